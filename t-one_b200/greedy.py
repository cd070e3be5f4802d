"""Greedy serving path above the acoustic model: the phrases the reference's host chain would produce, computed on the GPU.

In the reference, ``StreamingCTCPipeline.forward`` runs, per stream and chunk, ``StreamingLogprobSplitter`` and then
``GreedyCTCDecoder`` on the host (reference: tone/pipeline.py:146-171, tone/logprob_splitter.py:60-153,
tone/decoder.py:57-59).  Here the splitter state machine and the greedy collapse run in a kernel right after the decoder
epilogue (csrc/ctc_phrase.cuh); only finished phrases - frame interval + label ids - cross PCIe.  What is left for the
host is the mapping of label ids to characters and the frame -> seconds arithmetic of tone/pipeline.py:151-165.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .arch import LABELS
from .model import OUT_PHRASES

FRAME_SIZE, MEAN_TIME_BIAS, SAMPLE_RATE, PADDING = 0.03, 0.33, 8000, 2400   # tone/onnx_wrapper.py:30-33, pipeline.py:48


@dataclass
class TextPhrase:             # tone/pipeline.py:18-31
    text: str
    start_time: float
    end_time: float


def labels_to_text(ids) -> str:
    """Label ids (already collapsed, blank-free and stripped by the device) -> text (tone/decoder.py:23,59)."""
    return "".join(LABELS[int(t)] for t in ids)


def to_text_phrase(text: str, start_frame: int, end_frame: int) -> TextPhrase:
    """Frame indices -> seconds exactly as tone/pipeline.py:151-165."""
    start = max(0, round(start_frame * FRAME_SIZE - MEAN_TIME_BIAS - PADDING / SAMPLE_RATE, 2))
    end = max(start, round(end_frame * FRAME_SIZE - MEAN_TIME_BIAS - PADDING / SAMPLE_RATE, 2))
    return TextPhrase(text, start, end)


class GreedyStreamingPipeline:
    """Batched counterpart of ``StreamingCTCPipeline`` with ``GreedyCTCDecoder`` over an :class:`Engine`:
    ``forward(pcm (B, chunk))`` advances B resident streams and returns, per stream, the phrases finished by this chunk.
    ``submit`` / ``collect`` are the pipelined form (two chunks in flight: copies overlap the kernels)."""

    def __init__(self, engine, n_streams: int):
        self.engine = engine
        self.slots = engine.alloc_slots(n_streams)

    def submit(self, pcm: np.ndarray, *, is_last=False):
        B = len(self.slots)
        last = np.full(B, 1 if is_last else 0, dtype=np.uint8) if np.isscalar(is_last) else np.asarray(is_last, dtype=np.uint8)
        return self.engine.submit(self.slots, pcm, OUT_PHRASES, last)

    def collect(self, ticket) -> List[List[TextPhrase]]:
        out: List[List[TextPhrase]] = [[] for _ in range(len(self.slots))]
        for b, start, end, ids in self.engine.wait(ticket)["phrases"]:
            out[b].append(to_text_phrase(labels_to_text(ids), start, end))
        return out

    def forward(self, pcm: np.ndarray, *, is_last=False) -> List[List[TextPhrase]]:
        return self.collect(self.submit(pcm, is_last=is_last))

    def forward_offline(self, audio: np.ndarray) -> List[List[TextPhrase]]:
        """audio int (B, L): pad 2400 both sides, pad to a chunk multiple, stream (tone/pipeline.py:174-203)."""
        C = self.engine.chunk_samples
        a = np.pad(audio, ((0, 0), (PADDING, PADDING)))
        a = np.pad(a, ((0, 0), (0, -a.shape[1] % C)))
        n = a.shape[1] // C
        res: List[List[TextPhrase]] = [[] for _ in range(len(self.slots))]
        pending = None
        for i in range(n):
            t = self.submit(a[:, i * C:(i + 1) * C], is_last=(i == n - 1))
            if pending is not None:
                for b, ph in enumerate(self.collect(pending)):
                    res[b].extend(ph)
            pending = t
        for b, ph in enumerate(self.collect(pending)):
            res[b].extend(ph)
        return res

    def close(self):
        self.engine.release_slots(self.slots)
