"""Greedy fast path above the acoustic model: phrase splitting and greedy CTC decoding on what the device already
reduced - per frame the argmax token and the (space, blank) log-probs - instead of the full (T, 35) log-probs.

Behaviourally identical to the reference's ``StreamingLogprobSplitter`` + ``GreedyCTCDecoder`` chain inside
``StreamingCTCPipeline.forward`` (reference: tone/logprob_splitter.py:60-153, tone/decoder.py:57-59,
tone/pipeline.py:147-172): the splitter only ever looks at ``exp(lp[33]) + exp(lp[34])`` and greedy decoding only at
``argmax``, so nothing else has to leave the GPU (12 instead of 140 bytes per frame).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .arch import LABELS

SILENCE_THRESHOLD = 0.9       # tone/logprob_splitter.py:55-58
MIN_SILENCE_DURATION = 20
SPEECH_EXPAND_SIZE = 3
MAX_PHRASE_DURATION = 2000
FRAME_SIZE, MEAN_TIME_BIAS, SAMPLE_RATE, PADDING = 0.03, 0.33, 8000, 2400   # tone/onnx_wrapper.py:30-33, pipeline.py:48


@dataclass
class TextPhrase:             # tone/pipeline.py:18-31
    text: str
    start_time: float
    end_time: float


@dataclass
class GreedyStreamState:
    past_tokens: np.ndarray = field(default_factory=lambda: np.zeros((0,), dtype=np.int32))
    past_sil: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), dtype=np.float32))
    offset: int = 0


def collapse(tokens) -> str:
    """tone/decoder.py:57-59: drop repeats, drop ids >= len(LABELS) (blank), join, strip."""
    out, prev = [], None
    for t in tokens:
        t = int(t)
        if t != prev:
            out.append(t)
        prev = t
    return "".join(LABELS[t] for t in out if t < len(LABELS)).strip()


def _iterate_over_phrases(is_speech: np.ndarray, is_last: bool):
    """tone/logprob_splitter.py:60-89 restated: phrases are the speech runs between silences of >= 20 frames."""
    n = len(is_speech)
    pad = MIN_SILENCE_DURATION
    sp = np.pad(is_speech, (pad, pad if is_last else 0))
    changes = np.diff(np.pad(~sp, (1, 1)).astype(np.int32))
    starts = (changes == 1).nonzero()[0] - pad
    ends = (changes == -1).nonzero()[0] - pad
    keep = (ends - starts) >= MIN_SILENCE_DURATION
    starts, ends = starts[keep], ends[keep]
    speech_starts, speech_ends = ends.tolist(), starts.tolist()[1:] + [n]
    for i, (s, e) in enumerate(zip(speech_starts, speech_ends)):
        while e - s >= MAX_PHRASE_DURATION:
            yield s, s + MAX_PHRASE_DURATION
            s += MAX_PHRASE_DURATION
        if i < len(ends) - 1:
            yield s, e


class GreedyPhraseSplitter:
    """One stream's splitter + greedy decoder on (tokens, sil_logprobs)."""

    def forward(self, tokens: np.ndarray, sil: np.ndarray, state: Optional[GreedyStreamState] = None, *,
                is_last: bool = False) -> Tuple[List[Tuple[str, int, int]], GreedyStreamState]:
        if state is None:
            state = GreedyStreamState()
        tok = np.concatenate((state.past_tokens, np.asarray(tokens, dtype=np.int32)))
        sl = np.concatenate((state.past_sil, np.asarray(sil, dtype=np.float32)), axis=0)
        is_speech = np.exp(sl).sum(axis=-1) <= SILENCE_THRESHOLD          # logprob_splitter.py:129
        phrases, last = [], 0
        for ps, pe in _iterate_over_phrases(is_speech, is_last):
            text = collapse(tok[max(0, ps - SPEECH_EXPAND_SIZE): pe + SPEECH_EXPAND_SIZE])
            phrases.append((text, ps + state.offset, pe + state.offset))
            last = pe
        if not len(np.nonzero(is_speech[last:])[0]):                      # logprob_splitter.py:146-148
            last = max(last, len(tok) - SPEECH_EXPAND_SIZE)
        return phrases, GreedyStreamState(tok[last:], sl[last:], state.offset + last)


def to_text_phrase(text: str, start_frame: int, end_frame: int) -> TextPhrase:
    """Frame indices -> seconds exactly as tone/pipeline.py:151-165."""
    start = max(0, round(start_frame * FRAME_SIZE - MEAN_TIME_BIAS - PADDING / SAMPLE_RATE, 2))
    end = max(start, round(end_frame * FRAME_SIZE - MEAN_TIME_BIAS - PADDING / SAMPLE_RATE, 2))
    return TextPhrase(text, start, end)


class GreedyStreamingPipeline:
    """Batched counterpart of ``StreamingCTCPipeline`` with ``GreedyCTCDecoder`` over an :class:`Engine`:
    ``forward(pcm (B, chunk))`` advances B resident streams and returns the phrases finished by this chunk."""

    def __init__(self, engine, n_streams: int):
        self.engine = engine
        self.slots = engine.alloc_slots(n_streams)
        self.states = [None] * n_streams
        self.splitter = GreedyPhraseSplitter()

    def forward(self, pcm: np.ndarray, *, is_last: bool = False) -> List[List[TextPhrase]]:
        tokens, sil = self.engine.step_greedy(self.slots, pcm)
        out = []
        for b in range(len(self.slots)):
            ph, self.states[b] = self.splitter.forward(tokens[b], sil[b], self.states[b], is_last=is_last)
            out.append([to_text_phrase(*p) for p in ph])
        return out

    def forward_offline(self, audio: np.ndarray) -> List[List[TextPhrase]]:
        """audio int32 (B, L): pad 2400 both sides, pad to a chunk multiple, stream (tone/pipeline.py:174-203)."""
        C = self.engine.chunk_samples
        a = np.pad(audio, ((0, 0), (PADDING, PADDING)))
        a = np.pad(a, ((0, 0), (0, -a.shape[1] % C)))
        n = a.shape[1] // C
        res = [[] for _ in range(len(self.slots))]
        for i in range(n):
            for b, ph in enumerate(self.forward(a[:, i * C:(i + 1) * C], is_last=(i == n - 1))):
                res[b].extend(ph)
        return res

    def close(self):
        self.engine.release_slots(self.slots)
