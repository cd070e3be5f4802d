"""CPU mirror of the device-side phrase splitter + greedy decoder (t-one_b200/csrc/ctc_phrase.cuh).
TEST INFRASTRUCTURE ONLY (same rule as tone_oracle.py).

The CUDA kernel replaces the reference's per-chunk "re-derive the silence runs of the whole buffer" formulation
(tone/logprob_splitter.py:60-153) by an incremental state machine.  This file is that state machine, statement for
statement, in Python, so that the ALGORITHM can be pinned on the CPU against the reference's own
StreamingLogprobSplitter + GreedyCTCDecoder (tests/test_phrases.py, live when /root/reference is mounted, else against
oracle/pipeline_oracle.py); the GPU test then pins the KERNEL against the same chain.
"""
from __future__ import annotations

import numpy as np

MIN_SILENCE, EXPAND, MAX_PHRASE, BLANK, SPACE = 20, 3, 2000, 34, 33


class PhraseMachine:
    def __init__(self):
        self.buf = []          # (token, speech) since `offset`
        self.s = -1
        self.run = 0
        self.offset = 0

    @staticmethod
    def is_speech(sil: np.ndarray) -> np.ndarray:
        """sil (T,2) float32 log-probs of ' ' and blank -> bool (T,)   (tone/logprob_splitter.py:129)."""
        e = np.exp(sil.astype(np.float64)).astype(np.float32)
        return (e[:, 0] + e[:, 1]) <= np.float32(0.9)

    def step(self, tokens, sil, is_last=False):
        """-> list of (label_ids, start_frame, end_frame) finished by this chunk."""
        out = []
        sp = self.is_speech(np.asarray(sil, dtype=np.float32))
        n_before = len(self.buf)
        self.buf += [(int(t), bool(s)) for t, s in zip(tokens, sp)]
        n = len(self.buf)
        last_end = 0

        def emit(start, end):
            nonlocal last_end
            lo, hi = max(0, start - EXPAND), min(n, end + EXPAND)
            ids, prev = [], -1
            for i in range(lo, hi):
                t = self.buf[i][0]
                if t != prev and t < BLANK and not (not ids and t == SPACE):
                    ids.append(t)
                prev = t
            while ids and ids[-1] == SPACE:
                ids.pop()
            out.append((ids, self.offset + start, self.offset + end))
            last_end = end

        def emit_completed(s, e):
            while e - s >= MAX_PHRASE:
                emit(s, s + MAX_PHRASE)
                s += MAX_PHRASE
            emit(s, e)

        for t in range(n - n_before):
            i = n_before + t
            if sp[t]:
                if self.s < 0:
                    self.s = i
                self.run = 0
            else:
                self.run += 1
                if self.s >= 0 and self.run == MIN_SILENCE:
                    emit_completed(self.s, i + 1 - MIN_SILENCE)
                    self.s = -1
        if self.s >= 0:
            if is_last:
                emit_completed(self.s, n - self.run)
                self.s = -1
            else:
                split = False
                while n - self.s >= MAX_PHRASE:
                    emit(self.s, self.s + MAX_PHRASE)
                    self.s += MAX_PHRASE
                    split = True
                if split:
                    i = self.s
                    while i < n and not self.buf[i][1]:
                        i += 1
                    self.s = i if i < n else -1
        cut = last_end
        if self.s < 0:
            cut = max(cut, n - EXPAND)
        cut = max(cut, 0)
        if cut > 0:
            self.buf = self.buf[cut:]
            self.offset += cut
            if self.s >= 0:
                self.s -= cut
        if self.s < 0:
            self.run = min(self.run, len(self.buf))
        return out
