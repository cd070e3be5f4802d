"""CPU restatement of the reference's host-side chain above the acoustic model, on FULL log-probs.
TEST INFRASTRUCTURE ONLY (same rule as tone_oracle.py).

  splitter : tone/logprob_splitter.py:60-153 (StreamingLogprobSplitter)
  greedy   : tone/decoder.py:57-59          (GreedyCTCDecoder)
  timings  : tone/pipeline.py:141-171       (StreamingCTCPipeline.forward)

Pinned live against the reference's own classes in tests/test_greedy.py when /root/reference is mounted.
"""
from __future__ import annotations

import numpy as np

LABELS = "абвгдеёжзийклмнопрстуфхцчшщъыьэюя "


class SplitterState:
    def __init__(self):
        self.past = np.zeros((0, 35), dtype=np.float32)
        self.offset = 0


def _phrases(is_speech, is_last):
    n = len(is_speech)
    sp = np.pad(is_speech, (20, 20 if is_last else 0))
    ch = np.diff(np.pad(~sp, (1, 1)).astype(np.int32))
    st = (ch == 1).nonzero()[0] - 20
    en = (ch == -1).nonzero()[0] - 20
    keep = (en - st) >= 20
    assert keep[0]
    st, en = st[keep], en[keep]
    ss, se = en.tolist(), st.tolist()[1:] + [n]
    for i, (a, b) in enumerate(zip(ss, se)):
        while b - a >= 2000:
            yield a, a + 2000
            a += 2000
        if i < len(en) - 1:
            yield a, b


def split(logprobs, state=None, is_last=False):
    """-> (list of (phrase_logprobs, start_frame, end_frame), new state)"""
    if state is None:
        state = SplitterState()
    lp = np.concatenate((state.past, logprobs), axis=-2)
    is_speech = np.exp(lp[..., -2:]).sum(axis=-1) <= 0.9
    out, last = [], 0
    for a, b in _phrases(is_speech, is_last):
        out.append((lp[max(0, a - 3): b + 3], a + state.offset, b + state.offset))
        last = b
    if not len(np.nonzero(is_speech[last:])[0]):
        last = max(last, len(lp) - 3)
    ns = SplitterState()
    ns.past, ns.offset = lp[last:], state.offset + last
    return out, ns


def greedy(logprobs) -> str:
    toks = logprobs.argmax(axis=-1).tolist()
    out, prev = [], None
    for t in toks:
        if t != prev:
            out.append(t)
        prev = t
    return "".join(LABELS[t] for t in out if t < len(LABELS)).strip()


def pipeline_forward(logprobs, state=None, is_last=False):
    """One chunk of ONE stream: logprobs (T,35) -> (list of (text, start_s, end_s), state)."""
    phrases, ns = split(logprobs, state, is_last)
    res = []
    for lp, a, b in phrases:
        start = max(0, round(a * 0.03 - 0.33 - 2400 / 8000, 2))
        end = max(start, round(b * 0.03 - 0.33 - 2400 / 8000, 2))
        res.append((greedy(lp), start, end))
    return res, ns
