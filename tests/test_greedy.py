"""Greedy fast path (tokens + silence log-probs) == the reference's splitter + greedy decoder on full log-probs."""
import importlib.machinery
import os
import sys
import types

import numpy as np
import pytest

import pipeline_oracle as po

REF = os.environ.get("TONE_REFERENCE", "/root/reference")


def _synthetic_logprobs(rng, n_frames):
    """Frame posteriors with speech bursts and silences of assorted lengths around the 20-frame rule."""
    lp = np.empty((n_frames, 35), dtype=np.float32)
    t = 0
    speech = bool(rng.integers(0, 2))
    while t < n_frames:
        n = int(rng.choice([1, 2, 5, 19, 20, 21, 40, 75])) if not speech else int(rng.integers(1, 60))
        logits = rng.normal(0, 1.5, size=(n, 35)).astype(np.float32)
        if speech:
            logits[np.arange(n), rng.integers(0, 33, n)] += 6.0
        else:
            logits[:, 34] += rng.choice([3.0, 8.0])       # blank dominated, sometimes near the 0.9 threshold
        lp[t:t + n] = (logits - np.log(np.exp(logits).sum(-1, keepdims=True)))[: n_frames - t]
        t += n
        speech = not speech
    return lp


@pytest.mark.parametrize("seed", range(6))
def test_token_level_splitter_equals_logprob_level(tb, seed):
    rng = np.random.default_rng(seed)
    lp = _synthetic_logprobs(rng, 10 * int(rng.integers(30, 90)))
    sp = tb.greedy.GreedyPhraseSplitter()
    st_a, st_b, got, want = None, None, [], []
    n = len(lp) // 10
    for i in range(n):
        chunk = lp[i * 10:(i + 1) * 10]
        last = i == n - 1
        ref, st_a = po.pipeline_forward(chunk, st_a, is_last=last)
        want += ref
        ph, st_b = sp.forward(chunk.argmax(-1), chunk[:, 33:35], st_b, is_last=last)
        got += [(p.text, p.start_time, p.end_time) for p in (tb.greedy.to_text_phrase(*x) for x in ph)]
        assert st_b.offset == st_a.offset and len(st_b.past_tokens) == len(st_a.past)
    assert got == want
    assert len(want) >= 1


def test_forced_split_of_long_phrase(tb):
    """MAX_PHRASE_DURATION = 2000 frames forces a cut (tone/logprob_splitter.py:84-86)."""
    lp = np.full((2300, 35), -20.0, dtype=np.float32)
    lp[:, 5] = 0.0                                           # 2300 frames of one speech token
    lp[2250:, 5], lp[2250:, 34] = -20.0, 0.0                 # then silence
    sp = tb.greedy.GreedyPhraseSplitter()
    ref, _ = po.pipeline_forward(lp, None, is_last=True)
    ph, _ = sp.forward(lp.argmax(-1), lp[:, 33:35], None, is_last=True)
    got = [(p.text, p.start_time, p.end_time) for p in (tb.greedy.to_text_phrase(*x) for x in ph)]
    assert got == ref and len(ref) == 2


def _import_reference():
    """SURVEY 8c recipe (b): stub the absent third-party modules, then import the unchanged reference package."""
    for name in ("onnxruntime", "pyctcdecode", "pyctcdecode.decoder", "huggingface_hub"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__spec__ = importlib.machinery.ModuleSpec(name, None)
                sys.modules[name] = m
    sys.modules["onnxruntime"].InferenceSession = getattr(sys.modules["onnxruntime"], "InferenceSession", object)
    sys.modules["onnxruntime"].SessionOptions = getattr(sys.modules["onnxruntime"], "SessionOptions", object)
    d = sys.modules["pyctcdecode.decoder"]
    d.BeamSearchDecoderCTC = getattr(d, "BeamSearchDecoderCTC", object)
    d.build_ctcdecoder = getattr(d, "build_ctcdecoder", lambda *a, **k: None)
    hub = sys.modules["huggingface_hub"]
    hub.hf_hub_download = getattr(hub, "hf_hub_download", lambda *a, **k: None)
    sys.modules.pop("tone", None)
    sys.path.insert(0, REF)
    try:
        from tone.decoder import GreedyCTCDecoder
        from tone.logprob_splitter import StreamingLogprobSplitter
        from tone.pipeline import StreamingCTCPipeline
    finally:
        sys.path.remove(REF)
    return StreamingCTCPipeline, StreamingLogprobSplitter, GreedyCTCDecoder


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tone")), reason="reference tree not mounted")
def test_restatement_matches_unchanged_reference_pipeline():
    """pipeline_oracle == the reference's own StreamingCTCPipeline with a canned acoustic model."""
    Pipeline, Splitter, Greedy = _import_reference()
    rng = np.random.default_rng(3)
    lp = _synthetic_logprobs(rng, 600)

    class Canned:                                            # duck-typed model, as dev/triton/client_wer.py does
        def __init__(self):
            self.i = 0

        def forward(self, chunk, state):
            out = lp[self.i * 10:(self.i + 1) * 10][None]
            self.i += 1
            return out, state

    pipe = Pipeline(Canned(), Splitter(), Greedy())
    state, st, got, want = None, None, [], []
    for i in range(60):
        ph, state = pipe.forward(np.zeros(2400, dtype=np.int32), state, is_last=(i == 59))
        got += [(p.text, p.start_time, p.end_time) for p in ph]
        r, st = po.pipeline_forward(lp[i * 10:(i + 1) * 10], st, is_last=(i == 59))
        want += r
    assert got == want and len(got) >= 1


@pytest.mark.gpu
def test_greedy_pipeline_on_gpu_matches_full_logprob_path(tb, weights):
    """Engine.step_greedy + token-level splitter == Engine.step (full log-probs) + the restated reference chain."""
    B, C, n = 4, 2400, 40
    eng = tb.Engine(weights, chunk_samples=C, max_slots=2 * B, max_batch=B)
    pcm = tb.synth.telephony_pcm(B, C * n, seed=42)
    pipe = tb.greedy.GreedyStreamingPipeline(eng, B)
    ref_slots = eng.alloc_slots(B)
    states, want, got = [None] * B, [[] for _ in range(B)], [[] for _ in range(B)]
    for i in range(n):
        chunk = pcm[:, i * C:(i + 1) * C]
        last = i == n - 1
        lp, tk = eng.step(ref_slots, chunk)
        for b in range(B):
            r, states[b] = po.pipeline_forward(lp[b], states[b], is_last=last)
            want[b] += r
        for b, ph in enumerate(pipe.forward(chunk, is_last=last)):
            got[b] += [(p.text, p.start_time, p.end_time) for p in ph]
    assert got == want
    assert sum(len(w) for w in want) >= 1
    eng.close()
