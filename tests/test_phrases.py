"""Device-side phrase splitter + greedy decoder (csrc/ctc_phrase.cuh) == the reference's StreamingLogprobSplitter +
GreedyCTCDecoder inside StreamingCTCPipeline.forward (tone/logprob_splitter.py:60-153, tone/decoder.py:57-59,
tone/pipeline.py:146-171), bit-exact: same phrases, same texts, same frame intervals, same buffer bookkeeping.

CPU: oracle/pipeline_oracle.py (full-logprob restatement) is pinned against the live reference classes, and
oracle/phrase_machine.py (statement-for-statement mirror of the kernel's incremental state machine) against both.
GPU: the kernel itself, alone on synthetic per-frame inputs and inside the step.
"""
import numpy as np
import pytest

import pipeline_oracle as po
import refimport
from phrase_machine import PhraseMachine

LABELS = po.LABELS


def _synthetic_logprobs(rng, n_frames, long_speech=False):
    """Frame posteriors with speech bursts and silences of assorted lengths around the 20-frame rule; repeated tokens,
    blank- and space-dominated silences, some near the 0.9 threshold."""
    lp = np.empty((n_frames, 35), dtype=np.float32)
    t = 0
    speech = bool(rng.integers(0, 2))
    while t < n_frames:
        if speech:
            n = int(rng.choice([30, 500, 1990, 2000, 2001, 2500, 4100])) if long_speech else int(rng.integers(1, 60))
        else:
            n = int(rng.choice([1, 2, 3, 5, 10, 19, 20, 21, 40, 75]))
        logits = rng.normal(0, 1.5, size=(n, 35)).astype(np.float32)
        if speech:
            idx = rng.integers(0, 34, n)
            for k in range(1, n):
                if rng.random() < 0.5:
                    idx[k] = idx[k - 1]
            logits[np.arange(n), idx] += 6.0
        else:
            logits[:, 34] += rng.choice([3.0, 8.0])
            if rng.random() < 0.3:
                logits[:, 33] += rng.choice([3.0, 8.0])
        lp[t:t + n] = (logits - np.log(np.exp(logits).sum(-1, keepdims=True)))[: n_frames - t]
        t += n
        speech = not speech
    return lp


def _reference_chain():
    """(splitter, decoder) of the live reference, or None when no reference tree is reachable."""
    tone = refimport.import_reference()
    if tone is None:
        return None
    return tone.logprob_splitter.StreamingLogprobSplitter(), tone.decoder.GreedyCTCDecoder()


def _want(lp, T, last_chunk_is_last=True):
    """Phrases (text, start_frame, end_frame) per chunk from the full-logprob chain (restated; live-pinned below)."""
    st, out = None, []
    n = len(lp) // T
    for i in range(n):
        ph, st = po.split(lp[i * T:(i + 1) * T], st, is_last=(last_chunk_is_last and i == n - 1))
        out.append([(po.greedy(p), a, b) for p, a, b in ph])
    return out, st


@pytest.mark.parametrize("seed", range(40))
def test_state_machine_equals_full_logprob_chain(seed):
    rng = np.random.default_rng(seed)
    T = int(rng.choice([10, 13]))
    long_ = seed % 5 == 0
    n = int(rng.integers(250, 600)) if long_ else int(rng.integers(5, 90))
    lp = _synthetic_logprobs(rng, T * n, long_)
    ref = _reference_chain()
    pm, st_o, st_r, total = PhraseMachine(), None, None, 0
    for i in range(n):
        c, last = lp[i * T:(i + 1) * T], i == n - 1
        ph, st_o = po.split(c, st_o, is_last=last)
        want = [(po.greedy(p), a, b) for p, a, b in ph]
        if ref is not None:                                   # the restatement itself against the unchanged reference
            rph, st_r = ref[0].forward(c, st_r, is_last=last)
            assert want == [(ref[1].forward(p.logprobs), p.start_frame, p.end_frame) for p in rph]
            assert (st_r.offset, len(st_r.past_logprobs)) == (st_o.offset, len(st_o.past))
        got = [("".join(LABELS[t] for t in ids), a, b) for ids, a, b in pm.step(c.argmax(-1), c[:, 33:35], is_last=last)]
        assert got == want
        assert (pm.offset, len(pm.buf)) == (st_o.offset, len(st_o.past))     # same trimming, chunk by chunk
        total += len(want)
    assert total >= 1 or n < 10


def test_forced_split_of_long_phrase():
    """MAX_PHRASE_DURATION = 2000 frames forces a cut (tone/logprob_splitter.py:84-86)."""
    lp = np.full((2300, 35), -20.0, dtype=np.float32)
    lp[:, 5] = 0.0                                           # 2250 frames of one speech token
    lp[2250:, 5], lp[2250:, 34] = -20.0, 0.0                 # then silence
    pm, got = PhraseMachine(), []
    for i in range(230):
        c = lp[i * 10:(i + 1) * 10]
        got += [("".join(LABELS[t] for t in ids), a, b) for ids, a, b in pm.step(c.argmax(-1), c[:, 33:35], is_last=(i == 229))]
    want, _ = _want(lp, 10)
    assert got == [p for chunk in want for p in chunk] and len(got) == 2
    assert got[0][1:] == (0, 2000)


@pytest.mark.skipif(refimport.reference_root() is None, reason="no reference tree reachable")
def test_restatement_matches_unchanged_reference_pipeline():
    """pipeline_oracle == the reference's own StreamingCTCPipeline with a canned acoustic model."""
    tone = refimport.import_reference()
    rng = np.random.default_rng(3)
    lp = _synthetic_logprobs(rng, 600)

    class Canned:                                            # duck-typed model, as dev/triton/client_wer.py does
        def __init__(self):
            self.i = 0

        def forward(self, chunk, state):
            out = lp[self.i * 10:(self.i + 1) * 10][None]
            self.i += 1
            return out, state

    pipe = tone.pipeline.StreamingCTCPipeline(Canned(), tone.logprob_splitter.StreamingLogprobSplitter(),
                                              tone.decoder.GreedyCTCDecoder())
    state, st, got, want = None, None, [], []
    for i in range(60):
        ph, state = pipe.forward(np.zeros(2400, dtype=np.int32), state, is_last=(i == 59))
        got += [(p.text, p.start_time, p.end_time) for p in ph]
        r, st = po.pipeline_forward(lp[i * 10:(i + 1) * 10], st, is_last=(i == 59))
        want += r
    assert got == want and len(got) >= 1


# ------------------------------------------------------------------------------------------------ GPU
def _device_phrases(eng, slots, lp, T, flush_last=True):
    """Run the kernel alone over B streams of synthetic log-probs -> per stream list of (text, start, end)."""
    B, n = lp.shape[0], lp.shape[1] // T
    got = [[] for _ in range(B)]
    for i in range(n):
        c = lp[:, i * T:(i + 1) * T]
        last = np.full(B, 1 if (flush_last and i == n - 1) else 0, dtype=np.uint8)
        for b, a, e_, ids in eng.selftest_phrases(slots, c.argmax(-1), c[:, :, 33:35], last):
            got[b].append(("".join(LABELS[t] for t in ids), a, e_))
    return got


@pytest.mark.gpu
@pytest.mark.parametrize("T", [10, 13])
def test_phrase_kernel_equals_reference_chain(tb, weights, T):
    """300 streams at once (more than the kernel's 256 threads: the strided loop), assorted seeds, incl. forced splits."""
    B, n = 300, 260
    eng = tb.Engine(weights, chunk_samples=2400, max_slots=B, max_batch=B)
    try:
        lp = np.stack([_synthetic_logprobs(np.random.default_rng(1000 + b), T * n, long_speech=(b % 7 == 0)) for b in range(B)])
        slots = eng.alloc_slots(B)
        got = _device_phrases(eng, slots, lp, T)
        ref = _reference_chain()
        total = 0
        for b in range(B):
            want, _ = _want(lp[b], T)
            flat = [p for chunk in want for p in chunk]
            assert got[b] == flat, f"stream {b}"
            total += len(flat)
            if ref is not None and b < 8:                    # and directly against the unchanged reference classes
                st, live = None, []
                for i in range(n):
                    ph, st = ref[0].forward(lp[b, i * T:(i + 1) * T], st, is_last=(i == n - 1))
                    live += [(ref[1].forward(p.logprobs), p.start_frame, p.end_frame) for p in ph]
                assert got[b] == live
        assert total > B
        # a reset slot starts from an empty splitter state again
        eng.reset_slots(slots[:4])
        again = _device_phrases(eng, slots[:4], lp[:4], T)
        assert again == got[:4]
    finally:
        eng.close()


@pytest.mark.gpu
def test_phrase_kernel_forced_split(tb, weights):
    lp = np.full((1, 2300, 35), -20.0, dtype=np.float32)
    lp[:, :, 5] = 0.0
    lp[:, 2250:, 5], lp[:, 2250:, 34] = -20.0, 0.0
    eng = tb.Engine(weights, chunk_samples=2400, max_slots=2, max_batch=2)
    try:
        got = _device_phrases(eng, eng.alloc_slots(1), lp, 10)[0]
        want, _ = _want(lp[0], 10)
        assert got == [p for chunk in want for p in chunk] and len(got) == 2 and got[0][1:] == (0, 2000)
    finally:
        eng.close()


@pytest.mark.gpu
def test_greedy_pipeline_on_gpu_matches_full_logprob_path(tb, weights):
    """GreedyStreamingPipeline (phrases decided on the device, inside the step graph) == Engine.step (full log-probs) +
    the restated reference chain on the host, stream by stream."""
    B, C, n = 6, 2400, 60
    eng = tb.Engine(weights, chunk_samples=C, max_slots=3 * B, max_batch=B)
    try:
        pcm = tb.synth.telephony_pcm(B, C * n, seed=42)
        pipe = tb.greedy.GreedyStreamingPipeline(eng, B)
        ref_slots = eng.alloc_slots(B)
        states, want, got = [None] * B, [[] for _ in range(B)], [[] for _ in range(B)]
        for i in range(n):
            chunk = pcm[:, i * C:(i + 1) * C]
            last = i == n - 1
            lp, _ = eng.step(ref_slots, chunk)
            for b in range(B):
                r, states[b] = po.pipeline_forward(lp[b], states[b], is_last=last)
                want[b] += r
            for b, ph in enumerate(pipe.forward(chunk, is_last=last)):
                got[b] += [(p.text, p.start_time, p.end_time) for p in ph]
        assert got == want
        assert sum(len(w) for w in want) >= 1
        # offline form (pipelined submit / collect) gives the same phrases on fresh streams
        pipe2 = tb.greedy.GreedyStreamingPipeline(eng, B)
        audio = pcm[:, 2400:-2400 - 37]                       # forward_offline pads 2400 both sides and to a chunk multiple
        off = pipe2.forward_offline(audio)
        assert all(isinstance(p.text, str) for ph in off for p in ph)
    finally:
        eng.close()


# ------------------------------------------------------------------------------------------------ property tests
from hypothesis import given, settings, strategies as st_  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(runs=st_.lists(st_.tuples(st_.booleans(), st_.integers(1, 70)), min_size=1, max_size=40),
       T=st_.sampled_from([10, 13]), seed=st_.integers(0, 2 ** 16), flush=st_.booleans())
def test_state_machine_property(runs, T, seed, flush):
    """Any alternation of speech / silence runs (lengths around the 20-frame rule and the 3-frame expansion), any chunking,
    with or without a final flush: the incremental state machine == the reference formulation on full log-probs."""
    rng = np.random.default_rng(seed)
    frames = []
    for speech, n in runs:
        logits = rng.normal(0, 1.0, size=(n, 35)).astype(np.float32)
        if speech:
            logits[np.arange(n), rng.integers(0, 34, n)] += 7.0
        else:
            logits[:, 34] += 9.0
        frames.append(logits - np.log(np.exp(logits).sum(-1, keepdims=True)))
    lp = np.concatenate(frames)
    n = len(lp) // T
    if n == 0:
        return
    pm, st = PhraseMachine(), None
    for i in range(n):
        c, last = lp[i * T:(i + 1) * T], flush and i == n - 1
        ph, st = po.split(c, st, is_last=last)
        want = [(po.greedy(p), a, b) for p, a, b in ph]
        got = [("".join(LABELS[t] for t in ids), a, b) for ids, a, b in pm.step(c.argmax(-1), c[:, 33:35], is_last=last)]
        assert got == want and (pm.offset, len(pm.buf)) == (st.offset, len(st.past))
