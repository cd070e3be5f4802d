"""BASELINE configs[0]: the UNCHANGED reference pipeline - StreamingCTCPipeline.forward_offline with
StreamingLogprobSplitter + GreedyCTCDecoder (tone/pipeline.py:174-203) - on tone/demo/audio_examples/audio_short.flac
(50,880 samples -> 24 chunks), with B200StreamingCTCModel plugged in where the reference plugs StreamingCTCModel, against
the same pipeline fed by the reference's own torch model on the CPU.  The reference package is imported from
/root/reference or from its installed copy baseline/_ref (which is what travels to the GPU box).

Bars: log-probs within the stated tolerance (0.06 above a reference log-prob of -6, 0.08 above -10, 0.10 below) chunk by chunk; per-frame greedy tokens identical wherever the reference's top-2 margin
exceeds LP_TOL; phrase count and timings identical when no frame of the reference sits within the tolerance of the
splitter's 0.9 silence threshold (asserted to hold for this recording); texts identical for phrases without an
undecided frame.  Weights are the seeded synthetic set (the HF checkpoint is not cached offline); the golden transcript
of the notebook is asserted only when the checkpoint is present.
"""
import importlib
import os

import numpy as np
import pytest

import refimport

LP_TOL = 0.06
GOLDEN_TRANSCRIPT = "ну сейчас к тебе приедет бригада давай давай я жду"   # examples/triton_request_example.ipynb cells 5-6


def _flac_path(name="audio_short.flac"):
    root = refimport.reference_root()
    p = os.path.join(root, "tone", "demo", "audio_examples", name) if root else None
    return p if p and os.path.isfile(p) else None


def _flac():
    return importlib.import_module("t-one_b200.flac")


@pytest.mark.skipif(_flac_path() is None, reason="reference example recordings not reachable")
@pytest.mark.parametrize("name,n", [("audio_short.flac", 50880), ("audio_long.flac", 406080)])
def test_flac_reader_decodes_reference_recordings(name, n):
    """Known answer: the MD5 of the decoded PCM stored in STREAMINFO (read_flac verifies it), length and format."""
    pcm, sr = _flac().read_flac(_flac_path(name), verify=True)
    assert sr == 8000 and pcm.dtype == np.int32 and pcm.shape == (n,)       # 6.36 s, 8 kHz / 16-bit / mono
    assert -32768 <= pcm.min() and pcm.max() <= 32767 and np.abs(pcm).max() > 1000


def test_flac_reader_rejects_corruption(tmp_path):
    p = _flac_path()
    if p is None:
        pytest.skip("reference example recordings not reachable")
    raw = bytearray(open(p, "rb").read())
    raw[len(raw) // 2] ^= 0x55                                # flip bits inside a frame: decode error or MD5 mismatch
    q = tmp_path / "bad.flac"
    q.write_bytes(bytes(raw))
    with pytest.raises((ValueError, IndexError)):
        _flac().read_flac(str(q), verify=True)
    with pytest.raises(ValueError, match="not a FLAC"):
        q.write_bytes(b"RIFF" + bytes(100))
        _flac().read_flac(str(q))


class _Recording:
    """Wraps a model with the reference interface and keeps the log-probs it returned."""

    def __init__(self, model):
        self.model, self.logprobs = model, []
        for k in ("SAMPLE_RATE", "MEAN_TIME_BIAS", "AUDIO_CHUNK_SAMPLES", "FRAME_SIZE"):
            setattr(self, k, getattr(model, k))

    def forward(self, chunk, state):
        lp, st = self.model.forward(chunk, state)
        self.logprobs.append(np.array(lp[0]))
        return lp, st


def _audio(tb):
    p = _flac_path()
    if p is not None:
        return _flac().read_flac(p)[0], "audio_short.flac"
    return tb.synth.telephony_pcm(1, 50880, seed=1234)[0].astype(np.int32), "synthetic 50,880 samples"


@pytest.mark.gpu
@pytest.mark.skipif(refimport.reference_root() is None, reason="no reference tree reachable (baseline/_ref not installed)")
@pytest.mark.parametrize("state_mode", ["device", "numpy"])
def test_unchanged_reference_pipeline_with_b200_model(tb, weights, state_mode):
    tone = refimport.import_reference()
    Pipeline, Splitter, Greedy = (tone.pipeline.StreamingCTCPipeline, tone.logprob_splitter.StreamingLogprobSplitter,
                                  tone.decoder.GreedyCTCDecoder)
    audio, what = _audio(tb)
    ours = _Recording(tb.B200StreamingCTCModel(weights, state_mode=state_mode, max_streams=4))
    ref = _Recording(refimport.ReferenceStreamingModel(weights))
    got = Pipeline(ours, Splitter(), Greedy()).forward_offline(audio)
    want = Pipeline(ref, Splitter(), Greedy()).forward_offline(audio)
    a, b = np.concatenate(ours.logprobs), np.concatenate(ref.logprobs)       # (24 * 10, 35)
    assert a.shape == b.shape == (240, 35), what
    err = np.abs(a - b)                                        # the stated tolerance (tests/test_gpu_parity.py): two tiers
    assert (err[b > -6.0] <= LP_TOL).all() and (err[b > -10.0] <= 0.08).all() and err.max() <= 0.10
    top2 = np.sort(b, axis=-1)[:, -2:]
    decided = (top2[:, 1] - top2[:, 0]) > LP_TOL
    assert decided.mean() > 0.5
    assert (a.argmax(-1) == b.argmax(-1))[decided].all()
    # phrase boundaries: a frame can only flip between speech and silence if its silence probability is within the
    # propagated tolerance of the 0.9 threshold (d p <= p * d logprob)
    p_sil = np.exp(b[:, 33]) + np.exp(b[:, 34])
    clear = np.abs(p_sil - 0.9) > LP_TOL * p_sil + 1e-6
    assert all(isinstance(p, tone.pipeline.TextPhrase) for p in got)
    if clear.all():
        assert [(p.start_time, p.end_time) for p in got] == [(p.start_time, p.end_time) for p in want]
        for g, w in zip(got, want):
            # frames the phrase was decoded from (the splitter widens the interval by 3 frames on both sides)
            f0 = 0 if w.start_time == 0 else int(round((w.start_time + 0.33 + 0.3) / 0.03)) - 5   # times are clamped / rounded
            f1 = int(round((w.end_time + 0.33 + 0.3) / 0.03)) + 5
            if decided[max(0, f0): f1].all():
                assert g.text == w.text
    assert len(got) >= 1 or len(want) == 0


@pytest.mark.gpu
def test_golden_transcript_with_the_real_checkpoint(tb):
    """examples/triton_request_example.ipynb cells 5-6: the only end-to-end text pin the reference has.  Needs the HF
    checkpoint t-tech/T-one in the local cache (there is no network here) and the reference recording."""
    if _flac_path() is None:
        pytest.skip("reference example recordings not reachable")
    try:
        model = tb.B200StreamingCTCModel.from_hugging_face(state_mode="device", max_streams=2)
    except Exception as ex:  # noqa: BLE001
        pytest.skip(f"HF checkpoint t-tech/T-one is not in the local cache: {type(ex).__name__}")
    tone = refimport.import_reference()
    pipe = tone.pipeline.StreamingCTCPipeline(model, tone.logprob_splitter.StreamingLogprobSplitter(), tone.decoder.GreedyCTCDecoder())
    audio, _ = _flac().read_flac(_flac_path())
    text = " ".join(p.text for p in pipe.forward_offline(audio))
    assert text == GOLDEN_TRANSCRIPT
