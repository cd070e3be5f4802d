"""The oracle pinned against the reference: committed golden vectors (always) and the live
reference tree (only when /root/reference is mounted, i.e. in the build container)."""
import os
import sys

import numpy as np
import pytest
import torch

import tone_oracle as orc

import refimport


def _run(W, pcm, C, quant=None):
    B = pcm.shape[0]
    st = orc.zero_state(B)
    outs = []
    for i in range(pcm.shape[1] // C):
        lp, st = orc.step(W, torch.from_numpy(pcm[:, i * C:(i + 1) * C].astype(np.int32)), st, quant)
        outs.append(lp.numpy())
    return np.stack(outs, 0), st


@pytest.mark.parametrize("ms,C", [(300, 2400), (400, 3200)])
def test_oracle_matches_reference_golden(tb, weights, golden, ms, C):
    g = golden[ms]
    assert bytes(g["weights_digest"]).decode() == tb.weights.digest(weights), "seeded weights drifted"
    W = orc.to_torch(weights)
    lp, st = _run(W, g["pcm"].astype(np.int32), C)
    assert lp.shape == g["logprobs"].shape
    # fp32 vs fp32, different summation order only
    np.testing.assert_allclose(lp, g["logprobs"], atol=1e-4, rtol=0)
    assert (lp.argmax(-1) == g["logprobs"].argmax(-1)).all()
    # the reference's own fp16-autocast export path sits within its documented spread of the fp32 graph
    assert np.abs(lp - g["logprobs_export"]).max() < 5e-2
    for k in orc.STATE_KEYS:
        ref = g["state_" + k]
        got = st[k].numpy()[:1].reshape(ref.shape)
        if k == "mhsa_len":
            assert (got == ref).all() and (st[k].numpy() == g["state16_mhsa_len"].reshape(-1)).all()
        else:  # fixture stored as fp16: half-ulp of |x|<8 is 2^-9*... -> 4e-3 is ample
            np.testing.assert_allclose(got, ref.astype(np.float32), atol=4e-3, rtol=0)


def test_state_size_and_shapes(tb):
    a = tb.DEFAULT_ARCH
    assert a.state_size == orc.STATE_SIZE == 219729          # tone/onnx_wrapper.py:34
    assert tuple(s for _, s in a.state_layout()) == orc.STATE_SHAPES
    assert tb.weights.n_params() == 71685347                  # SURVEY.md quick facts [verified]


def test_pack_unpack_roundtrip():
    st = orc.zero_state(2)
    g = torch.Generator().manual_seed(3)
    for k in st:
        if k != "mhsa_len":
            st[k] = torch.randn(st[k].shape, generator=g).half().float()
    st["mhsa_len"] = torch.tensor([7, 30])
    flat = orc.pack_state(st)
    assert flat.shape == (2, 219729) and flat.dtype == np.float16
    back = orc.unpack_state(flat)
    for k in st:
        assert torch.equal(back[k], st[k]), k


def test_initial_state_is_zero_and_len_progression(weights):
    W = orc.to_torch(weights)
    st = orc.zero_state(1)
    pcm = torch.zeros(1, 2400, dtype=torch.int32)
    lens = []
    for _ in range(4):
        _, st = orc.step(W, pcm, st)
        lens.append(int(st["mhsa_len"][0]))
    assert lens == [10, 20, 30, 30]                           # conformer_blocks.py:191
    # layer-14 slot keeps its 15 left-pad rows at zero (conformer_blocks.py:161-163)
    assert float(st["mhsa"][:, 0, :15].abs().max()) == 0.0


def test_stream_independence(weights, tb):
    """Permuting the batch permutes the outputs: streams never interact."""
    W = orc.to_torch(weights)
    pcm = tb.synth.telephony_pcm(3, 4800, seed=5)
    a, _ = _run(W, pcm, 2400)
    b, _ = _run(W, pcm[[2, 0, 1]], 2400)
    np.testing.assert_allclose(a[:, [2, 0, 1]], b, atol=2e-5, rtol=0)


def test_bf16_emulation_within_stated_tolerance(weights, tb):
    """The tolerance the GPU parity tests state (logprobs 0.06, state 0.06) covers operand rounding to bf16."""
    W = orc.to_torch(weights)
    pcm = tb.synth.telephony_pcm(2, 2400 * 4, seed=11)
    a, sa = _run(W, pcm, 2400)
    b, sb = _run(W, pcm, 2400, quant=orc.bf16_round)
    assert np.abs(a - b).max() < 0.06
    for k in orc.STATE_KEYS:
        if k != "mhsa_len":
            assert float((sa[k] - sb[k]).abs().max()) < 0.06


def test_greedy_text():
    lp = np.full((6, 35), -10.0, dtype=np.float32)
    for t, tok in enumerate([34, 0, 0, 34, 0, 33]):          # blank a a blank a space
        lp[t, tok] = 0.0
    assert orc.greedy_text(lp) == "аа"                        # tone/decoder.py:57-59


def test_frontend_constants_match_torchaudio(tb):
    torchaudio = pytest.importorskip("torchaudio")
    fb = torchaudio.functional.melscale_fbanks(81, 0.0, 4000.0, 64, 8000, norm="slaney", mel_scale="slaney").T
    assert float((fb - orc.mel_fb()).abs().max()) < 1e-6
    assert float((fb - torch.from_numpy(tb.weights.mel_filterbank())).abs().max()) < 1e-6
    assert int((fb != 0).sum()) == 156


@pytest.mark.skipif(refimport.reference_root() is None, reason="no reference tree reachable")
def test_oracle_matches_live_reference(tb, weights):
    """The unmodified reference torch model (from /root/reference or its installed copy baseline/_ref), run live."""
    pcm = tb.synth.telephony_pcm(2, 2400 * 3, seed=77)
    model = refimport.ReferenceStreamingModel(weights)
    state, outs = None, []
    for i in range(3):
        lp_ref, state = model.forward(np.ascontiguousarray(pcm[:, i * 2400:(i + 1) * 2400, None]).astype(np.int32), state)
        outs.append(lp_ref)
    lp, _ = _run(orc.to_torch(weights), pcm, 2400)
    np.testing.assert_allclose(lp, np.stack(outs, 0), atol=1e-4, rtol=0)


def test_feature_input_mode_matches_reference_skip_preprocessor(weights):
    """Feature-input mode (reference skip_preprocessor=True, tone/nn/model.py:151-160).  The reference only runs this
    mode the way it is exported (fp16 features -> fp16 autocast), so the bar is the fp16-autocast spread measured on the
    audio goldens (1.8e-2 between their fp32 and export-mode logprobs), not the 1e-4 of the fp32 goldens."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "features_300ms.npz"))
    W = orc.to_torch(weights)
    st = orc.zero_state(g["feats"].shape[1])
    pre0 = st["preproc"].clone()
    for i in range(g["feats"].shape[0]):
        lp, st = orc.step(W, None, st, feats=torch.from_numpy(g["feats"][i].astype(np.float32)))
        assert np.abs(lp.numpy() - g["logprobs"][i]).max() <= 3e-2
    assert (st["mhsa_len"].numpy().ravel() == g["state_mhsa_len"].ravel()).all()
    assert torch.equal(st["preproc"], pre0)          # the waveform state is not touched in this mode
