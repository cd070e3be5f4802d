"""GPU parity tests (run with -m gpu on the B200 box): the CUDA step, called through the C ABI, against the CPU oracle
on the same seeded inputs and against the committed reference golden vectors.

Stated tolerance (floating point path, bf16 tensor-core operands with fp32 accumulation, fp32 norms / softmax):
  logprobs  |d| <= 0.10 vs the fp32 oracle / reference goldens (observed bf16 operand-rounding spread is 0.05,
            tests/test_oracle.py::test_bf16_emulation_within_stated_tolerance), <= 0.06 vs the bf16-emulating oracle
  state     |d| <= 0.10 (fp16 wire format, values up to ~4.5), mhsa_len exact
  tokens    identical wherever the oracle's top-2 logprob margin exceeds 0.10
"""
import numpy as np
import pytest
import torch

import tone_oracle as orc

pytestmark = pytest.mark.gpu

LP_TOL, LP_TOL_EMU, ST_TOL = 0.10, 0.06, 0.10


def _stream_oracle(W, pcm, C, quant=None):
    st = orc.zero_state(pcm.shape[0])
    outs = []
    for i in range(pcm.shape[1] // C):
        lp, st = orc.step(W, torch.from_numpy(pcm[:, i * C:(i + 1) * C].astype(np.int32)), st, quant)
        outs.append(lp.numpy())
    return np.stack(outs, 0), st


def _stream_engine(eng, slots, pcm, C):
    outs, toks = [], []
    for i in range(pcm.shape[1] // C):
        lp, tk = eng.step(slots, pcm[:, i * C:(i + 1) * C])
        outs.append(lp.copy())
        toks.append(tk.copy())
    return np.stack(outs, 0), np.stack(toks, 0)


def _check_tokens(tokens, lp_ref):
    top2 = np.sort(lp_ref, axis=-1)[..., -2:]
    margin = top2[..., 1] - top2[..., 0]
    decided = margin > LP_TOL
    assert decided.mean() > 0.5
    assert (tokens[decided] == lp_ref.argmax(-1)[decided]).all()


def _check_state(eng, slots, st_ref):
    flat_ref = orc.pack_state(st_ref).astype(np.float32)
    len_off = 80 + 2 * 30 * 384 + 16 * 384 * 30
    for b, s in enumerate(slots):
        got = eng.export_state(int(s)).astype(np.float32)
        assert got[len_off] == flat_ref[b, len_off]
        assert np.abs(got - flat_ref[b]).max() <= ST_TOL


@pytest.fixture(scope="module")
def engines(tb, weights):
    made = {}

    def get(C, **kw):
        key = (C, tuple(sorted(kw.items())))
        if key not in made:
            made[key] = tb.Engine(weights, chunk_samples=C, max_slots=80, max_batch=80, **kw)
        return made[key]

    yield get
    for e in made.values():
        e.close()


def test_library_is_native(tb):
    lib = tb.load_library()
    assert all(hasattr(lib, s) for s in tb.model.SYMBOLS)


@pytest.mark.parametrize("bn", [32, 64, 128])
def test_tcgen05_gemm_matches_matmul(engines, bn):
    eng = engines(2400)
    rng = np.random.default_rng(bn)
    for M, N, K in [(128, 128, 64), (640, 384, 1536), (333, 1152, 384)]:
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        ref = torch.from_numpy(A).bfloat16().float().numpy() @ torch.from_numpy(W).bfloat16().float().numpy().T
        got = eng.selftest_gemm(A, W, bn)
        np.testing.assert_allclose(got, ref, atol=2e-3, rtol=0)


@pytest.mark.parametrize("ms,C", [(300, 2400), (400, 3200)])
def test_step_matches_reference_golden(engines, golden, ms, C):
    """The committed outputs of the reference's own torch model (tests/golden/make_golden.py)."""
    g = golden[ms]
    eng = engines(C)
    pcm = g["pcm"].astype(np.int32)
    slots = eng.alloc_slots(pcm.shape[0])
    try:
        lp, tk = _stream_engine(eng, slots, pcm, C)
        assert lp.shape == g["logprobs"].shape
        assert np.isfinite(lp).all()
        assert np.abs(lp - g["logprobs"]).max() <= LP_TOL
        _check_tokens(tk, g["logprobs"])
        got = eng.export_state(int(slots[0])).astype(np.float32)
        ref = np.concatenate([g["state_" + k].astype(np.float32).reshape(-1) for k in orc.STATE_KEYS])
        assert np.abs(got - ref).max() <= ST_TOL
    finally:
        eng.release_slots(slots)


@pytest.mark.parametrize("C,B,n", [(2400, 5, 6), (3200, 4, 4), (2400, 64, 3)])
def test_step_matches_oracle(engines, weights, tb, C, B, n):
    eng = engines(C)
    W = orc.to_torch(weights)
    pcm = tb.synth.telephony_pcm(B, C * n, seed=100 + B)
    slots = eng.alloc_slots(B)
    try:
        lp, tk = _stream_engine(eng, slots, pcm, C)
        ref, st = _stream_oracle(W, pcm, C)
        emu, _ = _stream_oracle(W, pcm, C, quant=orc.bf16_round) if B <= 8 else (None, None)
        assert np.abs(lp - ref).max() <= LP_TOL
        if emu is not None:
            assert np.abs(lp - emu).max() <= LP_TOL_EMU
        _check_tokens(tk, ref)
        assert (tk == lp.argmax(-1)).all()                 # fused argmax == numpy argmax of our own logprobs
        np.testing.assert_allclose(np.exp(lp).sum(-1), 1.0, atol=1e-4)
        _check_state(eng, slots, st)
    finally:
        eng.release_slots(slots)


# ---- large-batch paths: 128-wide tiles (>= 2048 rows per lane), two concurrent lanes (>= 512 streams), and the
# persistent GEMM kernel.  Streams are independent, so B streams that replay a handful of distinct signals must all
# reproduce the oracle's answer for their signal (size-independent property; the oracle runs the distinct signals only).
@pytest.mark.parametrize("C,B,persist,mode", [(2400, 600, None, None), (2400, 600, "0", "0"), (2400, 600, "1", "0"), (2400, 600, "1", "1"),
                                               (2400, 600, "1", "2"), (3200, 420, "1", "2"), (2400, 230, "1", "2"),
                                               (2400, 333, "1", "1")])
def test_large_batch_paths_match_oracle(weights, tb, monkeypatch, C, B, persist, mode):
    # persist: smallest GEMM (in 128 x 128 tiles) that takes the persistent kernel, 0 = never; mode: its tile form
    # (0 = 128 wide, 1 = 256 wide, 2 = 256 wide on CTA pairs)
    if persist is not None:   # None = the engine's defaults
        monkeypatch.setenv("TONE_PERSIST_MIN_TILES", persist)
        monkeypatch.setenv("TONE_PERSIST_MODE", mode)
    n, D = 3, 6
    eng = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B)
    try:
        W = orc.to_torch(weights)
        distinct = tb.synth.telephony_pcm(D, C * n, seed=77)
        idx = np.arange(B) % D
        pcm = np.ascontiguousarray(distinct[idx])
        slots = eng.alloc_slots(B)
        lp, tk = _stream_engine(eng, slots, pcm, C)
        ref, st = _stream_oracle(W, distinct, C)
        assert np.isfinite(lp).all()
        assert np.abs(lp - ref[:, idx]).max() <= LP_TOL
        assert (tk == lp.argmax(-1)).all()
        _check_tokens(tk, ref[:, idx])
        flat_ref = orc.pack_state(st).astype(np.float32)
        for b in (0, 1, B // 2, B - 2, B - 1):
            got = eng.export_state(int(slots[b])).astype(np.float32)
            assert np.abs(got - flat_ref[idx[b]]).max() <= ST_TOL
    finally:
        eng.close()


# ---- experimental cluster (latency) path: the 16 layers + decoder as one thread-block-cluster kernel
@pytest.mark.parametrize("C,B,n,G", [(2400, 5, 5, None), (3200, 7, 4, None), (2400, 64, 3, None), (2400, 66, 2, "4"),
                                     (2400, 78, 2, "5"), (3200, 50, 2, "3")])
def test_cluster_path_matches_oracle(weights, tb, monkeypatch, C, B, n, G):
    """Same parity bar as the per-kernel path.  Cases cover: a ragged last group (B not a multiple of the group size),
    both group-size instantiations, and more groups than co-resident clusters (66 streams / 4 = 17 groups, 78 / 5 = 16,
    50 / 3 = 17 against 15 clusters), where clusters loop over groups."""
    if G is not None:
        monkeypatch.setenv("TONE_CLUSTER_G", G)
    eng = tb.Engine(weights, chunk_samples=C, max_slots=80, max_batch=80, cluster_max_batch=80)
    try:
        W = orc.to_torch(weights)
        pcm = tb.synth.telephony_pcm(B, C * n, seed=300 + B)
        slots = eng.alloc_slots(B)
        lp, tk = _stream_engine(eng, slots, pcm, C)
        assert eng._get_info().launches_per_step < 20       # the cluster kernel really ran
        ref, st = _stream_oracle(W, pcm, C)
        assert np.isfinite(lp).all()
        assert np.abs(lp - ref).max() <= LP_TOL
        _check_tokens(tk, ref)
        assert (tk == lp.argmax(-1)).all()
        _check_state(eng, slots, st)
    finally:
        eng.close()


def test_cluster_path_agrees_with_kernel_path_over_many_chunks(weights, tb):
    """Both device paths carry the same state layout: run 12 chunks on each (first three exercise the key masks) and
    compare logprobs and the exported state directly."""
    C, B, n = 2400, 9, 12
    pcm = tb.synth.telephony_pcm(B, C * n, seed=77)
    outs = []
    for cmb in (0, 80):
        eng = tb.Engine(weights, chunk_samples=C, max_slots=16, max_batch=16, cluster_max_batch=cmb)
        slots = eng.alloc_slots(B)
        lp, _ = _stream_engine(eng, slots, pcm, C)
        st = np.stack([eng.export_state(int(s)).astype(np.float32) for s in slots])
        outs.append((lp, st))
        eng.close()
    assert np.abs(outs[0][0] - outs[1][0]).max() <= LP_TOL_EMU
    assert np.abs(outs[0][1] - outs[1][1]).max() <= ST_TOL


def test_feature_input_mode(engines, weights, tb):
    """tone_step_features (reference skip_preprocessor=True): CUDA vs the oracle's feature path and vs the reference's
    own outputs (tests/golden/features_300ms.npz, export-mode reference: fp16-autocast spread on top of ours)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "features_300ms.npz"))
    n, B = g["feats"].shape[:2]
    eng = engines(2400)
    W = orc.to_torch(weights)
    slots = eng.alloc_slots(B)
    try:
        st = orc.zero_state(B)
        for i in range(n):
            lp, tk = eng.step_features(slots, g["feats"][i])
            ref, st = orc.step(W, None, st, feats=torch.from_numpy(g["feats"][i].astype(np.float32)))
            assert np.abs(lp - ref.numpy()).max() <= LP_TOL
            assert np.abs(lp - g["logprobs"][i]).max() <= LP_TOL
            _check_tokens(tk, ref.numpy())
        with pytest.raises(ValueError):
            eng.step_features(slots, g["feats"][0][:, :, :-1])
    finally:
        eng.release_slots(slots)


def test_simt_debug_path_agrees(engines, tb):
    """The SIMT debug GEMMs and the tcgen05 GEMMs see the same packed operands: results agree tightly."""
    C, B = 2400, 3
    pcm = tb.synth.telephony_pcm(B, C * 2, seed=9)
    a, b = engines(C), engines(C, gemm_impl=1, use_graph=False)
    sa, sb = a.alloc_slots(B), b.alloc_slots(B)
    try:
        la, _ = _stream_engine(a, sa, pcm, C)
        lb, _ = _stream_engine(b, sb, pcm, C)
        assert np.abs(la - lb).max() < 6e-2     # both sit within 0.03 of the oracle; rounding paths differ
    finally:
        a.release_slots(sa)
        b.release_slots(sb)


def test_graph_and_eager_are_bit_identical(engines, tb):
    C, B = 2400, 7
    pcm = tb.synth.telephony_pcm(B, C * 3, seed=21)
    a, b = engines(C), engines(C, use_graph=False)
    sa, sb = a.alloc_slots(B), b.alloc_slots(B)
    try:
        la, ta = _stream_engine(a, sa, pcm, C)
        lb, tb_ = _stream_engine(b, sb, pcm, C)
        assert np.array_equal(la, lb) and np.array_equal(ta, tb_)
    finally:
        a.release_slots(sa)
        b.release_slots(sb)


def test_streams_are_independent_and_slot_order_free(engines, tb):
    """Permuting the batch (and using different slots) permutes the outputs bit-exactly."""
    C, B = 2400, 6
    eng = engines(C)
    pcm = tb.synth.telephony_pcm(B, C * 3, seed=33)
    perm = np.array([4, 2, 0, 5, 1, 3])
    s1 = eng.alloc_slots(B)
    junk = eng.alloc_slots(3)
    s2 = eng.alloc_slots(B)[::-1].copy()
    try:
        l1, _ = _stream_engine(eng, s1, pcm, C)
        l2, _ = _stream_engine(eng, s2, pcm[perm], C)
        assert np.array_equal(l1[:, perm], l2)
    finally:
        eng.release_slots(np.concatenate([s1, junk, s2]))


def test_state_export_import_roundtrip_and_migration(engines, tb, weights):
    """export -> import into another slot continues the stream (checkpoint/resume, migration)."""
    C, B = 2400, 2
    eng = engines(C)
    pcm = tb.synth.telephony_pcm(B, C * 4, seed=55)
    s1 = eng.alloc_slots(B)
    s2 = eng.alloc_slots(B)
    try:
        _stream_engine(eng, s1, pcm[:, :2 * C], C)
        for b in range(B):
            st = eng.export_state(int(s1[b]))
            assert st.dtype == np.float16 and st.shape == (219729,)
            eng.import_state(int(s2[b]), st)
            assert np.array_equal(eng.export_state(int(s2[b])), st)      # wire format round-trips exactly
        la, _ = _stream_engine(eng, s1, pcm[:, 2 * C:], C)
        lb, _ = _stream_engine(eng, s2, pcm[:, 2 * C:], C)
        assert np.abs(la - lb).max() < 3e-2                              # only the fp16 quantisation of red/feat
    finally:
        eng.release_slots(np.concatenate([s1, s2]))


def test_model_class_numpy_state_matches_reference_contract(tb, weights):
    """forward(chunk, state) with the reference's numpy fp16 state, chained over chunks (tone/onnx_wrapper.py:84-123)."""
    m = tb.B200StreamingCTCModel(weights, state_mode="numpy", max_streams=4)
    W = orc.to_torch(weights)
    pcm = tb.synth.telephony_pcm(2, 2400 * 3, seed=77)
    state, st = None, orc.zero_state(2)
    for i in range(3):
        chunk = pcm[:, i * 2400:(i + 1) * 2400]
        out = m.forward(chunk[:, :, None].astype(np.int32), state)
        assert isinstance(out, list) and len(out) == 2
        lp, state = out
        assert lp.shape == (2, 10, 35) and lp.dtype == np.float32
        assert state.shape == (2, 219729) and state.dtype == np.float16
        ref, st = orc.step(W, torch.from_numpy(chunk), st)
        assert np.abs(lp - ref.numpy()).max() <= LP_TOL
    with pytest.raises(ValueError):
        m.forward(np.zeros((1, 2400, 1), dtype=np.int64), None)
    with pytest.raises(ValueError):
        m.forward(np.full((1, 2400, 1), 40000, dtype=np.int32), None)
    with pytest.raises(TypeError):
        m.forward([0] * 2400, None)


def test_long_stream_stays_finite_and_tracks_oracle(engines, weights, tb):
    """~24 s of audio through the state path: no drift / NaN; compare with the oracle on the last chunk."""
    C, B, n = 2400, 2, 80
    eng = engines(C)
    W = orc.to_torch(weights)
    pcm = tb.synth.telephony_pcm(B, C * n, seed=5)
    slots = eng.alloc_slots(B)
    try:
        lp, _ = _stream_engine(eng, slots, pcm, C)
        assert np.isfinite(lp).all()
        ref, _ = _stream_oracle(W, pcm, C)
        assert np.abs(lp[-5:] - ref[-5:]).max() <= LP_TOL
    finally:
        eng.release_slots(slots)
