"""GPU parity tests (run with -m gpu on the B200 box): the CUDA step, called through the C ABI, against the CPU oracle
on the same seeded inputs and against the committed reference golden vectors.

Stated tolerance (floating point path, bf16 tensor-core operands with fp32 accumulation, fp32 norms / softmax):
  logprobs  vs the fp32 oracle / reference goldens, by reference log-prob (bf16 operand rounding acts on the logits, so
            the absolute error grows with |log-prob|): |d| <= 0.06 above -6 (p > 0.25 %: every class greedy or beam
            decoding can act on), <= 0.08 above -10, <= 0.10 everywhere, and rms <= 0.02.  Measured on B200 (1024-stream
            case, 1.07 M log-probs; identical statistics on every kernel path): rms 0.014, max 0.043 above -6, 0.052 - 0.062
            above -10 (a 4.4-sigma tail whose exact value moves with the summation order of the path), 0.060 overall;
            0.053 worst over a one-hour stream.  The bf16 operand-rounding spread of the reference algorithm itself is
            0.05 (tests/test_oracle.py::test_bf16_emulation_within_stated_tolerance).  <= 0.05 vs the bf16-emulating oracle
  state     |d| <= 0.06 (fp16 wire format, values up to ~4.5; measured 0.02), mhsa_len exact
  tokens    identical wherever the oracle's top-2 logprob margin exceeds the logprob tolerance
  stages    residual stream after pre-encode and after every Conformer layer: per-stage bounds in STAGE_TOL
"""
import json
import os
import numpy as np
import pytest
import torch

import tone_oracle as orc

pytestmark = pytest.mark.gpu

LP_TOL, LP_TOL_EMU, ST_TOL = 0.06, 0.05, 0.06
LP_TIERS = ((-6.0, 0.06), (-10.0, 0.08), (-np.inf, 0.10))      # (reference log-prob above, max |d|)
LP_RMS = 0.02


def _lp_close(lp, ref):
    """The stated log-prob tolerance: tiered by reference log-prob, plus an rms bound (no rms bound on tiny samples)."""
    err = np.abs(lp - ref)
    ok = all((err[ref > lo] <= tol).all() for lo, tol in LP_TIERS)
    return bool(ok and (err.size < 1000 or np.sqrt((err ** 2).mean()) <= LP_RMS))


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
    got = eng.export_states(slots).astype(np.float32)
    assert np.array_equal(got[:, len_off], flat_ref[:, len_off])
    assert np.abs(got - flat_ref).max() <= ST_TOL


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
        assert _lp_close(lp, g["logprobs"])
        _check_tokens(tk, g["logprobs"])
        got = eng.export_state(int(slots[0])).astype(np.float32)          # the golden holds the state of stream 0
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
        assert _lp_close(lp, ref)
        if emu is not None:
            assert np.abs(lp - emu).max() <= LP_TOL_EMU
        _check_tokens(tk, ref)
        assert (tk == lp.argmax(-1)).all()                 # fused argmax == numpy argmax of our own logprobs
        np.testing.assert_allclose(np.exp(lp).sum(-1), 1.0, atol=1e-4)
        _check_state(eng, slots, st)
    finally:
        eng.release_slots(slots)


# ---- large-batch paths: 128-wide tiles (>= 2048 rows per lane), two concurrent lanes (from 628 streams), and the
# persistent GEMM kernel.  Streams are independent, so B streams that replay a handful of distinct signals must all
# reproduce the oracle's answer for their signal (size-independent property; the oracle runs the distinct signals only).
@pytest.mark.parametrize("C,B,persist,mode,ff", [(2400, 600, 0, 0, (0, 0)), (2400, 600, -1, 1, (1, 0)), (2400, 600, 1, 1, (1, 0)),
                                                  (2400, 600, 1, 2, (1, 0)), (2400, 600, 1, 3, (1, 0)), (3200, 420, 1, 3, (1, 0)),
                                                  (2400, 230, 1, 3, (1, 0)), (2400, 333, 1, 2, (1, 0)),
                                                  (2400, 230, 0, 0, (2, 1)), (2400, 230, 0, 0, (3, 1)), (3200, 333, 0, 0, (3, 1)),
                                                  (2400, 600, 0, 0, (2, 0)), (2400, 37, 0, 0, (3, 1))])
def test_large_batch_paths_match_oracle(weights, tb, C, B, persist, mode, ff):
    # persist: smallest GEMM (in 128 x 128 tiles) that takes the persistent kernel (tone_config.persist_min_tiles:
    # 0 = default, -1 = never); mode: its tile form (tone_config.persist_mode: 0 = default, 1 = 128 wide, 2 = 256 wide,
    # 3 = 256 wide on CTA pairs); ff = (tone_config.fused_ff, fused_ff_min_rows): fused feed-forward kernel
    # 1 = off, 2 = one CTA per 128 rows, 3 = CTA pairs (default from 2048 rows per lane; min_rows 1 forces it for every
    # layer, which also covers odd tile counts and ragged last tiles)
    n, D = 3, 6
    eng = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B, persist_min_tiles=persist, persist_mode=mode,
                    fused_ff=ff[0], fused_ff_min_rows=ff[1])
    try:
        W = orc.to_torch(weights)
        distinct = tb.synth.telephony_pcm(D, C * n, seed=77)
        idx = np.arange(B) % D
        pcm = np.ascontiguousarray(distinct[idx])
        slots = eng.alloc_slots(B)
        lp, tk = _stream_engine(eng, slots, pcm, C)
        ref, st = _stream_oracle(W, distinct, C)
        assert np.isfinite(lp).all()
        assert _lp_close(lp, ref[:, idx])
        assert (tk == lp.argmax(-1)).all()
        _check_tokens(tk, ref[:, idx])
        flat_ref = orc.pack_state(st).astype(np.float32)
        probe = np.array([0, 1, B // 2, B - 2, B - 1])
        got = eng.export_states(slots[probe]).astype(np.float32)
        assert np.abs(got - flat_ref[idx[probe]]).max() <= ST_TOL
    finally:
        eng.close()


# ---- the lane rule: a step is cut into two concurrent lanes when one lane's 384-column GEMMs exceed one 128 x 128 tile
# per SM (ceil(B T / 128) * 3 > 148): 576 streams x 10 frames = 135 tiles -> one lane, 640 streams = 150 tiles -> two lanes
# of 320 (3200-row lanes: a size no other test runs).  Both sides of the boundary must reproduce the oracle.
def test_lane_rule_follows_the_tile_count_and_both_sides_match_oracle(weights, tb):
    C, n, D = 2400, 3, 6
    W = orc.to_torch(weights)
    distinct = tb.synth.telephony_pcm(D, C * n, seed=79)
    ref, st = _stream_oracle(W, distinct, C)
    launches = {}
    for B in (576, 640):
        eng = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B)
        try:
            idx = (np.arange(B) * 7) % D
            pcm = np.ascontiguousarray(distinct[idx])
            slots = eng.alloc_slots(B)
            lp, tk = _stream_engine(eng, slots, pcm, C)
            launches[B] = eng._get_info().launches_per_step
            assert np.isfinite(lp).all()
            assert _lp_close(lp, ref[:, idx])
            assert (tk == lp.argmax(-1)).all()
            _check_tokens(tk, ref[:, idx])
            flat_ref = orc.pack_state(st).astype(np.float32)
            probe = np.array([0, 1, B // 2 - 1, B // 2, B - 1])
            got = eng.export_states(slots[probe]).astype(np.float32)
            assert np.abs(got - flat_ref[idx[probe]]).max() <= ST_TOL
        finally:
            eng.close()
    assert launches[576] < 200 and launches[640] >= 2 * launches[576] - 8, launches   # 184 kernels in one lane, 2 x 191 in two


# ---- large-batch kernel forms added in round 2b, forced on at sizes the oracle check can afford: the fused attention
# block (V projection + P.V + out projection + residual per tile of whole streams; ragged last tile, both frame rates,
# 400 ms chunks: 9 / 18 streams per tile), feed-forward 1 straight into the residual stream with norm_self_att as a row
# scale inside the projections (needs >= 2048 rows per lane), and the pipelined depthwise-conv kernel at every size.
@pytest.mark.parametrize("C,B,att,lazy,dw,ap,rg", [(2400, 600, 1, 1, 1, 1, -1), (3200, 333, 1, 1, 1, 1, -1), (2400, 230, 1, -1, 1, -1, -1),
                                                   (2400, 37, 1, 1, 1, 1, -1), (2400, 600, -1, 1, -1, -1, -1), (3200, 61, -1, -1, 1, 1, -1),
                                                   (2400, 5, -1, -1, 1, 1, -1), (2400, 600, -1, 1, 1, 1, 1), (3200, 333, -1, -1, 1, 1, 1),
                                                   (2400, 37, -1, -1, -1, -1, 1), (2400, 230, 1, 1, 1, 1, 1)])
def test_large_batch_fused_blocks_match_oracle(weights, tb, C, B, att, lazy, dw, ap, rg):
    n, D = 3, 6
    # ap: tone_config.att_pipe_min_batch (recompute attention layers as the pipelined persistent kernel); rg:
    # tone_config.rowgemm_min_rows (N = 384 projections as the row-owner CTA-pair kernel; odd tile counts, ragged tiles)
    eng = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B, att_block_min_rows=att, lazy_norm_min_rows=lazy,
                    dw_pipe_min_batch=dw, att_pipe_min_batch=ap, rowgemm_min_rows=rg)
    try:
        W = orc.to_torch(weights)
        distinct = tb.synth.telephony_pcm(D, C * n, seed=78)
        idx = (np.arange(B) * 5) % D
        pcm = np.ascontiguousarray(distinct[idx])
        slots = eng.alloc_slots(B)
        lp, tk = _stream_engine(eng, slots, pcm, C)
        ref, st = _stream_oracle(W, distinct, C)
        assert np.isfinite(lp).all()
        assert _lp_close(lp, ref[:, idx])
        assert (tk == lp.argmax(-1)).all()
        _check_tokens(tk, ref[:, idx])
        flat_ref = orc.pack_state(st).astype(np.float32)
        probe = np.array([0, 1, B // 2, B - 2, B - 1])
        got = eng.export_states(slots[probe]).astype(np.float32)
        assert np.abs(got - flat_ref[idx[probe]]).max() <= ST_TOL
    finally:
        eng.close()


# ---- BASELINE configs[2] / configs[3] shapes: 1024 streams on one GPU (two lanes of 5120 rows: persistent gated GEMMs,
# 128-wide tiles everywhere) and 4096 streams (the per-GPU share of 8192 streams on 2 GPUs; 20480 rows per lane).
# 32 / 40 distinct signals replayed across the batch; every stream must reproduce the oracle's answer for its signal,
# and the carried state of every stream is checked, not a sample.
@pytest.mark.parametrize("C,B,D,n", [(2400, 1024, 32, 3), (3200, 1024, 32, 2), (2400, 4096, 40, 2)])
def test_baseline_throughput_shapes_match_oracle(weights, tb, C, B, D, n):
    eng = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B)
    try:
        W = orc.to_torch(weights)
        distinct = tb.synth.telephony_pcm(D, C * n, seed=500 + B)
        idx = (np.arange(B) * 7) % D                        # neighbours in the batch carry different signals
        pcm = np.ascontiguousarray(distinct[idx])
        slots = eng.alloc_slots(B)
        lp, tk = _stream_engine(eng, slots, pcm, C)
        ref, st = _stream_oracle(W, distinct, C)
        assert np.isfinite(lp).all()
        assert _lp_close(lp, ref[:, idx])
        assert (tk == lp.argmax(-1)).all()
        _check_tokens(tk, ref[:, idx])
        flat_ref = orc.pack_state(st).astype(np.float32)
        len_off = 80 + 2 * 30 * 384 + 16 * 384 * 30
        worst = 0.0
        for i0 in range(0, B, 256):
            got = eng.export_states(slots[i0:i0 + 256]).astype(np.float32)
            want = flat_ref[idx[i0:i0 + 256]]
            assert np.array_equal(got[:, len_off], want[:, len_off])
            worst = max(worst, float(np.abs(got - want).max()))
        assert worst <= ST_TOL
        # streams that carry the same signal agree bit-exactly when they sit in the same lane (same kernel selection)
        same = np.nonzero(idx[: B // 2] == idx[0])[0]
        assert all(np.array_equal(lp[:, same[0]], lp[:, j]) for j in same[1:])
    finally:
        eng.close()


# ---- per-stage parity: the residual stream after pre-encode and after every Conformer layer (tone_step_debug taps)
# against the oracle's taps.  Bounds per stage are 2x what was measured on B200 (gpurun_out/stage_errors.json of the
# same test, copied to profiles/r02_stage_errors.json): a regression of one stage cannot hide below the end-to-end
# tolerance any more.
STAGE_NAMES = ["pre_encode"] + [f"layer{l}" for l in range(16)]
# measured max |d| on B200 (profiles/r02_stage_errors_{2400,3200}.json), worst of the two chunk lengths
_MEASURED = {"fp32": [0.0158, 0.0160, 0.0162, 0.0181, 0.0178, 0.0190, 0.0185, 0.0082, 0.0187, 0.0200, 0.0215, 0.0234, 0.0226,
                      0.0304, 0.0349, 0.0309, 0.0287],
             "bf16emu": [0.0017, 0.0033, 0.0042, 0.0053, 0.0057, 0.0069, 0.0091, 0.0039, 0.0112, 0.0121, 0.0134, 0.0155, 0.0141,
                         0.0206, 0.0198, 0.0206, 0.0179]}
STAGE_TOL = {mode: {name: 2.0 * v for name, v in zip(STAGE_NAMES, vals)} for mode, vals in _MEASURED.items()}


@pytest.mark.parametrize("C", [2400, 3200])
def test_per_stage_residual_stream_matches_oracle(weights, tb, C):
    B, n = 5, 4
    eng = tb.Engine(weights, chunk_samples=C, max_slots=8, max_batch=8, use_graph=False)
    try:
        W = orc.to_torch(weights)
        pcm = tb.synth.telephony_pcm(B, C * n, seed=1234)
        slots = eng.alloc_slots(B)
        st32, stemu = orc.zero_state(B), orc.zero_state(B)
        worst = {"fp32": dict.fromkeys(STAGE_NAMES, 0.0), "bf16emu": dict.fromkeys(STAGE_NAMES, 0.0)}
        scale = dict.fromkeys(STAGE_NAMES, 0.0)
        for i in range(n):
            chunk = pcm[:, i * C:(i + 1) * C]
            t32, temu = {}, {}
            _, st32 = orc.step(W, torch.from_numpy(chunk), st32, taps=t32)
            _, stemu = orc.step(W, torch.from_numpy(chunk), stemu, quant=orc.bf16_round, taps=temu)
            _, _, taps = eng.step_debug(slots, chunk)
            for k, name in enumerate(STAGE_NAMES):
                for mode, ref in (("fp32", t32), ("bf16emu", temu)):
                    r = ref[name].numpy().reshape(-1, 384)
                    d = float(np.abs(taps[k][: r.shape[0]] - r).max())
                    worst[mode][name] = max(worst[mode][name], d)
                scale[name] = max(scale[name], float(np.abs(t32[name].numpy()).max()))
        out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
        if os.path.isdir(out):
            with open(os.path.join(out, f"stage_errors_{C}.json"), "w") as f:
                json.dump({"chunk_samples": C, "streams": B, "chunks": n, "max_abs_err": worst, "ref_abs_max": scale}, f, indent=1)
        for mode in worst:
            for name in STAGE_NAMES:
                assert worst[mode][name] <= STAGE_TOL[mode][name], (mode, name, worst[mode][name])
    finally:
        eng.close()


def test_bad_batches_are_rejected_by_the_c_abi(engines, tb):
    """Duplicate / unallocated / out-of-range slot ids and out-of-range samples are errors, not silent corruption
    (tone/onnx_wrapper.py:100-121 for the sample range)."""
    eng = engines(2400)
    slots = eng.alloc_slots(3)
    try:
        pcm = tb.synth.telephony_pcm(3, 2400, seed=1)
        with pytest.raises(ValueError, match="twice"):
            eng.step(np.array([slots[0], slots[1], slots[0]], dtype=np.int32), pcm)
        with pytest.raises(ValueError, match="out of range"):
            eng.step(np.array([slots[0], 10 ** 6, slots[1]], dtype=np.int32), pcm)
        free = [s for s in range(80) if s not in set(slots.tolist())][0]
        with pytest.raises(tb.model.ToneError, match="not allocated"):
            eng.step(np.array([slots[0], free, slots[1]], dtype=np.int32), pcm)
        with pytest.raises(ValueError, match="twice"):
            eng.step_device(np.array([slots[2], slots[2]], dtype=np.int32), 0)
        bad = pcm.copy()
        bad[1, 100] = 40000
        before = eng.export_states(slots)
        with pytest.raises(ValueError, match=r"range \[-32768; 32767\]"):
            eng.step(slots, bad)
        assert np.array_equal(before, eng.export_states(slots))      # a rejected step leaves the state untouched
        eng.step(slots, pcm)                                          # and the engine keeps working
    finally:
        eng.release_slots(slots)


def test_int16_wire_format_and_pipelined_tickets_are_bit_identical(engines, tb):
    """tone_submit / tone_wait with int16 PCM and two tickets in flight == the synchronous int32 call."""
    C, B, n = 2400, 9, 6
    eng = engines(C)
    pcm = tb.synth.telephony_pcm(B, C * n, seed=61)
    s1, s2 = eng.alloc_slots(B), eng.alloc_slots(B)
    try:
        want, _ = _stream_engine(eng, s1, pcm, C)
        got, pending = [], None
        for i in range(n):
            stg_slots, stg_pcm, _ = eng.next_staging(B)             # zero-copy staging of the next set
            stg_slots[:] = s2
            stg_pcm[:] = pcm[:, i * C:(i + 1) * C].astype(np.int16)
            t = eng.submit(stg_slots, stg_pcm, tb.model.OUT_LOGPROBS | tb.model.OUT_TOKENS | tb.model.OUT_SIL)
            if pending is not None:
                got.append(eng.wait(pending))
            pending = t
        got.append(eng.wait(pending))
        for i in range(n):
            assert np.array_equal(got[i]["logprobs"], want[i])
            assert np.array_equal(got[i]["tokens"], want[i].argmax(-1))
            assert np.array_equal(got[i]["sil"], want[i][:, :, 33:35])
        with pytest.raises(tb.model.ToneError, match="not in flight"):
            eng.wait(pending)
        t1 = eng.submit(s1, pcm[:, :C])
        t2 = eng.submit(s2, pcm[:, :C])
        with pytest.raises(tb.model.ToneError, match="not been waited"):
            eng.submit(s1, pcm[:, :C])                                # both staging sets are busy
        eng.wait(t1)
        eng.wait(t2)
    finally:
        eng.release_slots(np.concatenate([s1, s2]))


def test_caller_stream_is_ordered_with_engine_streams(engines, tb):
    """a step launched on the caller's stream must see the staged inputs, and fetch must see its outputs,
    without the caller synchronising anything."""
    C, B = 2400, 8
    eng = engines(C)
    pcm = tb.synth.telephony_pcm(B, C * 4, seed=88)
    s1, s2 = eng.alloc_slots(B), eng.alloc_slots(B)
    side = torch.cuda.Stream()
    try:
        want, _ = _stream_engine(eng, s1, pcm, C)
        for i in range(4):
            eng.stage(s2, pcm[:, i * C:(i + 1) * C])
            eng.step_staged(B, side.cuda_stream)
            lp, tk = eng.fetch(B)
            assert np.array_equal(lp, want[i])
        # device-pointer form on the caller's stream, int16 and int32 device PCM
        eng.reset_slots(s2)
        for i in range(4):
            x = torch.from_numpy(pcm[:, i * C:(i + 1) * C].astype(np.int16 if i % 2 else np.int32)).cuda()
            d_lp = torch.empty((B, eng.T, 35), dtype=torch.float32, device="cuda")
            with torch.cuda.stream(side):
                eng.step_device(s2, x.data_ptr(), tb.model.PCM_I16 if i % 2 else tb.model.PCM_I32, d_lp.data_ptr(), 0,
                                side.cuda_stream)
            side.synchronize()
            assert np.array_equal(d_lp.cpu().numpy(), want[i])
    finally:
        eng.release_slots(np.concatenate([s1, s2]))


def test_triton_three_tensor_state_through_the_c_abi(engines, tb):
    """tone_export/import_states_triton == the Python restatement of tone/scripts/export.py:293-376 on the flat state."""
    C, B = 2400, 3
    eng = engines(C)
    pcm = tb.synth.telephony_pcm(B, C * 3, seed=91)
    s1, s2 = eng.alloc_slots(B), eng.alloc_slots(B)
    try:
        _stream_engine(eng, s1, pcm[:, :2 * C], C)
        flat = eng.export_states(s1)
        tm, ch, ln = eng.export_states_triton(s1)
        tm2, ch2, ln2 = tb.state_formats.flat_to_triton(flat)
        assert np.array_equal(tm, tm2) and np.array_equal(ch, ch2) and np.array_equal(ln, ln2)
        eng.import_states_triton(s2, tm, ch, ln)
        assert np.array_equal(eng.export_states(s2), flat)
        la, _ = eng.step(s1, pcm[:, 2 * C:])
        lb, _ = eng.step(s2, pcm[:, 2 * C:])
        assert np.abs(la - lb).max() < 3e-2
    finally:
        eng.release_slots(np.concatenate([s1, s2]))


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
            assert _lp_close(lp, ref.numpy())
            assert _lp_close(lp, g["logprobs"][i])
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
        assert _lp_close(lp, ref.numpy())
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
        assert _lp_close(lp[-5:], ref[-5:])
    finally:
        eng.release_slots(slots)


def test_results_do_not_depend_on_pool_size_or_slot_position(weights, tb):
    """Stand-in for a memcheck run (compute-sanitizer is closed on this pool): the same streams stepped in an engine
    sized exactly for them and in a much larger engine, at the far end of its slot pool, give bit-identical log-probs and
    state.  An out-of-bounds read that lands in a neighbouring buffer, or any dependence on the pool layout, shows up
    as a difference."""
    C, B, n = 2400, 11, 4
    pcm = tb.synth.telephony_pcm(B, C * n, seed=123)
    small = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B)
    big = tb.Engine(weights, chunk_samples=C, max_slots=8 * B + 5, max_batch=B)
    try:
        sa = small.alloc_slots(B)
        junk = big.alloc_slots(7 * B + 5)
        sb = big.alloc_slots(B)[::-1].copy()
        la, _ = _stream_engine(small, sa, pcm, C)
        lb, _ = _stream_engine(big, sb, pcm, C)
        assert np.array_equal(la, lb)
        assert np.array_equal(small.export_states(sa), big.export_states(sb))
        assert not np.abs(big.export_states(junk[:4]).astype(np.float32)).any()      # untouched neighbours stay all-zero
    finally:
        small.close()
        big.close()
