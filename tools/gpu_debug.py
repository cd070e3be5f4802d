"""GPU bring-up diagnostics (run on the B200 box): GEMM self-test, then a per-layer comparison of the CUDA step
against the oracle for both GEMM implementations.  Prints a compact report; writes gpurun_out/debug_report.txt."""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tone_oracle as orc  # noqa: E402

tb = importlib.import_module("t-one_b200")
out_lines = []


def P(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    out_lines.append(s)


def gemm_selftest(eng):
    rng = np.random.default_rng(0)
    ok = True
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 128, 128, 128), (200, 384, 384, 64), (640, 384, 1536, 32),
                          (333, 3072, 384, 128)]:
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K)
        Ab = torch.from_numpy(A).bfloat16().float().numpy()
        Wb = torch.from_numpy(W).bfloat16().float().numpy()
        ref = Ab @ Wb.T
        t0 = time.time()
        got = eng.selftest_gemm(A, W, bn)
        err = np.abs(got - ref).max()
        P(f"gemm selftest M={M} N={N} K={K} BN={bn}: max err {err:.3e} ({time.time()-t0:.2f}s)")
        if not err < 1e-2:
            ok = False
            bad = np.argwhere(np.abs(got - ref) > 1e-2)
            P("   first bad idx", bad[:5].tolist(), "got", got[tuple(bad[0])], "ref", ref[tuple(bad[0])])
            P("   row err profile (first 16 rows)", np.abs(got - ref).max(1)[:16].round(3).tolist())
            P("   col err profile (first 16 cols)", np.abs(got - ref).max(0)[:16].round(3).tolist())
    return ok


def compare_step(weights, impl, C=2400, B=3, n_chunks=4, use_debug=True):
    W = orc.to_torch(weights)
    eng = tb.Engine(weights, chunk_samples=C, max_slots=8, max_batch=8, gemm_impl=impl, use_graph=False)
    pcm = tb.synth.telephony_pcm(B, C * n_chunks, seed=1234)
    slots = eng.alloc_slots(B)
    st = orc.zero_state(B)
    T = eng.T
    worst = 0.0
    for i in range(n_chunks):
        chunk = pcm[:, i * C:(i + 1) * C]
        taps_ref = {}
        lp_ref, st = orc.step(W, torch.from_numpy(chunk), st, quant=orc.bf16_round, taps=taps_ref)
        lp, tk, taps = eng.step_debug(slots, chunk)
        d = np.abs(lp - lp_ref.numpy()).max()
        worst = max(worst, d)
        agree = (tk == lp_ref.numpy().argmax(-1)).mean()
        P(f"[impl={impl} C={C}] chunk {i}: max|dlogprob| {d:.4f}  token agree {agree:.3f}  nan={np.isnan(lp).any()}")
        if d > 0.08 or np.isnan(lp).any():
            names = ["pre_encode"] + [f"layer{l}" for l in range(16)]
            for idx, nm in enumerate(names):
                ref = taps_ref[nm].numpy().reshape(-1, 384)
                got = taps[idx][: ref.shape[0]]
                P(f"     tap {nm:11s} max|d| {np.abs(got-ref).max():.4f}  ref absmax {np.abs(ref).max():.3f} nan={np.isnan(got).any()}")
    # carried state
    flat_ref = orc.pack_state(st).astype(np.float32)
    for b in range(B):
        got = eng.export_state(int(slots[b])).astype(np.float32)
        o = 0
        for k, shp in zip(orc.STATE_KEYS, orc.STATE_SHAPES):
            n = int(np.prod(shp))
            dd = np.abs(got[o:o + n] - flat_ref[b, o:o + n]).max()
            if b == 0 or dd > 0.1:
                P(f"     state[{b}] {k:9s} max|d| {dd:.4f}")
            o += n
    eng.close()
    return worst


def main():
    weights = tb.weights.init_weights(0)
    P("device", torch.cuda.get_device_name(0))
    eng = tb.Engine(weights, max_slots=4, max_batch=4, use_graph=False)
    P("weights MB", eng.info.weight_bytes / 1e6, "state KB/slot", eng.info.state_bytes_per_slot / 1e3)
    ok = gemm_selftest(eng)
    eng.close()
    P("GEMM selftest", "OK" if ok else "FAILED")
    which = sys.argv[1:] or ["1", "0"]
    for impl in which:
        for C in (2400, 3200):
            try:
                w = compare_step(weights, int(impl), C=C)
                P(f"== impl {impl} C={C}: worst max|dlogprob| {w:.4f}")
            except Exception as ex:  # keep going: the other implementation still tells us something
                P(f"== impl {impl} C={C}: EXCEPTION {type(ex).__name__}: {ex}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "debug_report.txt"), "w") as f:
        f.write("\n".join(out_lines) + "\n")


if __name__ == "__main__":
    main()
