"""Error statistics of the engine against the fp32 oracle for one configuration (log-prob error by reference log-prob
range).  Usage: python tools/gpu_errstat.py B [chunk=2400] [chunks=3] [distinct=32] [engine kwargs ...]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tone_oracle as orc  # noqa: E402

tb = importlib.import_module("t-one_b200")


def main():
    B = int(sys.argv[1])
    kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
    C, n, D, seed = kw.pop("chunk", 2400), kw.pop("chunks", 3), kw.pop("distinct", 32), kw.pop("seed", 500 + B)
    w = tb.weights.init_weights(0)
    eng = tb.Engine(w, chunk_samples=C, max_slots=B, max_batch=B, **kw)
    W = orc.to_torch(w)
    D = min(D, B)
    distinct = tb.synth.telephony_pcm(D, C * n, seed=seed)
    idx = (np.arange(B) * 7) % D
    pcm = np.ascontiguousarray(distinct[idx])
    slots = eng.alloc_slots(B)
    st = orc.zero_state(D)
    outs, refs = [], []
    for i in range(n):
        lp, _ = eng.step(slots, pcm[:, i * C:(i + 1) * C])
        outs.append(lp.copy())
        r, st = orc.step(W, torch.from_numpy(distinct[:, i * C:(i + 1) * C].astype(np.int32)), st)
        refs.append(r.numpy())
    lp, ref = np.stack(outs), np.stack(refs)[:, idx]
    err = np.abs(lp - ref)
    print(f"B={B} C={C} {kw}: max {err.max():.4f} (at ref log-prob {ref.flat[err.argmax()]:.2f}), rms {np.sqrt((err ** 2).mean()):.4f}")
    for lo in (-2, -6, -10, -30):
        m = ref > lo
        print(f"   ref > {lo:3d}: n={int(m.sum()):8d}  max {err[m].max():.4f}  p99.99 {np.percentile(err[m], 99.99):.4f}  rms {np.sqrt((err[m] ** 2).mean()):.4f}")
    eng.close()


if __name__ == "__main__":
    main()
