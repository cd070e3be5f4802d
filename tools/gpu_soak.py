"""BASELINE.json configs[4]: offline long-form through the streaming state path - 256 streams x 400 ms chunks.
Runs `--steps` consecutive chunks per stream (9000 = 1 hour of audio per stream).  Every `--check-every` steps
("check no NaN / drift vs oracle on a sampled stream every 500 steps"):
  * log-probs and the exported state must be finite, probabilities must sum to one;
  * ORACLE CHECK: the carried state of `--sample` streams is exported (flat fp16 wire format), the oracle is re-seeded from
    exactly that state (unpack_state) and steps the next chunk; the engine's log-probs of that chunk must agree within
    the stated tolerance (0.06 above a reference log-prob of -6, 0.08 above -10, 0.10 everywhere).  Errors cannot hide behind a drifting state: every check restarts the oracle from the
    engine's own state at that point of the hour.
The first `--oracle-steps` steps are also compared chunk by chunk from the zero state.
Usage: python tools/gpu_soak.py [--steps 9000] [--out gpurun_out/soak.json]"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import tone_oracle as orc  # noqa: E402

tb = importlib.import_module("t-one_b200")
LP_TOL = 0.06


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=3200)
    ap.add_argument("--steps", type=int, default=9000)
    ap.add_argument("--check-every", type=int, default=500)
    ap.add_argument("--oracle-steps", type=int, default=20)
    ap.add_argument("--sample", type=int, default=4, help="streams re-checked against the oracle at every check")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "soak.json"))
    a = ap.parse_args()
    B, C = a.streams, a.chunk
    weights = tb.weights.init_weights(0)
    eng = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B)
    slots = eng.alloc_slots(B)
    n_distinct = 64                                   # 64 distinct chunks per stream, cycled (25.6 s of audio)
    pcm = tb.synth.telephony_pcm(B, C * n_distinct, seed=77).reshape(B, n_distinct, C)
    W = orc.to_torch(weights)
    # ---- parity from the zero state (two streams)
    st = orc.zero_state(2)
    worst0 = 0.0
    for i in range(a.oracle_steps):
        lp, _ = eng.step(slots, pcm[:, i % n_distinct])
        ref, st = orc.step(W, torch.from_numpy(pcm[:2, i % n_distinct].astype(np.int32)), st)
        worst0 = max(worst0, float(np.abs(lp[:2] - ref.numpy()).max()))
    # ---- the long run: device-resident int16 PCM, state carried in the slots
    d_pcm = torch.from_numpy(np.ascontiguousarray(pcm.transpose(1, 0, 2)).astype(np.int16)).cuda()
    d_lp = torch.empty((B, eng.T, 35), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    checks, t_gpu, done, worst = [], 0.0, a.oracle_steps, 0.0
    rng = np.random.default_rng(5)
    while done < a.steps:
        n = min(a.check_every, a.steps - done) - 1
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(n):
                eng.step_device(slots, d_pcm[(done + i) % n_distinct].data_ptr(), tb.model.PCM_I16, d_lp.data_ptr(), 0,
                                stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
        t_gpu += e0.elapsed_time(e1) / 1e3
        done += n
        # oracle check on sampled streams: restart the oracle from the engine's own carried state, step one chunk on both
        pick = np.sort(rng.choice(B, size=a.sample, replace=False))
        state = eng.export_states(slots[pick])
        chunk = pcm[:, done % n_distinct]
        lp, tk = eng.step(slots, chunk)
        done += 1
        ref, _ = orc.step(W, torch.from_numpy(chunk[pick].astype(np.int32)), orc.unpack_state(state))
        d = np.abs(lp[pick] - ref.numpy())
        err = float(d[ref.numpy() > -6.0].max())             # stated tolerance: 0.06 above a reference log-prob of -6,
        err_mid = float(d[ref.numpy() > -10.0].max())         # 0.08 above -10,
        err_tail = float(d.max())                            # 0.10 everywhere
        worst = max(worst, err_tail)
        top2 = np.sort(ref.numpy(), axis=-1)[..., -2:]
        decided = (top2[..., 1] - top2[..., 0]) > LP_TOL
        tok_ok = bool((tk[pick][decided] == ref.numpy().argmax(-1)[decided]).all())
        st32 = state.astype(np.float32)
        ok = bool(np.isfinite(lp).all() and np.isfinite(st32).all())
        checks.append({"step": done, "finite": ok, "streams_checked": pick.tolist(), "max_abs_dlogprob_vs_oracle": err, "max_abs_dlogprob_above_m10": err_mid, "max_abs_dlogprob_tail": err_tail,
                       "tokens_equal_above_margin": tok_ok, "logprob_min": float(lp.min()),
                       "state_absmax": float(np.abs(st32).max()), "prob_sum_err": float(np.abs(np.exp(lp).sum(-1) - 1).max())})
        print(checks[-1], flush=True)
        assert ok, "non-finite values"
        assert err <= LP_TOL and err_mid <= 0.08 and err_tail <= 0.10 and tok_ok, f"oracle check failed at step {done}: {err}"
    steps_timed = sum(min(a.check_every, a.steps - s) - 1 for s in range(a.oracle_steps, a.steps, a.check_every))
    audio_s = B * C / 8000.0 * steps_timed
    res = {"streams": B, "chunk_samples": C, "steps_per_stream": a.steps, "audio_hours_total": B * C / 8000.0 * a.steps / 3600,
           "max_abs_dlogprob_first_steps": worst0, "oracle_steps_from_zero_state": a.oracle_steps,
           "oracle_checks": len(checks), "max_abs_dlogprob_over_all_checks": worst, "tolerance": LP_TOL,
           "gpu_seconds": t_gpu, "rtfx": audio_s / t_gpu, "ms_per_step": 1e3 * t_gpu / steps_timed, "checks": checks}
    print(json.dumps({k: v for k, v in res.items() if k != "checks"}))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    eng.close()


if __name__ == "__main__":
    main()
