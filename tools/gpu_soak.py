"""BASELINE.json configs[4]: offline long-form through the streaming state path - 256 streams x 400 ms chunks.
Runs `--steps` consecutive chunks per stream (9000 = 1 hour of audio per stream), checks every `--check-every` steps
that logprobs and the exported state stay finite and bounded, compares the first `--oracle-steps` steps of two streams
with the CPU oracle, and reports the sustained throughput.  Usage: python tools/gpu_soak.py [--steps 9000]"""
import argparse
import importlib
import json
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=256)
    ap.add_argument("--chunk", type=int, default=3200)
    ap.add_argument("--steps", type=int, default=9000)
    ap.add_argument("--check-every", type=int, default=500)
    ap.add_argument("--oracle-steps", type=int, default=60)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "soak.json"))
    a = ap.parse_args()
    B, C = a.streams, a.chunk
    weights = tb.weights.init_weights(0)
    eng = tb.Engine(weights, chunk_samples=C, max_slots=B, max_batch=B)
    slots = eng.alloc_slots(B)
    n_distinct = 64                                   # 64 distinct chunks per stream, cycled (25.6 s of audio)
    pcm = tb.synth.telephony_pcm(B, C * n_distinct, seed=77).reshape(B, n_distinct, C)
    # ---- parity on the first steps (two streams)
    W = orc.to_torch(weights)
    st = orc.zero_state(2)
    worst = 0.0
    for i in range(a.oracle_steps):
        lp, _ = eng.step(slots, pcm[:, i % n_distinct])
        ref, st = orc.step(W, torch.from_numpy(pcm[:2, i % n_distinct].astype(np.int32)), st)
        worst = max(worst, float(np.abs(lp[:2] - ref.numpy()).max()))
    # ---- the long run: device-resident PCM, state carried in the slots
    d_pcm = torch.from_numpy(np.ascontiguousarray(pcm.transpose(1, 0, 2))).cuda()
    d_slots = torch.from_numpy(slots).cuda()
    d_lp = torch.empty((B, eng.T, 35), dtype=torch.float32, device="cuda")
    stream = torch.cuda.Stream()
    checks = []
    t_gpu = 0.0
    done = a.oracle_steps
    with torch.cuda.stream(stream):
        while done < a.steps:
            n = min(a.check_every, a.steps - done)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for i in range(n):
                eng.step_device(B, d_slots.data_ptr(), d_pcm[(done + i) % n_distinct].data_ptr(), d_lp.data_ptr(), 0,
                                stream.cuda_stream)
            e1.record(stream)
            stream.synchronize()
            t_gpu += e0.elapsed_time(e1) / 1e3
            done += n
            lp = d_lp.cpu().numpy()
            state = eng.export_state(int(slots[0])).astype(np.float32)
            ok = bool(np.isfinite(lp).all() and np.isfinite(state).all())
            checks.append({"step": done, "finite": ok, "logprob_min": float(lp.min()), "state_absmax": float(np.abs(state).max()),
                           "prob_sum_err": float(np.abs(np.exp(lp).sum(-1) - 1).max())})
            print(checks[-1], flush=True)
            assert ok, "non-finite values"
    steps_timed = a.steps - a.oracle_steps
    audio_s = B * C / 8000.0 * steps_timed
    res = {"streams": B, "chunk_samples": C, "steps_per_stream": a.steps, "audio_hours_total": B * C / 8000.0 * a.steps / 3600,
           "max_abs_dlogprob_first_steps": worst, "oracle_steps": a.oracle_steps, "gpu_seconds": t_gpu,
           "rtfx": audio_s / t_gpu, "ms_per_step": 1e3 * t_gpu / steps_timed, "checks": checks}
    print(json.dumps({k: v for k, v in res.items() if k != "checks"}))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    eng.close()


if __name__ == "__main__":
    main()
