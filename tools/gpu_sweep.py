"""BASELINE.json configs[2]: batch-size vs per-chunk latency sweep on one B200 (1024 resident streams).
For each batch size: device step time (CUDA events on the launching stream, inputs resident in HBM) and the end-to-end
per-step latency distribution through the C ABI with pinned host buffers (H2D of PCM + step + D2H of logprobs/tokens).
Usage: python tools/gpu_sweep.py [--chunk 2400] [--out gpurun_out/sweep.json] [B ...]"""
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
tb = importlib.import_module("t-one_b200")
FLOP = {2400: 1_287_738_880, 3200: 1_631_636_224}


def run(weights, B, chunk, steps, groups):
    eng = tb.Engine(weights, chunk_samples=chunk, max_slots=B * groups, max_batch=B)
    gs = [eng.alloc_slots(B) for _ in range(groups)]
    pcm = tb.synth.telephony_pcm(min(B, 64), chunk * 4, seed=1)
    pcm = np.tile(pcm, (B // min(B, 64) + 1, 1))[:B]
    d_pcm = torch.from_numpy(np.ascontiguousarray(pcm.reshape(B, 4, chunk).transpose(1, 0, 2))).cuda()
    d_slots = torch.from_numpy(np.stack(gs, 0)).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(10):
            eng.step_device(B, d_slots[i % groups].data_ptr(), d_pcm[i % 4].data_ptr(), 0, 0, st.cuda_stream)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            eng.step_device(B, d_slots[i % groups].data_ptr(), d_pcm[i % 4].data_ptr(), 0, 0, st.cuda_stream)
        e1.record(st)
        st.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    lat = []
    for i in range(10 + steps):
        eng.h_slots[:B] = gs[i % groups]
        eng.h_pcm[:B] = pcm[:, (i % 4) * chunk:(i % 4 + 1) * chunk]
        t0 = time.perf_counter()
        eng.step_pinned(B)
        if i >= 10:
            lat.append((time.perf_counter() - t0) * 1e3)
    audio = B * chunk / 8000.0
    r = {"streams": B, "chunk_samples": chunk, "steps": steps, "device_ms_per_step": dev_ms,
         "rtfx_device": audio / (dev_ms / 1e3), "tflops": B * FLOP[chunk] / (dev_ms / 1e3) / 1e12,
         "e2e_ms_p50": float(np.percentile(lat, 50)), "e2e_ms_p99": float(np.percentile(lat, 99)),
         "e2e_ms_max": float(np.max(lat)), "rtfx_e2e": audio / (float(np.mean(lat)) / 1e3),
         "launches_per_step": int(eng._get_info().launches_per_step)}
    eng.close()
    return r


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("batches", nargs="*", type=int, default=[1, 8, 16, 32, 64, 128, 256, 512, 1024])
    ap.add_argument("--chunk", type=int, default=2400)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    a = ap.parse_args()
    weights = tb.weights.init_weights(0)
    res = []
    for B in a.batches:
        groups = max(1, min(4, 1024 // B))
        r = run(weights, B, a.chunk, a.steps, groups)
        res.append(r)
        print(f"B={B:5d}: device {r['device_ms_per_step']*1e3:8.1f} us/step  RTFx {r['rtfx_device']:9.0f}  {r['tflops']:6.1f} TFLOP/s | "
              f"e2e p50 {r['e2e_ms_p50']:.3f} ms p99 {r['e2e_ms_p99']:.3f} ms  RTFx {r['rtfx_e2e']:9.0f}", flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump({"device": torch.cuda.get_device_name(0), "results": res}, f, indent=1)
