"""BASELINE.json configs[2]: batch-size vs per-chunk latency sweep on one B200.
For each batch size: device step time (CUDA events on the launching stream, int16 PCM resident in HBM), the synchronous
per-chunk latency distribution through the C ABI (tone_step: int32 host PCM in, log-probs + tokens out) and the pipelined
end-to-end throughput (tone_submit / tone_wait, two tickets in flight).
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


def run(weights, B, chunk, steps):
    M = tb.model
    G = 2 if B >= 512 else 8
    eng = tb.Engine(weights, chunk_samples=chunk, max_slots=B * G, max_batch=B)
    gs = [eng.alloc_slots(B) for _ in range(G)]
    pcm = tb.synth.telephony_pcm(min(B, 64), chunk * 4, seed=1).reshape(-1, 4, chunk)
    pcm = np.ascontiguousarray(np.tile(pcm, (B // pcm.shape[0] + 1, 1, 1))[:B].transpose(1, 0, 2))
    pcm16 = pcm.astype(np.int16)
    d_pcm = torch.from_numpy(pcm16).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(10):
            eng.step_device(gs[i % G], d_pcm[i % 4].data_ptr(), M.PCM_I16, 0, 0, st.cuda_stream)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            eng.step_device(gs[i % G], d_pcm[i % 4].data_ptr(), M.PCM_I16, 0, 0, st.cuda_stream)
        e1.record(st)
        st.synchronize()
    dev_ms = e0.elapsed_time(e1) / steps
    lat = []
    for i in range(10 + steps):
        t0 = time.perf_counter()
        eng.step(gs[i % G], pcm[i % 4])
        if i >= 10:
            lat.append(time.perf_counter() - t0)
    pend, t0 = None, None
    for i in range(10 + steps):
        if i == 10:
            if pend is not None:
                eng.wait(pend)
                pend = None
            t0 = time.perf_counter()
        s, p, l = eng.next_staging(B)
        s[:] = gs[i % G]
        p[:] = pcm16[i % 4]
        t = eng.submit(s, p, M.OUT_LOGPROBS | M.OUT_TOKENS)
        if pend is not None:
            eng.wait(pend)
        pend = t
    eng.wait(pend)
    pipe_s = time.perf_counter() - t0
    audio = B * chunk / 8000.0
    r = {"streams": B, "chunk_samples": chunk, "ms_per_step_device": dev_ms, "rtfx_device": audio / dev_ms * 1e3,
         "rtfx_e2e_pipelined": audio * steps / pipe_s, "sync_latency_ms_p50": float(np.percentile(lat, 50) * 1e3),
         "sync_latency_ms_p99": float(np.percentile(lat, 99) * 1e3), "rtfx_sync": audio / float(np.mean(lat)),
         "tflops": FLOP[chunk] * B / dev_ms / 1e9, "frac_of_1412.7": FLOP[chunk] * B / dev_ms / 1e9 / 1412.7,
         "launches_per_step": int(eng._get_info().launches_per_step)}
    print(json.dumps(r), flush=True)
    eng.close()
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunk", type=int, default=2400)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    ap.add_argument("B", nargs="*", type=int)
    a = ap.parse_args()
    weights = tb.weights.init_weights(0)
    res = [run(weights, B, a.chunk, a.steps) for B in (a.B or [1, 16, 64, 128, 256, 512, 1024])]
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
