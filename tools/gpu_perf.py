"""Device-time exploration on the B200 box: step time for batch sizes / PDL / graph settings (CUDA events, own stream).
Usage: python tools/gpu_perf.py [B ...]   Env: TONE_PDL=0|1"""
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")


def run(B, chunk=2400, steps=100, warm=10, use_graph=True, groups=4):
    eng = tb.Engine(weights, chunk_samples=chunk, max_slots=B * groups, max_batch=B, use_graph=use_graph)
    gs = [eng.alloc_slots(B) for _ in range(groups)]
    pcm = tb.synth.telephony_pcm(min(B, 64), chunk * 4, seed=1)
    pcm = np.tile(pcm, (B // min(B, 64) + 1, 1))[:B]
    d_pcm = torch.from_numpy(np.ascontiguousarray(pcm.reshape(B, 4, chunk).transpose(1, 0, 2))).cuda()
    d_slots = torch.from_numpy(np.stack(gs, 0)).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(warm):
            eng.step_device(B, d_slots[i % groups].data_ptr(), d_pcm[i % 4].data_ptr(), 0, 0, st.cuda_stream)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            eng.step_device(B, d_slots[i % groups].data_ptr(), d_pcm[i % 4].data_ptr(), 0, 0, st.cuda_stream)
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # e2e host path
    eng.h_slots[:B] = gs[0]
    eng.h_pcm[:B] = pcm[:, :chunk]
    for _ in range(5):
        eng.step_pinned(B)
    t0 = time.perf_counter()
    n2 = max(10, steps // 4)
    for _ in range(n2):
        eng.step_pinned(B)
    e2e = (time.perf_counter() - t0) / n2 * 1e3
    rtf = B * chunk / 8000.0 / (ms / 1e3)
    tf = B * (1.2877e9 if chunk == 2400 else 1.6316e9) / (ms / 1e3) / 1e12
    print(f"B={B:5d} chunk={chunk} graph={int(use_graph)} pdl={os.environ.get('TONE_PDL','1')}: {ms*1e3:8.1f} us/step  "
          f"RTFx {rtf:9.0f}  {tf:6.1f} TFLOP/s  launches {eng._get_info().launches_per_step}  e2e {e2e*1e3:8.1f} us", flush=True)
    eng.close()


if __name__ == "__main__":
    weights = tb.weights.init_weights(0)
    Bs = [int(x) for x in sys.argv[1:]] or [64]
    for B in Bs:
        run(B)
