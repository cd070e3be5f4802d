"""Step time of the engine at given batch sizes on the B200 box (device-timed, PCM resident, rotating slot groups).
Usage: python tools/gpu_perf.py [B ...] [key=value ...]   e.g.  python tools/gpu_perf.py 1024 64 fused_ff=1 lanes=1
Engine keyword arguments (lanes, lane_min_batch, persist_min_tiles, persist_mode, split_k, flags, fused_ff,
fused_ff_min_rows, chunk) are passed through."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")


def run(B, chunk=2400, steps=60, warm=10, **kw):
    G = 2 if B >= 512 else 8
    eng = tb.Engine(tb.weights.init_weights(0), chunk_samples=chunk, max_slots=B * G, max_batch=B, **kw)
    groups = [eng.alloc_slots(B) for _ in range(G)]
    pcm = tb.synth.telephony_pcm(min(B, 128), chunk * 4, seed=1).reshape(-1, 4, chunk)
    pcm = np.ascontiguousarray(np.tile(pcm, ((B + pcm.shape[0] - 1) // pcm.shape[0], 1, 1))[:B].transpose(1, 0, 2)).astype(np.int16)
    d_pcm = torch.from_numpy(pcm).cuda()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(warm):
            eng.step_device(groups[i % G], d_pcm[i % 4].data_ptr(), tb.model.PCM_I16, 0, 0, st.cuda_stream)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(steps):
            eng.step_device(groups[i % G], d_pcm[i % 4].data_ptr(), tb.model.PCM_I16, 0, 0, st.cuda_stream)
        e1.record(st)
        st.synchronize()
    ms = e0.elapsed_time(e1) / steps
    flops = {2400: 1_287_738_880, 3200: 1_631_636_224}[chunk] * B
    print(f"B={B:5d} chunk={chunk} {kw} : {ms:.4f} ms/step  {B * chunk / 8000 / ms * 1e3:9.0f} RTFx  "
          f"{flops / ms / 1e9:7.1f} TFLOP/s ({flops / ms / 1e9 / 1412.7 * 100:.1f} %)  launches {eng._get_info().launches_per_step}", flush=True)
    eng.close()
    return ms


if __name__ == "__main__":
    Bs, kw = [], {}
    for a in sys.argv[1:]:
        if "=" in a:
            k, v = a.split("=")
            kw[k] = int(v)
        else:
            Bs.append(int(a))
    chunk = kw.pop("chunk", 2400)
    for B in Bs or [64, 1024]:
        run(B, chunk, **kw)
