"""Profiling target: `warm` untimed steps, then `steps` steps of the engine at batch B between cudaProfilerStart / Stop
(run under `ncu --profile-from-start off ...`), device-resident int16 PCM, rotating slot groups.
Usage: python tools/gpu_one.py B [steps=2] [warm=5] [engine kwargs: lanes=.. fused_ff=.. chunk=..]"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")


def main():
    B = int(sys.argv[1])
    kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}
    steps, warm, chunk = kw.pop("steps", 2), kw.pop("warm", 5), kw.pop("chunk", 2400)
    G = 2
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
        torch.cuda.profiler.start()
        for i in range(steps):
            eng.step_device(groups[i % G], d_pcm[i % 4].data_ptr(), tb.model.PCM_I16, 0, 0, st.cuda_stream)
        st.synchronize()
        torch.cuda.profiler.stop()
    print(f"B={B} steps={steps} launches_per_step={eng._get_info().launches_per_step}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
