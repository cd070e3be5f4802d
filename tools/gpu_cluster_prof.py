"""In-kernel timeline of the cluster (latency) path on the B200 box: TONE_CL_PROF=1 python tools/gpu_cluster_prof.py [B]
Prints, for cluster 0 / CTA 0, the phase durations of selected layers (worker thread 0) and the MMA lane's per-GEMM
time and weight-stall share."""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

os.environ.setdefault("TONE_CL_PROF", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
weights = tb.weights.init_weights(0)
eng = tb.Engine(weights, chunk_samples=2400, max_slots=B, max_batch=B, use_graph=True)
slots = eng.alloc_slots(B)
pcm = tb.synth.telephony_pcm(B, 2400 * 6, seed=1)
for i in range(6):
    eng.step(slots, pcm[:, i * 2400:(i + 1) * 2400])
lib = tb.load_library()
buf = np.zeros(6144, dtype=np.uint64)
ma = ctypes.c_int32(0)
lib.tone_cluster_prof_read.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]
rc = lib.tone_cluster_prof_read(eng._h, buf.ctypes.data, ctypes.byref(ma))
print("rc", rc, "max active clusters", ma.value)
marks = buf[:2048].astype(np.int64)
n = int((marks > 0).sum())
print("worker marks", n, "total us", (marks[n - 1] - marks[0]) / 1e3)
# marks per layer: 1 (initial owner) then per layer 24 (+6 for l==6, +2 for l==14)
names = ["ff1:wait_up", "ff1:swiglu", "ff1:rel_h", "ff1:wait_dn", "ff1:send", "ff1:wait_x", "ff1:owner",
         "att:wait_qkv", "att:core", "att:rel_h", "att:wait_wo", "att:send", "att:wait_x", "att:owner",
         "cv:wait_pw1", "cv:dw", "cv:rel_h", "cv:wait_pw2", "cv:send", "cv:wait_x", "cv:owner",
         "ff2:wait_up", "ff2:swiglu", "ff2:rel_h", "ff2:wait_dn", "ff2:send", "ff2:wait_x", "ff2:owner"]
pos = 1
for l in range(16):
    k = len(names) + (6 if l == 6 else 0) + (2 if l == 14 else 0)
    seg = marks[pos - 1:pos + k]
    d = np.diff(seg) / 1e3
    if l in (1, 8, 15):
        print(f"layer {l}: total {d.sum():.1f} us")
        for nm, v in zip(names, d[:len(names)]):
            print(f"   {nm:14s} {v:7.2f}")
    else:
        print(f"layer {l}: total {d.sum():.1f} us")
    pos += k
mm = buf[2048:].astype(np.int64).reshape(-1, 4)
nm_ = int((mm[:, 0] > 0).sum())
mm = mm[:nm_]
dur = (mm[:, 1] - mm[:, 0]) / 1e3
stall = mm[:, 2] / 1.965e3
print("mma ops", nm_, "sum gemm us", dur.sum(), "sum stall us", stall.sum(), "chunks", mm[:, 3].sum())
# layer 1 ops: index 8.. (layer 0 has 8 ops)
for i in range(8, 16):
    print(f"   op {i}: chunks {mm[i,3]:3d} dur {dur[i]:7.2f} us stall {stall[i]:7.2f} us")
eng.close()
