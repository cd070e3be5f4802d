"""In-kernel timeline of the cluster (latency) path on the B200 box: python tools/gpu_cluster_prof.py [B] [layer]
Prints, for cluster 0 / CTA 0, the raw mark-to-mark durations of one layer (worker thread 0) and the MMA lane's per-GEMM
time and weight-stall share.  Mark order follows ClWorker::run_group / reduce_tail in csrc/encoder_cluster.cuh."""
import ctypes
import importlib
import os
import sys

import numpy as np

os.environ.setdefault("TONE_CL_PROF", "1")
os.environ.setdefault("TONE_CLUSTER_MAX_B", "128")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")

B = int(sys.argv[1]) if len(sys.argv) > 1 else 60
LAYER = int(sys.argv[2]) if len(sys.argv) > 2 else 2
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
print("rc", rc, "max active clusters (G small*1000 + G large)", ma.value)
marks = buf[:2048].astype(np.int64)
n = int((marks > 0).sum())
print("worker marks", n, "total us", (marks[n - 1] - marks[0]) / 1e3)
# tail of a sub-block (reduce_tail): wait_acc | prefetch | wait_z | send:tmem+stores | send:signal | wait_x | owner:rows | owner:signal
tail = ["wait_acc2", "prefetch", "wait_z", "send:stores", "send:signal", "wait_x", "owner:rows", "owner:signal"]
names = (["ff1:wait_acc", "ff1:swiglu", "ff1:rel_h"] + ["ff1:" + t for t in tail] +
         ["att:wait_acc", "att:core+z", "att:rel_h"] + ["att:" + t for t in tail] +
         ["cv:wait_acc", "cv:dw+z", "cv:rel_h"] + ["cv:" + t for t in tail] +
         ["ff2:wait_acc", "ff2:swiglu", "ff2:rel_h"] + ["ff2:" + t for t in tail])
per = len(names)
# marks before layer 0: initial owner_rows emits 1 (before signal) + 1 after = 2
pos = 2
d_all = np.diff(marks[:n]) / 1e3
for l in range(16):
    k = per
    if l == 6:
        k += len(tail)           # reduction: reduce_tail only (exchange itself has no marks)
    if l == 14:
        k += 2                   # upsample owner_rows: 1 inside + 1 after
    seg = d_all[pos - 1:pos - 1 + k]
    print(f"layer {l}: total {seg.sum():.1f} us")
    if l == LAYER:
        for nm, v in zip(names, seg[:per]):
            print(f"   {nm:16s} {v:7.2f}")
    pos += k
mm = buf[2048:].astype(np.int64).reshape(-1, 4)
nm_ = int((mm[:, 0] > 0).sum())
mm = mm[:nm_]
dur = (mm[:, 1] - mm[:, 0]) / 1e3
stall = mm[:, 2] / 1.965e3
print("mma ops", nm_, "sum gemm us", round(dur.sum(), 1), "sum stall us", round(stall.sum(), 1), "chunks", mm[:, 3].sum())
ops_per_layer = 9
for i in range(ops_per_layer * 1 + 0, ops_per_layer * 2):
    if i < nm_:
        print(f"   op {i}: chunks {mm[i,3]:3d} dur {dur[i]:7.2f} us stall {stall[i]:7.2f} us")
eng.close()
