"""In-kernel timeline of one step (diagnostic build libtone_b200_prof.so): per kernel start, duration, gap, phases."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["TONE_B200_LIB"] = os.path.join(ROOT, "t-one_b200", "libtone_b200_prof.so")
sys.path.insert(0, ROOT)
tb = importlib.import_module("t-one_b200")

NAMES = {1: "begin_step", 2: "norm", 3: "upsample_norm", 4: "attention", 5: "dwconv", 6: "reduction_dw", 7: "ff_fused", 8: "att_block", 9: "rowgemm"}
KINDS = ["store_f32", "resid", "swiglu", "glu", "conv0", "conv1", "kv", "decoder", "partial", "glu_dw", "vatt"]


def name(i):
    if i >= 1000:
        k, bn = (i - 1000) // 100, ((i - 1000) % 100) * 8
        return f"gemm_{KINDS[k]}_bn{bn}"
    return NAMES.get(int(i), str(i))


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    graph = int(os.environ.get("GRAPH", "1"))
    kw = {k: int(v) for k, v in (a.split("=") for a in sys.argv[2:])}     # engine keyword arguments, e.g. lanes=1 fused_ff=3
    eng = tb.Engine(tb.weights.init_weights(0), max_slots=B, max_batch=B, use_graph=bool(graph), **kw)
    lib = eng._lib
    lib.tone_prof_start.argtypes = [C.c_void_p, C.c_int32]
    lib.tone_prof_read.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32, C.POINTER(C.c_int32)]
    slots = eng.alloc_slots(B)
    pcm = tb.synth.telephony_pcm(min(B, 64), 2400, seed=1)
    pcm = np.tile(pcm, (B // min(B, 64) + 1, 1))[:B]
    for _ in range(5):
        eng.step(slots, pcm)
    assert lib.tone_prof_start(eng._h, 4096) == 0
    for _ in range(2):
        eng.step(slots, pcm)
    buf = np.zeros((4096, 10), dtype=np.uint64)
    n = C.c_int32()
    assert lib.tone_prof_read(eng._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), 4096, C.byref(n)) == 0
    n = n.value
    rec = buf[:n].astype(np.int64)
    per = n // 2
    rec = rec[per:]                      # second step
    t0 = rec[0, 0]
    lines = [f"# B={B} graph={graph} {kw}  kernels/step={per}",
             "# seq name grid start_us dur_us gap_us | cycles: prologue wait first_stage acc_ready total"]
    prev_end = t0
    agg = {}
    for i, r in enumerate(rec):
        g0, g1, c = r[0], r[1], r[2:8]
        nm = name(r[8])
        dur, gap = (g1 - g0) / 1e3, (g0 - prev_end) / 1e3
        ph = [int(c[j] - c[0]) if c[j] else 0 for j in range(1, 6)]
        lines.append(f"{i:4d} {nm:22s} {int(r[9]):5d} {(g0 - t0) / 1e3:9.2f} {dur:7.2f} {gap:7.2f} | "
                     + " ".join(f"{x:7d}" for x in ph))
        a = agg.setdefault(nm, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += dur
        a[2] += gap
        prev_end = max(prev_end, g1)
    total = (rec[-1, 1] - t0) / 1e3
    lines.append(f"# step span {total:.1f} us")
    lines.append("# per-kernel-type: count  sum_dur_us  sum_gap_us  avg_dur  avg_gap")
    for nm, (c, d, g) in sorted(agg.items(), key=lambda x: -(x[1][1] + x[1][2])):
        lines.append(f"#   {nm:22s} {c:4d} {d:9.1f} {g:9.1f} {d / c:7.2f} {g / c:7.2f}")
    out = "\n".join(lines)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    tag = f"timeline_B{B}_g{graph}_" + "_".join(f"{k}{v}" for k, v in kw.items())
    with open(os.path.join(ROOT, "gpurun_out", tag + ".txt"), "w") as f:
        f.write(out + "\n")
    print("\n".join(lines[:40]))
    print("...")
    print("\n".join(lines[-16:]))


if __name__ == "__main__":
    main()
