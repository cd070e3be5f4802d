"""Per-kernel roofline table for the 1024-stream step (2 lanes x 512 streams = 5120 rows per lane, 2560 at the reduced
rate): algorithmic FLOPs (GEMM kinds) or algorithmic HBM bytes (bandwidth kernels) per launch divided by the kernel's ncu
duration (`profiles/r02_step_B1024.csv`: cold-cache, serialised replay - so these are LOWER bounds of what the kernels
reach inside a live step, where two lanes overlap and L2 is warm), against the measured peaks of MEASURED_PEAKS.json
(burst figures: a kernel timed alone).  Usage: python tools/roofline_table.py > profiles/r02_roofline_B1024.md"""
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"bf16_tflops": 1672.7, "hbm_gbs": 6545.0}
S, T, T2, D, DFF = 512, 10, 5, 384, 1536        # streams per lane, frames, reduced frames, d_model, d_ff


def main():
    path = os.path.join(ROOT, "profiles", "r02_step_B1024.csv")
    with open(path) as f:
        rws = list(csv.DictReader(l for l in f if not l.startswith("==")))
    launches = collections.OrderedDict()
    for r in rws:
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": tuple(int(x) for x in re.findall(r"\d+", r["Grid Size"]))})
        if r["Metric Name"] == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            d["us"] = v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3}.get(r["Metric Unit"], 1)
    # shapes per (kernel, grid): rows of the lane at that point of the graph
    R, R2 = S * T, S * T2
    F = lambda m, n, k: 2.0 * m * n * k                             # noqa: E731
    table = [
        # (regex on name, grid, label, FLOP, HBM bytes)
        (r"persist_kernel<2,", (148, 1, 1), "FF up + SwiGLU (persistent, 256-wide), rows 5120 | 2560 alternate", (F(R, 2 * DFF, D) + F(R2, 2 * DFF, D)) / 2, None),
        (r"gemm_tc_kernel<8, 128", (40, 3, 1), "FF down (partial sums), 5120 rows", F(R, D, DFF), None),
        (r"gemm_tc_kernel<8, 128", (20, 3, 2), "FF down (split-K 2), 2560 rows", F(R2, D, DFF), None),
        (r"gemm_tc_kernel<1, 128", (40, 3, 1), "W_o / pw2 + residual, 5120 rows", F(R, D, D), None),
        (r"gemm_tc_kernel<1, 128", (20, 3, 1), "W_o / pw2 + residual, 2560 rows", F(R2, D, D), None),
        (r"persist_kernel<3,", (120, 1, 1), "conv-module pw1 + GLU (persistent), 5120 rows", F(R, 2 * D, D), None),
        (r"gemm_tc_kernel<3, 128", (20, 6, 1), "conv-module pw1 + GLU, 2560 rows", F(R2, 2 * D, D), None),
        (r"gemm_tc_kernel<0, 128", (40, 3, 1), "V / Q projection, 5120 rows", F(R, D, D), None),
        (r"gemm_tc_kernel<0, 128", (20, 3, 1), "V / Q projection, 2560 rows", F(R2, D, D), None),
        (r"gemm_tc_kernel<5, 128", (43, 17, 1), "conv1 implicit GEMM (K = 11 x 384), 512 streams", 2.0 * S * T * 34 * 64 * 32 * 121, None),
        (r"gemm_tc_kernel<4, 128", (128, 11, 1), "conv0 implicit GEMM (banded), 512 streams", 2.0 * S * 30 * 44 * 32 * 231, None),
        (r"norm_kernel", (640, 1, 1), "RMSNorm + partial fold, 5120 rows (r fp32 r/w, partial fp16, n bf16)", None, R * D * (4 + 4 + 2 + 2)),
        (r"norm_kernel", (320, 1, 1), "RMSNorm + partial fold, 2560 rows", None, R2 * D * (4 + 4 + 2 * 2 + 2)),
        (r"dwconv_kernel<5>", (512, 2, 1), "depthwise conv k=31, T=10 (cache 30 rows r/w, g in, e out)", None, S * D * 2 * (30 + 30 + T + T)),
        (r"dwconv_kernel<3>", (512, 2, 1), "depthwise conv k=31, T=5", None, S * D * 2 * (30 + 30 + T2 + T2)),
        (r"attention_kernel<0>", (512, 1, 1), "P.V with shared scores (V fp32 in, P in, ctx bf16 out)", None, S * T * D * (4 + 2) + S * 8 * T * T * 4),
        (r"begin_step_kernel", (512, 1, 1), "log-mel front end + cache rolls (PCM int16 in, 70 KB of rolls per stream)", None, S * (2400 * 2 + 2 * 70000)),
    ]
    agg = collections.OrderedDict()
    for l in launches.values():
        for pat, grid, label, flop, byts in table:
            if re.search(pat, l["name"]) and l["grid"] == grid:
                a = agg.setdefault(label, [0, 0.0, flop, byts])
                a[0] += 1
                a[1] += l.get("us", 0.0)
                break
    print("# Round 2 - per-kernel roofline, 1024 streams (2 lanes x 512 streams)\n")
    print("Algorithmic FLOPs (GEMM kinds) or algorithmic HBM bytes (bandwidth kernels) per launch / the launch's ncu duration "
          "(`profiles/r02_step_B1024.csv`: cold caches, serialised replay - a lower bound of the live rate), against the measured "
          f"burst peaks ({PEAK['bf16_tflops']:.0f} TFLOP/s bf16, {PEAK['hbm_gbs']:.0f} GB/s; `MEASURED_PEAKS.json`).  The whole step, "
          "timed live by `bench.py`, runs at 450 TFLOP/s = 31.9 % of the sustained peak (1412.7).\n")
    print("| kernel | launches | avg us | algorithmic work per launch | achieved | of measured peak |")
    print("|---|---|---|---|---|---|")
    for label, (n, us, flop, byts) in agg.items():
        avg = us / n
        if flop:
            tf = flop / avg / 1e6
            print(f"| {label} | {n} | {avg:.2f} | {flop / 1e9:.2f} GFLOP | {tf:.0f} TFLOP/s | {100 * tf / PEAK['bf16_tflops']:.1f} % (tensor) |")
        elif byts:
            gbs = byts / avg / 1e3
            print(f"| {label} | {n} | {avg:.2f} | {byts / 1e6:.1f} MB | {gbs:.0f} GB/s | {100 * gbs / PEAK['hbm_gbs']:.1f} % (HBM) |")
        else:
            print(f"| {label} | {n} | {avg:.2f} | - | - | - |")


if __name__ == "__main__":
    main()
