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
        (r"persist_kernel<2,", (120, 1, 1), "FF up + SwiGLU (persistent, 256-wide tiles), rows 5120 / 2560 alternate", (F(R, 2 * DFF, D) + F(R2, 2 * DFF, D)) / 2, None),
        (r"gemm_tc_kernel<8, 128", (40, 3, 1), "FF2 down (fp16 partial sums), 5120 rows", F(R, D, DFF), None),
        (r"gemm_tc_kernel<8, 128", (20, 3, 2), "FF down (split-K 2), 2560 rows", F(R2, D, DFF), None),
        (r"gemm_tc_kernel<1, 128", (40, 3, 1), "W_o / pw2 (K = 384: 32 launches) and FF1 down (K = 1536: 14) + residual, 5120 rows", (32 * F(R, D, D) + 14 * F(R, D, DFF)) / 46, None),
        (r"gemm_tc_kernel<1, 128", (20, 3, 1), "W_o / pw2 + residual, 2560 rows", F(R2, D, D), None),
        (r"persist_kernel<3,", (120, 1, 1), "conv-module pw1 + GLU (persistent), 5120 rows", F(R, 2 * D, D), None),
        (r"gemm_tc_kernel<3, 128", (20, 6, 1), "conv-module pw1 + GLU, 2560 rows", F(R2, 2 * D, D), None),
        (r"gemm_tc_kernel<10, 48", (43, 8, 1), "V projection + P.V (one head per CTA), 5120 rows", F(R, D, D), None),
        (r"gemm_tc_kernel<10, 48", (21, 8, 1), "V projection + P.V, 2560 rows", F(R2, D, D), None),
        (r"gemm_tc_kernel<0, 128", (40, 3, 1), "Q projection (layers 14 / 15 full rate), 5120 rows", F(R, D, D), None),
        (r"gemm_tc_kernel<0, 128", (20, 3, 1), "Q projection / reduction pointwise, 2560 rows", F(R2, D, D), None),
        (r"persist_kernel<0, 1", (120, 1, 1), "q, k, v projection, layer 0 (N = 1152), 5120 rows", F(R, 3 * D, D), None),
        (r"persist_kernel<0, 1", (90, 1, 1), "q, k, v projection, layer 7, 2560 rows", F(R2, 3 * D, D), None),
        (r"gemm_tc_kernel<6, 128", (171, 6, 1), "K, V projection of [cache, new] rows, layer 15 (40 rows per stream)", F(S * 40, 2 * D, D), None),
        (r"gemm_tc_kernel<6, 128", (86, 6, 1), "K, V projection, layer 14 (20 rows per stream)", F(S * 20, 2 * D, D), None),
        (r"gemm_tc_kernel<5, 128", (43, 17, 1), "conv1 implicit GEMM (K = 11 x 384), 512 streams", 2.0 * S * T * 34 * 64 * 32 * 121, None),
        (r"gemm_tc_kernel<4, 128", (128, 11, 1), "conv0 implicit GEMM (banded), 512 streams", 2.0 * S * 30 * 44 * 32 * 231, None),
        (r"norm_kernel", (640, 1, 1), "RMSNorm + partial fold, 5120 rows (r fp32 r/w, partial fp16, n bf16)", None, R * D * (4 + 4 + 2 + 2)),
        (r"norm_kernel", (320, 1, 1), "RMSNorm + partial fold, 2560 rows", None, R2 * D * (4 + 4 + 2 * 2 + 2)),
        (r"dwconv_pipe_kernel<10>", (296, 1, 1), "depthwise conv k=31, T=10 (ring cache: 30 rows in, T rows out; g in, e out)", None, S * D * 2 * (30 + T + T + T)),
        (r"dwconv_pipe_kernel<5>", (296, 1, 1), "depthwise conv k=31, T=5", None, S * D * 2 * (30 + T2 + T2 + T2)),
        (r"attention_pipe_kernel<512>", (148, 1, 1), "cached-context attention, layers 14 / 15 (k, v fp32 in: 20 / 40 rows, q, P out, ctx out), average", None,
         S * (30 * 2 * D * 4 + 1.5 * T2 * D * 4 + 8 * 7.5 * 30 * 4 + 7.5 * D * 2)),
        (r"attention_pipe_kernel<256>", (296, 1, 1), "attention of layers 0 / 7 (q, k, v fp32 in, P out, ctx out), average", None,
         S * (7.5 * 3 * D * 4 + 8 * (100 + 25) / 2 * 4 + 7.5 * D * 2)),
        (r"begin_step_kernel", (148, 1, 1), "log-mel front end + cache rolls (PCM int16 in, 57 KB of rolls per stream)", None, S * (2400 * 2 + 2 * 57088 + 30 * 64 * 2)),
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
          "timed live by `bench.py`, runs at 498 TFLOP/s = 35.3 % of the sustained peak (1412.7).\n")
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
