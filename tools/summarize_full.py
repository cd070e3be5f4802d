"""Summarise an `ncu --set full` report into a markdown table (one row per captured launch).

usage: python tools/summarize_full.py gpurun_out/x.{ncu-rep,csv} "title" > profiles/x.md
Reads the report with `ncu -i ... --page raw --csv`; picks the counters the north_star asks for: tensor-pipe
utilisation for the GEMMs, DRAM bytes / throughput for the bandwidth kernels.
"""
import csv
import io
import subprocess
import sys

COLS = [
    ("duration us", "gpu__time_duration.sum"),
    ("DRAM rd MB", "dram__bytes_read.sum"),
    ("DRAM wr MB", "dram__bytes_write.sum"),
    ("DRAM % peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 % peak", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor pipe % (active)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("tensor pipe % (elapsed)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("SM busy %", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps active %", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("regs", "launch__registers_per_thread"),
]

UNIT_SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3,
              "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}

KINDS = {0: "STORE_F32", 1: "RESID", 2: "SWIGLU", 3: "GLU", 4: "CONV0", 5: "CONV1", 6: "KV", 7: "DECODER",
         8: "PARTIAL", 9: "GLU_DW", 10: "VATT"}


def main():
    rep, title = sys.argv[1], sys.argv[2]
    if rep.endswith(".csv"):   # already exported on the GPU box
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def find(metric):
        if metric in idx:
            return idx[metric]
        for h, i in idx.items():
            if h.startswith(metric):
                return i
        return None

    print(f"# {title}\n")
    print("Read with `ncu -i <report> --page raw --csv` (tools/summarize_full.py).  Cold caches, serialised replay: use the")
    print("percentages and byte counts, not the absolute durations.\n")
    print("| # | kernel | grid | " + " | ".join(c for c, _ in COLS) + " |")
    print("|---|---|---|" + "---|" * len(COLS))
    for n, r in enumerate(data):
        name = r[idx["Kernel Name"]]
        if name.startswith("void "):
            name = name[5:]
        if name.startswith("gemm_tc_kernel<"):
            try:
                k = int(name.split("<")[1].split(",")[0])
                name = name.split("(")[0] + " " + KINDS.get(k, "?")
            except ValueError:
                pass
        name = name.split("(")[0] if "gemm_tc" not in name else name
        grid = r[idx["Grid Size"]].replace(" ", "")
        cells = []
        for _, m in COLS:
            i = find(m)
            if i is None or r[i] == "":
                cells.append("-")
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if u in UNIT_SCALE:
                v *= UNIT_SCALE[u]
            cells.append(f"{v:.2f}" if v < 100 else f"{v:.0f}")
        print(f"| {n} | `{name}` | {grid} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
