"""Summarise an `ncu --replay-mode app-range --cache-control none --clock-control none --metrics
dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,gpu__time_duration.sum`
capture of tools/gpu_one.py (the cudaProfilerStart / Stop range = `steps` consecutive LIVE steps: warm caches, lanes
running concurrently, nothing replayed per kernel) into DRAM / L2 bytes per step.
Usage: python tools/summarize_range.py profiles/r02_range_B1024_16.csv 16 profiles/r02_dram_live_B1024.json"""
import csv
import json
import sys


def main(path, steps, out):
    steps = int(steps)
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    m = {r["Metric Name"]: float(r["Metric Value"].replace(",", "")) for r in rows}
    d = {"steps_in_range": steps,
         "dram_read_bytes": m["dram__bytes_read.sum"] / steps,
         "dram_write_bytes": m["dram__bytes_write.sum"] / steps,
         "dram_total_bytes": (m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]) / steps,
         "l2_read_bytes": m["lts__t_sectors_op_read.sum"] * 32 / steps,
         "l2_write_bytes": m["lts__t_sectors_op_write.sum"] * 32 / steps,
         "range_ms_per_step_under_ncu": m["gpu__time_duration.sum"] / 1e6 / steps,
         "mode": "ncu --replay-mode app-range --cache-control none --clock-control none: one range of consecutive live steps",
         "source": path}
    with open(out, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
