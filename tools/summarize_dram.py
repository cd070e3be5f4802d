"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list:
finds one complete step (begin_step_kernel ... decoder GEMM), sums DRAM traffic and writes a JSON + markdown summary.
Usage: python tools/summarize_dram.py gpurun_out/r01_dram.csv profiles/r01_dram_traffic_B64"""
import collections
import csv
import json
import sys


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)


def main(path, out):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            d[m] = to_bytes(r["Metric Value"], r["Metric Unit"])
        elif m == "gpu__time_duration.sum":
            d["us"] = to_us(r["Metric Value"], r["Metric Unit"])
    ls = list(launches.values())
    starts = [i for i, l in enumerate(ls) if "begin_step_kernel" in l["name"]]
    assert len(starts) >= 2, "need one complete step in the capture"
    step = ls[starts[0]:starts[1]]
    step = [l for l in step if "at::" not in l["name"]]
    rd = sum(l.get("dram__bytes_read.sum", 0) for l in step)
    wr = sum(l.get("dram__bytes_write.sum", 0) for l in step)
    agg = collections.OrderedDict()
    for l in step:
        a = agg.setdefault(l["name"].split("(")[0][:60], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += l.get("dram__bytes_read.sum", 0)
        a[2] += l.get("dram__bytes_write.sum", 0)
        a[3] += l.get("us", 0)
    summary = {"launches": len(step), "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_total_bytes": rd + wr,
               "sum_kernel_us_serialised": sum(l.get("us", 0) for l in step), "source": path}
    with open(out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write("# DRAM traffic of one step (ncu, per-launch `dram__bytes_read.sum + dram__bytes_write.sum`)\n\n")
        f.write(f"launches {len(step)}; DRAM read {rd / 1e6:.1f} MB, write {wr / 1e6:.1f} MB, total {(rd + wr) / 1e6:.1f} MB per step "
                "(serialised replay: every kernel starts with a cold L1, L2 keeps what the previous launch left).\n\n")
        f.write("| kernel | launches | read MB | write MB | sum us |\n|---|---|---|---|---|\n")
        for k, (n, r_, w_, us) in sorted(agg.items(), key=lambda x: -(x[1][1] + x[1][2])):
            f.write(f"| `{k}` | {n} | {r_ / 1e6:.2f} | {w_ / 1e6:.2f} | {us:.1f} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
