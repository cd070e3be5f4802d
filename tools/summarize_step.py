"""Summarise an ncu per-launch CSV of ONE step (tools/gpu_one.py B steps=1 under
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum`)
into a per-kernel table: launches, time share, DRAM read / write bytes and the bytes written into / read from L2.

ncu replays every launch on its own after flushing the caches, so DRAM reads are an upper bound of the live traffic
(everything a kernel touches is a miss) while DRAM writes are a LOWER bound: a kernel's output usually still sits in the
126 MB write-back L2 when the kernel ends, which is why `dram__bytes_write.sum` reads near zero for most launches.  The
L2 write sectors (x 32 B) are what the kernels actually wrote.
Usage: python tools/summarize_step.py gpurun_out/r02_step_B1024.csv profiles/r02_step_B1024 "title" """
import collections
import csv
import json
import sys

UNIT_B = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
UNIT_T = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("tone::", "")
    return n[:64]


def main(path, out, title):
    with open(path) as f:
        rows = list(csv.DictReader(l for l in f if not l.startswith("==")))
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        m, v, u = r["Metric Name"], float(r["Metric Value"].replace(",", "")), r["Metric Unit"]
        if m.startswith("dram__bytes"):
            d[m] = v * UNIT_B.get(u, 1)
        elif m.startswith("lts__t_sectors"):
            d[m] = v * 32.0
        elif m == "gpu__time_duration.sum":
            d["us"] = v * UNIT_T.get(u, 1)
    ls = [l for l in launches.values() if "at::" not in l["name"]]
    agg = collections.OrderedDict()
    keys = ("us", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum")
    tot = dict.fromkeys(keys, 0.0)
    for l in ls:
        a = agg.setdefault((short(l["name"]), l["grid"]), dict.fromkeys(keys, 0.0) | {"n": 0})
        a["n"] += 1
        for k in keys:
            a[k] += l.get(k, 0.0)
            tot[k] += l.get(k, 0.0)
    summary = {"launches": len(ls), "sum_kernel_us_serialised": tot["us"], "dram_read_bytes": tot["dram__bytes_read.sum"],
               "dram_write_bytes": tot["dram__bytes_write.sum"],
               "dram_total_bytes": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
               "l2_write_bytes": tot["lts__t_sectors_op_write.sum"], "l2_read_bytes": tot["lts__t_sectors_op_read.sum"], "source": path}
    with open(out + ".json", "w") as f:
        json.dump(summary, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# {title}\n\n")
        f.write("One step under `ncu --clock-control none` (per-launch replay, caches flushed before every launch: compare SHARES of time, "
                "not absolutes; DRAM reads are an upper bound of the live traffic, DRAM writes a lower bound - outputs stay in the "
                "write-back L2 - so the L2 write bytes are listed beside them).\n\n")
        f.write(f"launches {len(ls)}, sum of durations {tot['us']:.1f} us; DRAM read {tot['dram__bytes_read.sum'] / 1e6:.1f} MB, "
                f"DRAM write {tot['dram__bytes_write.sum'] / 1e6:.1f} MB, written into L2 {tot['lts__t_sectors_op_write.sum'] / 1e6:.1f} MB, "
                f"read from L2 {tot['lts__t_sectors_op_read.sum'] / 1e6:.1f} MB\n\n")
        f.write("| kernel | grid | launches | sum us | avg us | share | DRAM rd MB | DRAM wr MB | L2 wr MB | L2 rd MB |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for (name, grid), a in sorted(agg.items(), key=lambda x: -x[1]["us"]):
            f.write(f"| `{name}` | {grid} | {a['n']} | {a['us']:.1f} | {a['us'] / a['n']:.2f} | {100 * a['us'] / max(tot['us'], 1e-9):.1f} % | "
                    f"{a['dram__bytes_read.sum'] / 1e6:.2f} | {a['dram__bytes_write.sum'] / 1e6:.2f} | "
                    f"{a['lts__t_sectors_op_write.sum'] / 1e6:.2f} | {a['lts__t_sectors_op_read.sum'] / 1e6:.2f} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "one step, per kernel")
