"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import sys


def main(path, out, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg = collections.OrderedDict()
    total = 0.0
    n = 0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v = v / 1000 if unit in ("ns", "nsecond") else (v * 1000 if unit in ("ms", "msecond") else v)
        name = r["Kernel Name"]
        grid = r.get("Grid Size", "")
        key = (name, grid)
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
        n += 1
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                "(cold-cache, serialised: compare SHARES, not absolutes).\n\n")
        f.write(f"launches: {n}, sum of durations: {total:.1f} us\n\n")
        f.write("| kernel | grid | launches | sum us | avg us | share |\n|---|---|---|---|---|---|\n")
        for (name, grid), (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{name[:70]}` | {grid} | {c} | {t:.1f} | {t / c:.2f} | {100 * t / total:.1f} % |\n")
    print(open(out).read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "launch list")
