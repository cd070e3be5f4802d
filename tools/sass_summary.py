"""SASS evidence table: per kernel of libtone_b200.so, how many tcgen05 MMAs (UTC*MMA), TMEM loads / stores (LDTM / STTM),
TMA loads (UTMALDG / UBLKCP), TMA prefetches (UTMAPF), tcgen05 commits (UTCBAR), legacy tensor ops (HMMA) and packed fp32
FMAs (FFMA2) the compiled code contains.  Usage: python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "t-one_b200", "libtone_b200.so")
PATS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTMALDG", r"\bUTMALDG"), ("UTMALDG.2CTA", r"\bUTMALDG\.[0-9D.]*2CTA"), ("UBLKCP", r"\bUBLKCP"), ("UTMAPF", r"\bUTMAPF"),
        ("UTCBAR", r"\bUTCBAR"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("FFMA2", r"\bFFMA2"), ("MUFU", r"\bMUFU")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, lines = collections.OrderedDict(), None, collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        lines[cur] += 1
        for name, pat in PATS:
            if re.search(pat, line):
                counts[cur][name] += 1
    dm = demangle(list(counts))
    print("# SASS evidence per kernel (`cuobjdump -sass t-one_b200/libtone_b200.so`, built with `nvcc -gencode arch=compute_100a,code=sm_100a`)\n")
    print("`UTC*MMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` = cp.async.bulk.tensor (TMA), `UBLKCP` = cp.async.bulk, "
          "`UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier ops, `HMMA` = legacy mma.sync (only the 160-point framed DFT of the log-mel "
          "front end, 0.15 % of the FLOPs: an fp16 hi / lo split basis on m16n8k16, too small for a tcgen05 tile).\n")
    print("| kernel | SASS lines | " + " | ".join(n for n, _ in PATS) + " |")
    print("|---|---|" + "---|" * len(PATS))
    tot = collections.Counter()
    for k, c in counts.items():
        name = dm.get(k, k).replace("tone::", "").replace("CUtensorMap_st", "TMap")
        name = re.sub(r"\(.*", "", name)
        print(f"| `{name[:70]}` | {lines[k]} | " + " | ".join(str(c.get(n, 0)) for n, _ in PATS) + " |")
        tot.update(c)
    print(f"| **total ({len(counts)} kernels)** | {sum(lines.values())} | " + " | ".join(str(tot.get(n, 0)) for n, _ in PATS) + " |")


if __name__ == "__main__":
    main()
