"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table.
usage: python tools/summarize_launches.py gpurun_out/launches_r01.csv > profiles/r01_launches.md"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
agg, tot = collections.OrderedDict(), 0.0
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    key = (re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki])), r[gi])
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1; a[1] += v; tot += v
print(f"# ncu launch list summary: {sys.argv[1]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised launches: compare SHARES, not absolutes).\n")
print(f"{len(data)} launches, {tot:.1f} us total\n")
print("| kernel | grid | launches | total us | share | avg us |\n|---|---|---:|---:|---:|---:|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[0]}` | {k[1]} | {c} | {t:.1f} | {100 * t / tot:.1f}% | {t / c:.2f} |")
