"""Sum an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` launch list of ONE Gauss-Newton
step into the step's measured HBM traffic (total and per kernel).  usage: python tools/summarize_traffic.py launches.csv [out.json] > table.md"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ii, ki, ni, vi, ui, gi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}
agg = collections.OrderedDict()
launches = set()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * SC.get(r[ui], 1.0)
    key = (re.sub(r"^void ", "", re.sub(r"\(.*", "", r[ki])), r[gi])
    a = agg.setdefault(key, dict(n=set(), rd=0.0, wr=0.0, us=0.0))
    a["n"].add(r[ii]); launches.add(r[ii])
    if r[ni] == "dram__bytes_read.sum": a["rd"] += v
    elif r[ni] == "dram__bytes_write.sum": a["wr"] += v
    elif r[ni] == "gpu__time_duration.sum": a["us"] += v
tot_b = sum(a["rd"] + a["wr"] for a in agg.values()); tot_us = sum(a["us"] for a in agg.values())
print(f"# measured HBM traffic of one Gauss-Newton step: {sys.argv[1]}\n")
print("`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --cache-control none` over the profiled step "
      "(caches as the running step leaves them; launches serialised by the profiler, so compare time SHARES, not absolutes).\n")
print(f"{len(launches)} launches, {tot_b / 1e9:.3f} GB DRAM traffic (read {sum(a['rd'] for a in agg.values()) / 1e9:.3f} + write {sum(a['wr'] for a in agg.values()) / 1e9:.3f}), {tot_us / 1e3:.2f} ms of serialised kernel time\n")
print("| kernel | grid | launches | DRAM MB | share of bytes | MB / launch | total us | GB/s |\n|---|---|---:|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -(kv[1]["rd"] + kv[1]["wr"])):
    b = a["rd"] + a["wr"]
    print(f"| `{k[0]}` | {k[1]} | {len(a['n'])} | {b / 1e6:.1f} | {100 * b / tot_b:.1f}% | {b / 1e6 / len(a['n']):.2f} | {a['us']:.1f} | {b / max(a['us'], 1e-9) / 1e3:.0f} |")
if len(sys.argv) > 2:
    json.dump({"dram_bytes_per_step": tot_b, "launches": len(launches), "source": sys.argv[1]}, open(sys.argv[2], "w"))
