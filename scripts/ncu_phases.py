"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by line ranges of one source file.
Usage: ncu_phases.py dump.csv file_prefix name:lo-hi [name:lo-hi ...]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
prefix = sys.argv[2]
bounds = []
for spec in sys.argv[3:]:
    n, r = spec.split(":")
    lo, hi = r.split("-")
    bounds.append((int(lo), int(hi), n))
hdr = cur = curfile = None
agg = defaultdict(lambda: defaultdict(float))
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        curfile = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0] not in ("", "-"):
        try:
            cur = (curfile, int(r[0]))
        except ValueError:
            cur = None
        continue
    if cur is None or len(r) < len(hdr) - 5:
        continue
    d = dict(zip(hdr[:2] + ["Address", "SASS"] + hdr[4:], r))
    for k, v in d.items():
        if k.startswith("stall_") or k in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared"):
            try:
                agg[cur][k] += float(v or 0)
            except ValueError:
                pass
ph = defaultdict(lambda: defaultdict(float))
for (f, l), v in agg.items():
    name = "other:" + f
    if f.startswith(prefix):
        for lo, hi, n in bounds:
            if lo <= l <= hi:
                name = n
    for k, x in v.items():
        ph[name][k] += x
tot = sum(v["# Samples"] for v in ph.values())
toti = sum(v["Instructions Executed"] for v in ph.values())
totw = sum(v["L1 Wavefronts Shared"] for v in ph.values())
for n, v in sorted(ph.items(), key=lambda kv: -kv[1]["# Samples"]):
    st = sorted(((k, x) for k, x in v.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:6]
    print(f"{n:28s} samples {100*v['# Samples']/tot:5.1f}% inst {100*v['Instructions Executed']/toti:5.1f}% smem {100*v['L1 Wavefronts Shared']/max(totw,1):5.1f}%  ",
          ", ".join(f"{k[6:]}={100*x/max(v['# Samples'],1):.0f}%" for k, x in st))
