"""Per-phase table of an `ncu --page source --csv --print-source cuda,sass` dump of mfHexPlanesKernel: share of the stall samples,
of the executed instructions, static SASS instructions (code size) and the top stall reasons.
Usage: ncu_phase_table.py dump.csv [name:first_line ...]   (phase marks are read from the current source unless given)"""
import csv
import re
import sys
from collections import defaultdict

SRC = __import__("os").path.join(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))),
                                 "l3ster_b200", "csrc", "mf_hex_planes.cuh")
marks = []
for i, line in enumerate(open(SRC), 1):
    m = re.search(r"// ---- ([A-E]):", line)
    if m:
        marks.append((i, m.group(1)))
    elif "for (int it = 0; it < n_it" in line:
        marks.append((i, "loop head"))
marks.append((10**9, "end"))
if len(sys.argv) > 2:  # explicit marks for a capture of an older revision of the source: name:first_line ...
    marks = [(int(a.split(":")[1]), a.split(":")[0]) for a in sys.argv[2:]] + [(10**9, "end")]
rows = list(csv.reader(open(sys.argv[1])))
hdr = cur = curfile = None
agg, static = defaultdict(lambda: defaultdict(float)), defaultdict(int)
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
    name = "other: " + cur[0]
    if cur[0].startswith("mf_hex_planes"):
        name = "prologue"
        for (lo, n), (hi, _) in zip(marks, marks[1:]):
            if lo <= cur[1] < hi:
                name = n
    static[name] += 1
    for k, v in d.items():
        if k.startswith("stall_") or k in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared"):
            try:
                agg[name][k] += float(v or 0)
            except ValueError:
                pass
tot = sum(v["# Samples"] for v in agg.values())
toti = sum(v["Instructions Executed"] for v in agg.values())
print(f"samples {tot:.0f}, warp instructions {toti:.0f}, static SASS instructions {sum(static.values())} ({16 * sum(static.values()) / 1024:.0f} kB)")
for n, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    st = sorted(((k, x) for k, x in v.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:5]
    print(f"{n:24s} samples {100 * v['# Samples'] / tot:5.1f} %  instructions {100 * v['Instructions Executed'] / toti:5.1f} %  static {static[n]:5d}  "
          + ", ".join(f"{k[6:]} {100 * x / max(v['# Samples'], 1):.0f} %" for k, x in st))
