"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
instructions executed, stall samples, shared-memory wavefronts (ideal / excessive). Usage: ncu_lines.py dump.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr, cur_line = None, None, None
agg = defaultdict(lambda: defaultdict(float))
text = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    d = dict(zip(hdr[:2] + ["Address", "SASS"] + hdr[4:], r))
    if r[0] not in ("", "-"):
        cur_line = (cur_file, int(r[0]))
        text[cur_line] = r[1].strip()[:90]
        continue  # line summary row: the SASS rows below carry the same counts
    if cur_line is None or r[2] in ("-", "..."):
        continue
    for k in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive", "L1 Wavefronts Shared Ideal"):
        try:
            agg[cur_line][k] += float(d.get(k, 0) or 0)
        except ValueError:
            pass
tot = defaultdict(float)
for v in agg.values():
    for k, x in v.items():
        tot[k] += x
print("totals:", dict(tot))
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>7s} {'smem wf%':>8s} {'excess':>9s}  source")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print(f"{key[0][:20]}:{key[1]:<6d} {100*v['# Samples']/max(tot['# Samples'],1):8.2f} {100*v['Instructions Executed']/max(tot['Instructions Executed'],1):7.2f} "
          f"{100*v['L1 Wavefronts Shared']/max(tot['L1 Wavefronts Shared'],1):8.2f} {v['L1 Wavefronts Shared Excessive']:9.0f}  {text.get(key,'')}")
