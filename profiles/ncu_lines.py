#!/usr/bin/env python
"""Hot source lines of an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` export:
stall samples and executed warp-instructions per CUDA-C line (rows with Address == '-')."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and cur and len(r) == len(hdr) and r[2] == "-":
        try:
            out.append((int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), cur, int(r[0]), r[1].strip()[:100]))
        except ValueError:
            pass
ts, ti = sum(o[0] for o in out) or 1, sum(o[1] for o in out) or 1
print(f"total samples {ts}, warp-instructions {ti}")
byfile = {}
for o in out:
    a = byfile.setdefault(o[2], [0, 0]); a[0] += o[0]; a[1] += o[1]
for f, (s, i) in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:24s} samples {100 * s / ts:5.1f} %  instructions {100 * i / ti:5.1f} %")
for o in sorted(out, reverse=True)[:top]:
    print(f"{100 * o[0] / ts:5.1f}% smp {100 * o[1] / ti:5.1f}% ins  {o[2]}:{o[3]}  {o[4]}")
