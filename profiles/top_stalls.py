#!/usr/bin/env python
"""Stall summary of an ncu `--page source --csv` dump (first SASS section of the file):
    ncu -i X.ncu-rep --page source --csv --kernel-name regex:k_collide > src.csv
    python profiles/top_stalls.py src.csv [N]"""
import csv, sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
# sections start with a "Kernel Name" row followed by a header row; keep the first section
starts = [k for k, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = rows[starts[0]:(starts[1] if len(starts) > 1 else len(rows))]
print(sec[0][1][:100])
hdr = sec[1]
body = [r for r in sec[2:] if len(r) == len(hdr) and r[0] != "Address"]
ci, si, ie = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(float(r[si] or 0) for r in body)
print("total samples", tot, " SASS instructions", len(body), " warp instructions executed", sum(float(r[ie] or 0) for r in body))
for h in hdr:
    if h.startswith("stall_") and "Not Issued" not in h:
        s = sum(float(r[hdr.index(h)] or 0) for r in body)
        if s / tot > 0.005:
            print(f"  {h:24s} {100 * s / tot:5.1f} %")
lsb, ssb, mio = hdr.index("stall_long_sb"), hdr.index("stall_short_sb"), hdr.index("stall_mio")
for k, r in sorted(enumerate(body), key=lambda kr: -float(kr[1][si] or 0))[:n]:
    prev = body[k - 1][ci][:50] if k else ""
    print(f"{k:5d} {float(r[si]):8.0f} lsb={r[lsb]:>6} ssb={r[ssb]:>6} mio={r[mio]:>6}  {r[ci][:64]:64s} | prev: {prev}")
