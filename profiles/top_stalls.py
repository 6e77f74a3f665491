#!/usr/bin/env python
"""Top stall sites of an ncu `--page source --csv` dump (SASS view): python top_stalls.py src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
ci, si = hdr.index('Source'), hdr.index('# Samples')
lsb, ssb, mio = hdr.index('stall_long_sb'), hdr.index('stall_short_sb'), hdr.index('stall_mio')
body = [r for r in rows[2:] if len(r) > si]
tot = sum(float(r[si] or 0) for r in body)
print('total samples', tot)
idx = {id(r): k for k, r in enumerate(body)}
for r in sorted(body, key=lambda r: -float(r[si] or 0))[:n]:
    k = idx[id(r)]
    prev = body[k - 1][ci][:60] if k else ''
    print(f"{k:5d} {float(r[si]):8.0f} lsb={r[lsb]:>6} ssb={r[ssb]:>6} mio={r[mio]:>6}  {r[ci][:70]:70s} | prev: {prev}")
