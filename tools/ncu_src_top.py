#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv --print-source sass` export (optionally gzipped).
   python tools/ncu_src_top.py <src.csv[.gz]> [n]"""
import csv, gzip, sys
path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
f = gzip.open(path, "rt", errors="ignore") if path.endswith(".gz") else open(path, errors="ignore")
rows = list(csv.reader(f))
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
h = rows[hi]
idx = {k: i for i, k in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
S = lambda r: int(r[idx["# Samples"]] or 0)
E = lambda r: int(r[idx["Instructions Executed"]] or 0)
print("samples", sum(map(S, data)), "warp-instructions", sum(map(E, data)), "sass lines", len(data))
for r in sorted(data, key=lambda r: -S(r))[:n]:
    print(f"{r[idx['Address']][-5:]} {S(r):7d} {E(r):10d}  {r[idx['Source']][:100]}")
