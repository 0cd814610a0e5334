#!/usr/bin/env python
"""Stall-sample summary of one kernel from an `ncu --set full --import-source on` capture exported with
`ncu -i x.ncu-rep --page source --csv --print-source sass [| gzip]`:
   python tools/ncu_src_top.py <src.csv[.gz]> [n] [note]
Prints (1) the n SASS instructions with the most warp-stall samples and (2) the samples grouped by how often the
instruction was executed -- in a warp-specialised kernel every role (loader / MMA / epilogue / gather ...) runs its loop a
different number of times, so the groups separate the roles: busy instructions of a role share one count, its barrier
polls show up as a few lines with a very large count."""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
note = sys.argv[3] if len(sys.argv) > 3 else ""
f = gzip.open(path, "rt", errors="ignore") if path.endswith(".gz") else open(path, errors="ignore")
rows = list(csv.reader(f))
kern = next((r[1] for r in rows if r and r[0] == "Kernel Name"), "")
hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
h = rows[hi]
idx = {k: i for i, k in enumerate(h)}
nxt = next((i for i in range(hi + 1, len(rows)) if "Address" in rows[i] and "Source" in rows[i]), len(rows))
data = [r for r in rows[hi + 1:nxt] if len(r) == len(h)]   # an export of several launches repeats the header: first one
S = lambda r: int(r[idx["# Samples"]] or 0)
E = lambda r: int(r[idx["Instructions Executed"]] or 0)
if note:
    print(note)
print("kernel:", kern[:160])
print("warp-stall samples", sum(map(S, data)), "| warp-instructions executed", sum(map(E, data)), "| SASS lines", len(data))
print("\naddress  samples   executed    instruction")
for r in sorted(data, key=lambda r: -S(r))[:n]:
    print(f"{r[idx['Address']][-5:]} {S(r):8d} {E(r):10d}  {r[idx['Source']][:100]}")
b = collections.defaultdict(lambda: [0, 0])
for r in data:
    b[E(r)][0] += 1
    b[E(r)][1] += S(r)
print("\nsamples by execution count (= by warp role / loop):  executed  lines  samples")
for e, (k, s) in sorted(b.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{e:12d} {k:6d} {s:8d}")
