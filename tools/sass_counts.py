"""Per-kernel counts of the Blackwell-specific SASS instructions in the in-tree libltgnn.so, so that the tcgen05 / TMA /
tensor-memory claims can be audited without rebuilding:  python tools/sass_counts.py > profiles/r02_sass_counts.txt"""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

LIB = Path(__file__).resolve().parents[1] / "leak_det_gnn_b200" / "libltgnn.so"
WATCH = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "SYNCS", "LDGSTS", "ATOM", "RED", "LDL", "STL", "HMMA", "FFMA"]


def main() -> None:
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    kernels: "OrderedDict[str, Counter]" = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("ltgnn::", "").replace("void ", "")
            cur = kernels.setdefault(name[:110], Counter())
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            cur["total"] += 1
            if op in WATCH:
                cur[op] += 1
    print(f"# {LIB.name}: SASS instruction counts per kernel (cuobjdump -sass); UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, UBLKCP = TMA 1-D bulk copy,")
    print("# LDTM / STTM = tcgen05.ld / st (tensor memory), UTCBAR = tcgen05.commit, LDGSTS = cp.async, ATOM / RED = global atomics,")
    print("# LDL / STL = register spills")
    print(f"{'kernel':112s} {'total':>7s} " + " ".join(f"{w:>7s}" for w in WATCH))
    tot = Counter()
    for k, c in kernels.items():
        print(f"{k:112s} {c['total']:7d} " + " ".join(f"{c[w]:7d}" for w in WATCH))
        tot.update(c)
    print(f"{'ALL KERNELS':112s} {tot['total']:7d} " + " ".join(f"{tot[w]:7d}" for w in WATCH))


if __name__ == "__main__":
    sys.exit(main())
