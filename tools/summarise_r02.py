#!/usr/bin/env python
"""Round-2 profile summaries.
   python tools/summarise_r02.py launches <launches.csv> <out.csv>     per-kernel totals of an ncu launch list
   python tools/summarise_r02.py full <prof.ncu-rep | raw.csv[.gz]> <out.csv> [traffic.json key]   --set full metrics per kernel"""
import collections
import csv
import json
import re
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def short(name: str) -> str:
    name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("ltgnn::", "")
    m = re.search(r"tgrad_kernel<(.+?), (.+?), (?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(-?\d+)>", name)
    if m:
        return f"tgrad<{m.group(1).split('::')[-1]},{m.group(2).split('::')[-1]},mt={m.group(3)},seg={m.group(6)}>"
    m = re.search(r"spmm_staged_kernel<(?:\(bool\))?(\d), (?:\(bool\))?(\d)>", name)
    if m:
        return f"spmm_staged_kernel<epi={m.group(1)},gate={m.group(2)}>"
    m = re.search(r"([A-Za-z0-9_]+_kernel)", name)
    if m and "at::" not in name and "elementwise" not in name:
        return m.group(1)
    return "torch/lib: " + name.split("(")[0][-70:]


def launches(path: str, out: str) -> None:
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        t = float(r[vi].replace(",", ""))
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += t
    unit = next((r[hdr.index("Metric Unit")] for r in rows if r is not hdr and len(r) > vi and r[mi] == "gpu__time_duration.sum"), "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    total = sum(v[1] for v in agg.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "mean_ms", "share_of_all_launches"])
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, n, round(t * scale, 4), round(t * scale / n, 4), round(t / total, 4)])
    print(f"{out}: {len(agg)} kernels, {total * scale:.2f} ms of GPU time in the capture")


WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wavefronts_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__sass_inst_executed_op_local_ld.sum": "local_loads",
}


def full(rep: str, out: str, key: str = "") -> None:
    if rep.endswith(".gz"):           # `ncu -i x.ncu-rep --page raw --csv | gzip` exported on the GPU box
        import gzip
        raw = gzip.open(rep, "rt", errors="ignore").read()
    elif rep.endswith(".csv"):
        raw = open(rep, errors="ignore").read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    cols = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    for r in rows[2:]:
        k = short(r[ki])
        d = per.setdefault(k, collections.defaultdict(list))
        for m, nm in WANT.items():
            if m in cols and r[cols[m]] not in ("", "n/a"):
                d[nm].append((float(r[cols[m]].replace(",", "")), units[cols[m]]))
    to_bytes = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    to_ms = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "s": 1e3, "second": 1e3}
    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "mean_ms", "dram_read_GB", "dram_write_GB", "dram_GBs", "dram_pct_of_ncu_peak",
                    "tensor_pipe_pct", "issue_active_pct", "l1_wavefronts_pct", "l2_hit_pct", "registers", "grid", "block",
                    "local_loads"])
        for k, d in per.items():
            def mean(nm, conv=None):
                v = d.get(nm)
                if not v:
                    return None
                return sum(x * (conv.get(u, 1.0) if conv else 1.0) for x, u in v) / len(v)
            ms, rd, wr = mean("duration", to_ms), mean("dram_read", to_bytes), mean("dram_write", to_bytes)
            gbs = (rd + wr) / (ms * 1e-3) / 1e9 if ms and rd is not None else None
            traffic[k] = None if rd is None else rd + wr
            fmt = lambda x, n=3: "" if x is None else round(x, n)
            w.writerow([k, len(d["duration"]), fmt(ms, 4), fmt(rd and rd / 1e9), fmt(wr and wr / 1e9), fmt(gbs, 0),
                        fmt(mean("dram_pct"), 1), fmt(mean("tensor_pipe_pct"), 1), fmt(mean("issue_active_pct"), 1),
                        fmt(mean("l1_wavefronts_pct"), 1), fmt(mean("l2_hit_pct"), 1), fmt(mean("registers"), 0),
                        fmt(mean("grid"), 0), fmt(mean("block"), 0), fmt(mean("local_loads"), 0)])
    print(f"{out}: {len(per)} kernels")
    if key:
        tp = REPO / "profiles" / "ncu_traffic_r02.json"
        data = json.loads(tp.read_text()) if tp.exists() else {}
        data.setdefault(key, {}).update({k: v for k, v in traffic.items() if v is not None})
        tp.write_text(json.dumps(data, indent=1))
        # the per-op table bench.py reads (profiles/ncu_traffic.json): kernel -> the instrument name of the op it implements
        op_of = {"gcn_layer_fwd_kernel": "gcn_layer_fwd", "pipe_head_fwd_kernel": "pipe_head_fwd", "mean_pool_kernel": "mean_pool_fwd",
                 "pipe_head_bwd_dx_kernel": "pipe_head_bwd_dx", "tgrad<HeadLive,HeadFeatSlice,mt=1,seg=-2>": "pipe_head_bwd_w",
                 "tgrad<StackedRows,StackedRows,mt=1,seg=-1>": "wgrad_tc", "linear_ts_kernel": "linear_tc",
                 "spmm_staged_kernel<epi=0,gate=1>": "spmm_fused_bwd", "spmm_staged_kernel<epi=1,gate=0>": "spmm_fused_fwd",
                 "node_init_fwd_kernel": "node_init_fwd", "gate_extract_kernel": "node_init_bwd"}
        bp = REPO / "profiles" / "ncu_traffic.json"
        table = json.loads(bp.read_text()) if bp.exists() else {}
        ops = table.setdefault(key, {})
        for k, v in traffic.items():
            name = "pipe_head_bwd_w" if k.startswith("tgrad<HeadLive,HeadFeatSlice") else op_of.get(k)
            if v is not None and name:
                ops[name] = v
        table["_note_r02"] = ("GNN-stack entries refreshed from profiles/r02*_kernels_ncu_full.csv (ncu --set full, per launch); "
                              "node_init_bwd = its streaming pass (gate_extract_kernel) only; GRU entries are round 1's (kernels unchanged)")
        bp.write_text(json.dumps(table, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
