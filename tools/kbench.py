#!/usr/bin/env python
"""Kernel micro-benchmarks (CUDA events, L2 flushed by working-set size): python tools/kbench.py [B]"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200 import lib as L  # noqa: E402
from leak_det_gnn_b200 import ops  # noqa: E402

PEAK = json.loads((REPO / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (REPO / "MEASURED_PEAKS.json").exists() else 6650.0


def timeit(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    z = np.load(REPO / "tests/golden/graph_LTA.npz")
    pg = ops.PipeGraph(torch.from_numpy(z["edge_index"]), 661)
    for d in (64, 128):
        x = torch.randn(b, 661, d, device="cuda")
        y = torch.empty_like(x)
        gb = 2 * x.numel() * 4 / 1e9
        for name, algo in (("staged", L.SPMM_STAGED), ("gather", L.SPMM_GATHER)):
            for tr in (False, True):
                ms = timeit(lambda: ops.spmm(pg, x, transpose=tr, algo=algo, out=y))
                print(f"spmm {name:6s} D={d:3d} B={b} transpose={int(tr)}: {ms:.4f} ms  {gb / ms * 1e3:7.1f} GB/s  "
                      f"{gb / ms * 1e3 / PEAK * 100:5.1f}% of {PEAK:.0f}")
        ms = timeit(lambda: y.copy_(x))
        print(f"torch copy  D={d:3d}: {ms:.4f} ms {gb / ms * 1e3:7.1f} GB/s")
    x = torch.randn(b * 661, 64, device="cuda")
    w = torch.randn(64, 64, device="cuda") * 0.1
    bias = torch.randn(64, device="cuda")
    gb = 2 * x.numel() * 4 / 1e9
    ms = timeit(lambda: ops.linear_tc(x, w, bias, True))
    print(f"linear_tc {x.shape[0]}x64x64: {ms:.4f} ms  {gb / ms * 1e3:7.1f} GB/s  {2 * x.shape[0] * 64 * 64 / ms / 1e9:.1f} TFLOP/s(fp32-equiv)")
    ms = timeit(lambda: torch.relu(torch.nn.functional.linear(x, w, bias)))
    print(f"torch linear+relu        : {ms:.4f} ms")


if __name__ == "__main__":
    main()
