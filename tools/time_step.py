#!/usr/bin/env python
"""Per-kernel times of the GNN-stack step at bench shapes (instrument timing, no ncu):  python tools/time_step.py [B] [iters]"""
import sys
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200 import instrument as inst  # noqa: E402
from leak_det_gnn_b200.models import LeakDetector  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
g = np.load(REPO / "tests/golden/graph_LTA.npz")
torch.manual_seed(42)
m = LeakDetector(REPO / "tests/golden/L-TOWN-A.topo.inp", [str(s) for s in g["sensor_node_ids"]],
                 [str(p) for p in g["pipe_ids"]]).cuda().train()
h_s = torch.randn(b, 29, 64, device="cuda", requires_grad=True)
label = torch.randint(0, 765, (b,), device="cuda")
for it in range(iters + 2):
    if it == 2:
        torch.cuda.synchronize()
        inst.reset(timing=True)
    m.zero_grad(set_to_none=True)
    h_s.grad = None
    torch.nn.functional.cross_entropy(m.gnn_stack(h_s), label).backward()
torch.cuda.synchronize()
s = inst.summary()
print({k: round(v["mean_ms"], 4) for k, v in s.items()}, "sum", round(sum(v["total_ms"] for v in s.values()) / iters, 3))
