#!/usr/bin/env python
"""Head-kernel timing matrix: python tools/hbench.py [B]"""
import sys
from pathlib import Path
import numpy as np, torch
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200 import ops  # noqa: E402
from tools.kbench import timeit  # noqa: E402

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = np.load(REPO / "tests/golden/graph_LTA.npz")
ends = torch.from_numpy(g["pipe_ends"]).to(torch.int32).cuda()
x = torch.randn(b, 661, 64, device="cuda").relu()
w1 = (torch.randn(128, 192) * 0.1).cuda(); b1 = (torch.randn(128) * 0.1).cuda(); w2 = (torch.randn(1, 128) * 0.1).cuda()
for P in (764, 64):
    e = ends[:P].contiguous()
    xg = x.clone().requires_grad_(True)
    with torch.no_grad():
        print(f"P={P} fwd eval no-save : {timeit(lambda: ops.heads(x, e, w1, b1, w2, 0.1, False)):.3f} ms")
        print(f"P={P} fwd train no-save: {timeit(lambda: ops.heads(x, e, w1, b1, w2, 0.1, True)):.3f} ms")
    print(f"P={P} fwd eval save    : {timeit(lambda: ops.heads(xg, e, w1, b1, w2, 0.1, False)):.3f} ms")
    print(f"P={P} fwd train save   : {timeit(lambda: ops.heads(xg, e, w1, b1, w2, 0.1, True)):.3f} ms")

# backward pieces
P = 764
e = ends[:P].contiguous()
xg = x.clone().requires_grad_(True)
w1g = w1.clone().requires_grad_(True); b1g = b1.clone().requires_grad_(True); w2g = w2.clone().requires_grad_(True)
part, pooled = ops.heads(xg, e, w1g, b1g, w2g, 0.1, True)
loss = part.sum() + pooled.sum()
from leak_det_gnn_b200 import instrument as inst
inst.reset(timing=True)
for _ in range(5):
    loss.backward(retain_graph=True)
torch.cuda.synchronize()
print({k: round(v["mean_ms"], 3) for k, v in inst.summary().items()})
