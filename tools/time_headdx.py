import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from leak_det_gnn_b200 import ops, instrument as inst
B, N, P = 4096, 661, 764
x = torch.randn(B, N, 64, device="cuda").relu().requires_grad_(True)
import numpy as np; z = np.load(Path(__file__).resolve().parents[1] / "tests/golden/graph_LTA.npz"); ends = torch.from_numpy(z["pipe_ends"]).to(device="cuda", dtype=torch.int32)
w1 = (torch.randn(128, 192, device="cuda") * 0.1).requires_grad_(True)
b1 = (torch.randn(128, device="cuda") * 0.1).requires_grad_(True)
w2 = (torch.randn(1, 128, device="cuda") * 0.2).requires_grad_(True)
inc = ops.pipe_incidence(ends, N)
for it in range(4):
    if it == 1:
        torch.cuda.synchronize(); inst.reset(timing=True)
    part, pooled = ops.heads(x, ends, w1, b1, w2, 0.1, True, inc)
    (part.sum() + pooled.sum()).backward()
torch.cuda.synchronize()
print(os.environ.get("LTGNN_DBG", "0"), {k: round(v["mean_ms"], 3) for k, v in inst.summary().items()})
