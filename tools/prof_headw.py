import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from leak_det_gnn_b200 import ops
B, N, P = 4096, 661, 764
x = torch.randn(B, N, 64, device="cuda").relu().requires_grad_(True)
ends = torch.randint(0, N, (P, 2), device="cuda", dtype=torch.int32)
w1 = (torch.randn(128, 192, device="cuda") * 0.1).requires_grad_(True)
b1 = (torch.randn(128, device="cuda") * 0.1).requires_grad_(True)
w2 = (torch.randn(1, 128, device="cuda") * 0.2).requires_grad_(True)
for _ in range(2):
    part, pooled = ops.heads(x, ends, w1, b1, w2, 0.1, True)
    (part.sum() + pooled.sum()).backward()
torch.cuda.synchronize()
