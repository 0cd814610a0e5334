import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from leak_det_gnn_b200 import ops
m = 2707456
g = torch.randn(m, 64, device="cuda"); x = torch.randn(m, 64, device="cuda")
for _ in range(3):
    ops.wgrad(g, x)
torch.cuda.synchronize()
