import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from leak_det_gnn_b200 import ops
m = 2707456
g = torch.randn(m, 64, device="cuda"); x = torch.randn(m, 64, device="cuda")
for _ in range(3): ops.wgrad(g, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.wgrad(g, x)
e1.record(); torch.cuda.synchronize()
print(os.environ.get("LTGNN_TGRAD_DBG", "0"), f"{e0.elapsed_time(e1) / 10:.4f} ms")
