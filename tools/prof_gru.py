import sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200 import ops
b = int(sys.argv[1]) if len(sys.argv) > 1 else 640
l = int(sys.argv[2]) if len(sys.argv) > 2 else 72
gru = torch.nn.GRU(10, 64, batch_first=True).cuda()
w = [getattr(gru, n).detach().clone().requires_grad_(True) for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
r = torch.randn(b, l, 29, device="cuda"); tf = torch.randn(b, l, 9, device="cuda")
dh = torch.randn(b, 29, 64, device="cuda")
for _ in range(2):
    for x in w: x.grad = None
    (ops.gru_encode(r, tf, *w) * dh).sum().backward()
torch.cuda.synchronize(); print("ok")
