import sys
from pathlib import Path
import torch
REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200 import ops
from tools.kbench import timeit
b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
l = int(sys.argv[2]) if len(sys.argv) > 2 else 288
gru = torch.nn.GRU(10, 64, batch_first=True).cuda()
w = [getattr(gru, n).detach() for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")]
r = torch.randn(b, l, 29, device="cuda"); tf = torch.randn(b, l, 9, device="cuda")
print(f"gru_fwd B={b} L={l} no-save: {timeit(lambda: ops.gru_fwd(r, tf, *w), iters=5):.2f} ms")
print(f"gru_fwd B={b} L={l} save   : {timeit(lambda: ops.gru_fwd(r, tf, *w, save=True), iters=5):.2f} ms")

w = [x.clone().requires_grad_(True) for x in w]
dh = torch.randn(b, 29, 64, device="cuda")
def step():
    for x in w: x.grad = None
    (ops.gru_encode(r, tf, *w) * dh).sum().backward()
from leak_det_gnn_b200 import instrument as inst
step(); torch.cuda.synchronize(); inst.reset(timing=True)
for _ in range(3): step()
torch.cuda.synchronize()
print({k: round(v["mean_ms"], 2) for k, v in inst.summary().items()}, f"peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
