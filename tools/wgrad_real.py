"""Weight-gradient accuracy on REAL (ill-conditioned) operands: the conv-0 and pipe-head gradients of a CE step."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch
import test_benchscale_gpu as T
from leak_det_gnn_b200 import ops

def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()

ours, o32, o64, g = T._models("LTA", 764)
residual, tfeat, label = T._inputs(128, 36, 765)
ours.eval()
cap = {}
orig = ops.wgrad
def spy(gz, x):
    cap.setdefault("calls", []).append((gz.detach().clone(), x.detach().clone()))
    return orig(gz, x)
ops.wgrad = spy
lo = ours(residual.cuda(), tfeat.cuda())
torch.nn.functional.cross_entropy(lo, label.cuda()).backward()
ops.wgrad = orig
for i, (gz, x) in enumerate(cap["calls"]):
    gz2, x2 = gz.reshape(-1, gz.shape[-1]), x.reshape(-1, x.shape[-1])
    want = gz2.double().t() @ x2.double()
    cond = (gz2.double().abs().t() @ x2.double().abs()).max() / want.abs().max()
    line = f"call {i}: M={gz2.shape[0]} cond~{cond.item():.1f}  torch fp32 {rel(gz2.t() @ x2, want):.2e}"
    line += f"  tgrad {rel(orig(gz2, x2), want):.2e}"
    print(line, flush=True)
