"""Error of the tensor-core weight gradient vs fp64 as the row count grows (accumulator length per CTA)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from leak_det_gnn_b200 import ops

def rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max()).item()

torch.manual_seed(0)
for m in (1000, 10000, 84608, 338432, 1000000, 2707456):
    for kind in ("randn", "relu"):
        g = torch.randn(m, 64, device="cuda")
        x = torch.randn(m, 64, device="cuda")
        if kind == "relu":
            x = x.relu()
            g = g * (torch.rand_like(g) > 0.5)
        want = g.double().t() @ x.double()
        got = ops.wgrad(g, x)
        t32 = g.t() @ x
        print(f"M={m:8d} {kind:6s} tgrad {rel(got, want):.2e}   torch fp32 matmul {rel(t32, want):.2e}", flush=True)
