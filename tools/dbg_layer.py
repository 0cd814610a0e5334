import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from leak_det_gnn_b200 import ops
z = np.load(Path(__file__).resolve().parents[1] / "tests/golden/graph_LTA.npz")
n = len(z["node_names"]); graph = ops.PipeGraph(torch.from_numpy(z["edge_index"]), n)
torch.manual_seed(0)
x = torch.randn(1, n, 64, device="cuda"); w = torch.randn(64, 64, device="cuda") * 0.2
want = ops.spmm(graph, ops.linear_tc(x, w)); got = ops.gcn_layer_fwd(graph, x, w)
err = (got - want).abs()[0]
print("max err", err.max().item(), "rows wrong", (err.max(1).values > 1e-5).sum().item(), "cols wrong", (err.max(0).values > 1e-5).sum().item())
print("per col block of 8:", [round(err[:, c:c+8].max().item(), 3) for c in range(0, 64, 8)])
print("per row block of 32 (first 10):", [round(err[r:r+32].max().item(), 3) for r in range(0, 320, 32)])
# identity graph check: use W = I to see XW = X path
wi = torch.eye(64, device="cuda")
got2 = ops.gcn_layer_fwd(graph, x, wi); want2 = ops.spmm(graph, x.clone())
e2 = (got2 - want2).abs()[0]; print("W=I max err", e2.max().item(), [round(e2[:, c:c+8].max().item(), 3) for c in range(0, 64, 8)])
# identity graph: A_hat = I, so Y = X W^T exactly
g2 = ops.PipeGraph(torch.zeros(2, 0, dtype=torch.long), n)
for wname, ww in (("I", wi), ("rand", w)):
    got3 = ops.gcn_layer_fwd(g2, x, ww); want3 = ops.linear_tc(x, ww)
    e3 = (got3 - want3).abs()[0]
    print("A=I W=", wname, "max err", e3.max().item(), [round(e3[:, c:c+8].max().item(), 3) for c in range(0, 64, 8)],
          "rows", [round(e3[r:r+32].max().item(), 3) for r in range(0, 256, 32)])
print(got3[0, :3, :12]); print(want3[0, :3, :12])
wp = torch.zeros(64, 64, device="cuda")
for nn_ in range(64): wp[nn_, (nn_ + 8) % 64] = 1.0
got4 = ops.gcn_layer_fwd(g2, x, wp); want4 = ops.linear_tc(x, wp)
e4 = (got4 - want4).abs()[0]
print("perm W: per 8 cols", [round(e4[:, c:c+8].max().item(), 3) for c in range(0, 64, 8)])
# which source feature does each output column equal?
xs = x[0]
for c in (0, 8, 9, 16, 31, 32, 40):
    col = got4[0, :, c]
    match = [(k) for k in range(64) if torch.allclose(col, xs[:, k], atol=1e-5)]
    print("out col", c, "expected src", (c + 8) % 64, "matches src", match)
