"""Operator-level drop-ins (leak_det_gnn_b200.nn) vs the PyG restatement (oracle), forward and backward."""
import pytest
import torch

from conftest import rel_err
from leak_det_gnn_b200 import nn as lnn
from leak_det_gnn_b200.graph import batchify_edge_index
from oracle import pyg_restatement as pyg

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _pair(din, dout, seed=0):
    torch.manual_seed(seed)
    ours = lnn.GCNConv(din, dout)
    torch.manual_seed(seed)
    ref = pyg.GCNConv(din, dout)
    assert list(ours.state_dict()) == list(ref.state_dict()) == ["bias", "lin.weight"]
    assert torch.equal(ours.lin.weight, ref.lin.weight)  # same seeded init stream (glorot drawn twice)
    with torch.no_grad():
        b = torch.randn(dout) * 0.1
        ours.bias.copy_(b)
        ref.bias.copy_(b)
    return ours.cuda(), ref.double()


@pytest.mark.parametrize("net,bsz,din,dout", [("LTA", 4, 64, 64), ("LT", 3, 64, 64), ("LTA", 2, 128, 128),
                                              ("LTA", 3, 64, 128), ("LTA", 2, 40, 24)])
def test_gcnconv_fast_path(graph_golden, net, bsz, din, dout):
    g0 = graph_golden(net)
    ei = torch.from_numpy(g0["edge_index"])
    n = len(g0["node_names"])
    ours, ref = _pair(din, dout)
    graph = lnn.PipeGraph(ei, n)
    gen = torch.Generator().manual_seed(3)
    x = torch.randn(bsz, n, din, generator=gen)
    dy = torch.randn(bsz, n, dout, generator=gen)
    xr = x.double().reshape(bsz * n, din).requires_grad_(True)
    yr = ref(xr, batchify_edge_index(ei, n, bsz))
    yr.backward(dy.double().reshape(bsz * n, dout))
    xo = x.cuda().requires_grad_(True)
    yo = ours(xo, graph)
    assert yo.shape == (bsz, n, dout)
    yo.backward(dy.cuda())
    assert rel_err(yo.reshape(bsz * n, dout), yr) <= TOL
    assert rel_err(xo.grad.reshape(bsz * n, din), xr.grad) <= TOL
    assert rel_err(ours.lin.weight.grad, ref.lin.weight.grad) <= TOL
    assert rel_err(ours.bias.grad, ref.bias.grad) <= TOL


def test_gcnconv_pyg_signature(graph_golden):
    """conv(x[B*N, D], edge_index[2, B*E]) exactly as the reference calls PyG (detector.py:199)."""
    g0 = graph_golden("LTA")
    ei = torch.from_numpy(g0["edge_index"])
    n, bsz = 661, 3
    ours, ref = _pair(64, 64, seed=1)
    big = batchify_edge_index(ei, n, bsz)
    x = torch.randn(bsz * n, 64, generator=torch.Generator().manual_seed(4))
    yr = ref(x.double(), big)
    yo = ours(x.cuda(), big.cuda())
    assert yo.shape == (bsz * n, 64) and rel_err(yo, yr) <= TOL
    assert rel_err(ours(x.cuda(), big.cuda()), yr) <= TOL  # cached graph


def test_global_mean_pool():
    x = torch.randn(5 * 661, 64, device="cuda")
    batch = torch.arange(5, device="cuda").repeat_interleave(661)
    want = pyg.global_mean_pool(x.cpu().double(), batch.cpu())
    assert rel_err(lnn.global_mean_pool(x, batch), want) <= TOL
    assert rel_err(lnn.global_mean_pool(x, batch, size=5), want) <= TOL
