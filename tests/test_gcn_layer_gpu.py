"""The one-kernel GCN layer forward (csrc/gcn_layer.cu; BASELINE north_star "fused aggregate -> node Linear ->
activation", reference models/detector.py:198-201) against (a) the two-launch path it replaces -- required to be
BIT-IDENTICAL, train-mode dropout and live mask included -- and (b) the PyG restatement in fp64."""
import pytest
import torch

from conftest import rel_err
from leak_det_gnn_b200 import ops
from leak_det_gnn_b200.graph import batchify_edge_index
from oracle import pyg_restatement as pyg

pytestmark = pytest.mark.gpu


def _setup(graph_golden, net, bsz, k, d, seed=0):
    g0 = graph_golden(net)
    ei = torch.from_numpy(g0["edge_index"])
    n = len(g0["node_names"])
    graph = ops.PipeGraph(ei, n)
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(bsz, n, k, generator=gen)
    w = torch.randn(d, k, generator=gen) * 0.2
    b = torch.randn(d, generator=gen) * 0.1
    return graph, ei, n, x, w, b


@pytest.mark.parametrize("net,bsz,k,d", [("LTA", 1, 64, 64), ("LTA", 5, 64, 64), ("LT", 3, 64, 64), ("LTA", 300, 64, 64),
                                         ("LT", 200, 64, 64), ("LTA", 3, 128, 64), ("LTA", 7, 32, 32)])
def test_fused_layer_is_bit_identical_to_linear_then_aggregate(graph_golden, net, bsz, k, d):
    """300 windows x 2 slices over 148 CTAs: every CTA runs several windows (accumulator reuse, stage reuse, ring wrap)."""
    graph, ei, n, x, w, b = _setup(graph_golden, net, bsz, k, d)
    xc, wc, bc = x.cuda(), w.cuda(), b.cuda()
    assert ops.gcn_layer_supported(graph, xc, wc)
    for relu, p, seed in ((False, 0.0, 0), (True, 0.0, 0), (True, 0.1, 12345)):
        live_a = ops.new_live_mask(bsz, n, d, xc.device)
        live_b = ops.new_live_mask(bsz, n, d, xc.device)
        want = ops.spmm_fused(graph, ops.linear_tc(xc, wc), bias=bc, relu=relu, drop_p=p, drop_seed=seed, live_out=live_a)
        got = ops.gcn_layer_fwd(graph, xc, wc, bias=bc, relu=relu, drop_p=p, drop_seed=seed, live_out=live_b)
        assert torch.equal(got, want), (relu, p, (got - want).abs().max().item())
        assert torch.equal(live_a, live_b)
        assert torch.equal(ops.gcn_layer_fwd(graph, xc, wc, bias=bc, relu=relu, drop_p=p, drop_seed=seed), got)  # deterministic
    got = ops.gcn_layer_fwd(graph, xc, wc)                                   # no bias, no activation
    assert torch.equal(got, ops.spmm(graph, ops.linear_tc(xc, wc)))


@pytest.mark.parametrize("net,bsz", [("LTA", 6), ("LT", 160)])
def test_fused_layer_vs_pyg_restatement_fp64(graph_golden, net, bsz):
    graph, ei, n, x, w, b = _setup(graph_golden, net, bsz, 64, 64, seed=3)
    conv = pyg.GCNConv(64, 64).double()
    conv.load_state_dict({"lin.weight": w.double(), "bias": b.double()})
    want = torch.relu(conv(x.double().reshape(bsz * n, 64), batchify_edge_index(ei, n, bsz))).view(bsz, n, 64)
    got = ops.gcn_layer_fwd(graph, x.cuda(), w.cuda(), bias=b.cuda(), relu=True)
    # ReLU boundary: compare where the fp64 pre-activation is not within rounding of zero
    assert rel_err(got, want) <= 1e-5


def test_body_uses_fused_layer_and_matches_unfused(graph_golden, monkeypatch):
    """gnn_body with and without the fused layer: same forward bits, same gradients (the backward is shared)."""
    g0 = graph_golden("LTA")
    n = len(g0["node_names"])
    graph = ops.PipeGraph(torch.from_numpy(g0["edge_index"]), n)
    slot = torch.full((n,), -1, dtype=torch.int32)
    slot[torch.from_numpy(g0["sensor_node_idx"])] = torch.arange(29, dtype=torch.int32)
    gen = torch.Generator().manual_seed(1)
    h_s = torch.randn(9, 29, 64, generator=gen).cuda().requires_grad_(True)
    prm = [(torch.randn(s, generator=gen) * 0.2).cuda().requires_grad_(True) for s in ((64, 65), (64,), (64, 64), (64,), (64, 64), (64,))]
    outs = []
    for fused in (True, False):
        monkeypatch.setattr(ops, "FUSED_LAYER", fused)
        for t in [h_s] + prm:
            t.grad = None
        torch.manual_seed(5)
        y = ops.gnn_body(h_s, slot.cuda(), graph, 0.1, True, prm[0], prm[1], prm[2:])
        y.square().sum().backward()
        outs.append([y.detach().clone()] + [t.grad.clone() for t in [h_s] + prm])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
