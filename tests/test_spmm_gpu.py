"""Aggregation kernels vs the C oracle through the C ABI: BIT-EXACT (same per-entry mul, add and
summation order as a sequential scatter-add over PyG's edge list)."""
import numpy as np
import pytest
import torch

from leak_det_gnn_b200 import lib as L
from leak_det_gnn_b200.ops import PipeGraph, aggregate, spmm
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _graph(graph_golden, net):
    g0 = graph_golden(net)
    return PipeGraph(torch.from_numpy(g0["edge_index"]), len(g0["node_names"]))


def _bits(t):
    return t.detach().cpu().numpy().view(np.int32)


@pytest.mark.parametrize("net", ["LTA", "LT"])
@pytest.mark.parametrize("algo", [L.SPMM_STAGED, L.SPMM_GATHER, L.SPMM_AUTO])
@pytest.mark.parametrize("b,d", [(1, 64), (5, 32), (3, 128), (301, 64)])
def test_spmm_bit_exact(graph_golden, net, algo, b, d):
    pg = _graph(graph_golden, net)
    c = pg.csr
    gen = torch.Generator().manual_seed(1000 + b + d)
    x = torch.randn(b, pg.num_nodes, d, generator=gen)
    for transpose, arrs in ((False, (c.rowptr, c.col, c.val)), (True, (c.t_rowptr, c.t_col, c.t_val))):
        want = c_oracle.spmm(*arrs, x.numpy())
        got = spmm(pg, x.cuda(), transpose=transpose, algo=algo)
        torch.cuda.synchronize()
        assert np.array_equal(_bits(got), want.view(np.int32)), (net, algo, b, d, transpose)


@pytest.mark.parametrize("d", [4, 12, 48])
def test_spmm_gather_odd_widths(graph_golden, d):
    pg = _graph(graph_golden, "LTA")
    c = pg.csr
    x = torch.randn(7, pg.num_nodes, d, generator=torch.Generator().manual_seed(d))
    got = spmm(pg, x.cuda())
    assert np.array_equal(_bits(got), c_oracle.spmm(c.rowptr, c.col, c.val, x.numpy()).view(np.int32))
    with pytest.raises(L.LtgnnError, match="STAGED"):
        spmm(pg, x.cuda(), algo=L.SPMM_STAGED)


def test_spmm_random_graphs_and_edge_cases():
    rng = np.random.default_rng(7)
    for n, e in ((1, 0), (2, 1), (37, 90), (300, 5), (1000, 2300)):
        src = rng.integers(0, n, e)
        dst = rng.integers(0, n, e)
        pg = PipeGraph(torch.from_numpy(np.stack([src, dst])), n)  # directed, duplicates, self loops, isolated nodes
        c = pg.csr
        x = torch.randn(4, n, 64, generator=torch.Generator().manual_seed(n))
        for algo in (L.SPMM_STAGED, L.SPMM_GATHER):
            for tr, arrs in ((False, (c.rowptr, c.col, c.val)), (True, (c.t_rowptr, c.t_col, c.t_val))):
                got = spmm(pg, x.cuda(), transpose=tr, algo=algo)
                assert np.array_equal(_bits(got), c_oracle.spmm(*arrs, x.numpy()).view(np.int32)), (n, e, algo, tr)
    # empty batch is a no-op
    out = spmm(pg, torch.empty(0, 1000, 64, device="cuda"))
    assert out.shape == (0, 1000, 64)
    # argument errors surface as exceptions, never as a silent fallback
    with pytest.raises(ValueError, match="CUDA"):
        spmm(pg, torch.zeros(1, 1000, 64))
    with pytest.raises(ValueError, match="float32"):
        spmm(pg, torch.zeros(1, 1000, 64, device="cuda", dtype=torch.float64))
    with pytest.raises(ValueError, match="whole number"):
        spmm(pg, torch.zeros(1, 999, 64, device="cuda"))
    with pytest.raises(L.LtgnnError, match="multiple of 4"):
        spmm(pg, torch.zeros(1, 1000, 6, device="cuda"))


def test_large_graph_gather(graph_golden):
    """Scaled network of BASELINE config 5 shape (smaller): only the L2-gather kernel fits."""
    rng = np.random.default_rng(198)
    n = 20000
    par = np.arange(1, n) - 1 - rng.integers(0, np.minimum(np.arange(1, n), 64))
    src = np.concatenate([np.arange(1, n), par])
    dst = np.concatenate([par, np.arange(1, n)])
    pg = PipeGraph(torch.from_numpy(np.stack([src, dst])), n)
    c = pg.csr
    x = torch.randn(2, n, 128, generator=torch.Generator().manual_seed(3))
    got = spmm(pg, x.cuda())
    assert np.array_equal(_bits(got), c_oracle.spmm(c.rowptr, c.col, c.val, x.numpy()).view(np.int32))
    with pytest.raises(L.LtgnnError, match="STAGED"):
        spmm(pg, x.cuda(), algo=L.SPMM_STAGED)


def test_full_size_properties(graph_golden):
    """BASELINE config 3 size (4096 windows, L-TOWN-A, D=64): size-independent checks.
    Linearity is exact in structure: A(x) for window b depends on window b only, and the adjoint
    identity <A x, y> = <x, A^T y> ties the forward and backward kernels together."""
    pg = _graph(graph_golden, "LTA")
    b, n, d = 4096, 661, 64
    gen = torch.Generator(device="cuda").manual_seed(198)
    x = torch.randn(b, n, d, device="cuda", generator=gen)
    y = torch.randn(b, n, d, device="cuda", generator=gen)
    ax = spmm(pg, x)
    aty = spmm(pg, y, transpose=True)
    lhs = (ax.double() * y.double()).sum()
    rhs = (x.double() * aty.double()).sum()
    assert abs(lhs - rhs) <= 1e-9 * abs(lhs) + 1e-3
    # window independence + determinism: a permuted batch gives the permuted result, bitwise
    perm = torch.randperm(b, device="cuda", generator=gen)
    assert torch.equal(spmm(pg, x[perm].contiguous()), ax[perm])
    # staged and gather kernels agree bitwise at full size
    assert torch.equal(spmm(pg, x, algo=L.SPMM_GATHER), ax)
    # spot check 8 windows against the oracle
    c = pg.csr
    sel = [0, 1, 147, 148, 2047, 4000, 4094, 4095]
    want = c_oracle.spmm(c.rowptr, c.col, c.val, x[sel].cpu().numpy())
    assert np.array_equal(_bits(ax[sel]), want.view(np.int32))


def test_aggregate_autograd(graph_golden):
    pg = _graph(graph_golden, "LTA")
    x = torch.randn(3, 661, 64, device="cuda", requires_grad=True)
    w = torch.randn(3, 661, 64, device="cuda")
    (aggregate(x, pg) * w).sum().backward()
    assert torch.equal(x.grad, spmm(pg, w, transpose=True))


@pytest.mark.parametrize("net,b,d", [("LTA", 5, 64), ("LT", 3, 64), ("LTA", 300, 128), ("LTA", 2, 32)])
def test_spmm_fused_epilogue_and_gate_bit_exact(graph_golden, net, b, d):
    from leak_det_gnn_b200.ops import spmm_fused
    pg = _graph(graph_golden, net)
    c = pg.csr
    gen = torch.Generator().manual_seed(77 + b)
    x = torch.randn(b, pg.num_nodes, d, generator=gen)
    bias = torch.randn(d, generator=gen)
    # forward epilogue: relu(A x + bias), same roundings as the oracle
    want = np.maximum(c_oracle.spmm(c.rowptr, c.col, c.val, x.numpy()) + bias.numpy(), 0.0).astype(np.float32)
    got = spmm_fused(pg, x.cuda(), bias=bias.cuda(), relu=True)
    assert np.array_equal(_bits(got), want.view(np.int32))
    # input gate on the transposed (backward) aggregation + column sums of the gated input
    gate = torch.randn(b, pg.num_nodes, d, generator=gen).relu()  # an upstream ReLU output: ~half exact zeros
    scale = 1.0 / 0.9
    xg = torch.where(gate > 0, x * np.float32(scale), torch.zeros(()))
    want = c_oracle.spmm(c.t_rowptr, c.t_col, c.t_val, xg.numpy())
    got, colsum = spmm_fused(pg, x.cuda(), transpose=True, gate=gate.cuda(), gate_scale=scale, want_colsum=True)
    assert np.array_equal(_bits(got), want.view(np.int32))
    truth = xg.double().sum(dim=(0, 1))
    assert (colsum.cpu().double() - truth).abs().max() <= 1e-5 * xg.double().abs().sum(dim=(0, 1)).max()
    # run-to-run deterministic
    got2, colsum2 = spmm_fused(pg, x.cuda(), transpose=True, gate=gate.cuda(), gate_scale=scale, want_colsum=True)
    assert torch.equal(got, got2) and torch.equal(colsum, colsum2)


@pytest.mark.parametrize("net,b,d", [("LTA", 5, 64), ("LT", 3, 64), ("LTA", 9, 128), ("LTA", 2, 32)])
def test_spmm_fused_live_mask_bit_exact(graph_golden, net, b, d):
    """The 1-bit gate (live_out / live_in) is the same gate as the float activations it was derived from."""
    from leak_det_gnn_b200.ops import new_live_mask, spmm_fused, unpack_live_mask
    pg = _graph(graph_golden, net)
    n = pg.num_nodes
    gen = torch.Generator().manual_seed(5 + b)
    x = torch.randn(b, n, d, generator=gen).cuda()
    bias = torch.randn(d, generator=gen).cuda()
    live = new_live_mask(b, n, d, x.device)
    live.fill_(-1)  # every word must be overwritten
    y = spmm_fused(pg, x, bias=bias, relu=True, drop_p=0.25, drop_seed=99, live_out=live)
    assert torch.equal(y, spmm_fused(pg, x, bias=bias, relu=True, drop_p=0.25, drop_seed=99))
    assert torch.equal(unpack_live_mask(live), y > 0)
    g = torch.randn(b, n, d, generator=gen).cuda()
    want, want_cs = spmm_fused(pg, g, transpose=True, gate=y, gate_scale=4.0 / 3.0, want_colsum=True)
    got, got_cs = spmm_fused(pg, g, transpose=True, live_in=live, gate_scale=4.0 / 3.0, want_colsum=True)
    assert torch.equal(got, want) and torch.equal(got_cs, want_cs)
    with pytest.raises(Exception):
        spmm_fused(pg, g, transpose=True, gate=y, live_in=live)


@pytest.mark.parametrize("n,b,d", [(12000, 3, 128), (9001, 2, 64), (20000, 1, 32)])
def test_gather_path_live_mask_bit_exact(n, b, d):
    """Graphs too large for shared memory (BASELINE configs[4] shape, scaled down): the L2-gather kernel writes and reads
    the same 1-bit gates -- identical to gating with the float activations, column sums included."""
    from leak_det_gnn_b200 import ops
    from leak_det_gnn_b200.ops import PipeGraph, new_live_mask, spmm_fused, unpack_live_mask
    rng = np.random.default_rng(n)
    par = np.arange(1, n) - 1 - rng.integers(0, np.minimum(np.arange(1, n), 64))
    ei = torch.from_numpy(np.stack([np.concatenate([np.arange(1, n), par]), np.concatenate([par, np.arange(1, n)])]))
    pg = PipeGraph(ei, n)
    assert not ops._staged_ok(pg, d)
    gen = torch.Generator().manual_seed(n + b)
    x = torch.randn(b, n, d, generator=gen).cuda()
    bias = torch.randn(d, generator=gen).cuda()
    live = new_live_mask(b, n, d, x.device)
    live.fill_(-1)
    y = spmm_fused(pg, x, bias=bias, relu=True, drop_p=0.25, drop_seed=99, live_out=live)
    assert torch.equal(y, spmm_fused(pg, x, bias=bias, relu=True, drop_p=0.25, drop_seed=99))
    assert torch.equal(unpack_live_mask(live), y > 0)
    g = torch.randn(b, n, d, generator=gen).cuda()
    want, want_cs = spmm_fused(pg, g, transpose=True, gate=y, gate_scale=4.0 / 3.0, want_colsum=True)
    got, got_cs = spmm_fused(pg, g, transpose=True, live_in=live, gate_scale=4.0 / 3.0, want_colsum=True)
    assert torch.equal(got, want) and torch.equal(got_cs, want_cs)


def test_spmm_fused_dropout_statistics(graph_golden):
    from leak_det_gnn_b200.ops import spmm_fused
    pg = _graph(graph_golden, "LTA")
    x = torch.randn(64, 661, 64, device="cuda").abs() + 0.1   # strictly positive -> relu never zeroes
    base = spmm_fused(pg, x, relu=True)
    assert (base > 0).all()
    p = 0.1
    a = spmm_fused(pg, x, relu=True, drop_p=p, drop_seed=1234)
    b = spmm_fused(pg, x, relu=True, drop_p=p, drop_seed=1234)
    c2 = spmm_fused(pg, x, relu=True, drop_p=p, drop_seed=1235)
    assert torch.equal(a, b) and not torch.equal(a, c2)          # keyed, reproducible
    dropped = (a == 0)
    frac = dropped.float().mean().item()
    assert abs(frac - p) < 2e-3, frac                            # 2.7M samples: sigma ~ 1.8e-4
    kept = ~dropped
    p_eff = round(p * 65536) / 65536                               # 16 random bits per element
    assert torch.equal(a[kept], (base * np.float32(1.0 / np.float32(1.0 - np.float32(p_eff))))[kept])  # inverted dropout scaling, exact
    # no structure across features / nodes / windows
    for dim in (0, 1, 2):
        other = tuple(i for i in range(3) if i != dim)
        m = dropped.float().mean(dim=other)
        assert (m - p).abs().max() < 0.02
    both = ((a == 0) & (c2 == 0)).float().mean().item()
    assert abs(both - p * p) < 2e-3                              # independent streams
