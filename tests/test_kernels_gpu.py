"""Unit parity of the remaining hand-written kernels vs plain fp64 torch restatements of the reference ops."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from leak_det_gnn_b200 import ops

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.parametrize("m,do,di", [(1, 64, 64), (31, 64, 64), (33, 64, 64), (84608, 64, 64), (5000, 128, 128),
                                     (777, 64, 128), (100, 32, 64), (1000, 24, 40)])
def test_wgrad(m, do, di):
    gen = torch.Generator().manual_seed(m + do)
    g = torch.randn(m, do, generator=gen)
    x = torch.randn(m, di, generator=gen)
    want = g.double().t() @ x.double()
    got = ops.wgrad(g.cuda(), x.cuda())
    assert got.shape == (do, di)
    # a sum of m products: compare to fp64 relative to the size of the terms (sqrt(m) growth)
    assert (got.cpu().double() - want).abs().max() <= TOL * max(want.abs().max().item(), m ** 0.5)
    assert rel_err(got, want) <= 1e-4
    assert torch.equal(got, ops.wgrad(g.cuda(), x.cuda()))  # deterministic


def _node_init_ref(h_s, idx, n, w, b):
    """reference detector.py:178-189 in fp64 (dropout off)."""
    bsz, s, ds = h_s.shape
    h0 = torch.zeros(bsz, n, ds, dtype=torch.float64)
    h0[:, idx, :] = h_s
    mask = torch.zeros(n, 1, dtype=torch.float64)
    mask[idx, 0] = 1.0
    h = torch.cat([h0, mask.unsqueeze(0).expand(bsz, -1, -1)], dim=-1)
    return torch.relu(torch.nn.functional.linear(h, w, b))


@pytest.mark.parametrize("bsz,n,s,ds,d", [(1, 661, 29, 64, 64), (37, 661, 29, 64, 64), (3, 785, 29, 32, 128),
                                          (700, 50, 7, 16, 32), (2, 5000, 900, 64, 128)])  # last: sensors too many to stage
def test_node_init_fwd_bwd(bsz, n, s, ds, d):
    gen = torch.Generator().manual_seed(bsz + n)
    idx = torch.randperm(n, generator=gen)[:s]
    slot = torch.full((n,), -1, dtype=torch.int32)
    slot[idx] = torch.arange(s, dtype=torch.int32)
    h_s = torch.randn(bsz, s, ds, generator=gen)
    w = torch.randn(d, ds + 1, generator=gen) * 0.2
    b = torch.randn(d, generator=gen) * 0.5
    x64 = _node_init_ref(h_s.double(), idx, n, w.double(), b.double())
    dx = torch.randn(bsz, n, d, generator=gen)

    x0 = ops.node_init_fwd(h_s.cuda(), slot.cuda(), n, w.cuda(), b.cuda())
    assert rel_err(x0, x64) <= TOL
    # fp64 truth of the backward, through autograd, with the ReLU mask the fp32 forward actually produced
    # (a pre-activation within fp32 rounding of zero may land on the other side in fp64; the backward is
    # defined by the forward that ran)
    h64 = h_s.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    b64 = b.double().requires_grad_(True)
    h0 = torch.zeros(bsz, n, ds, dtype=torch.float64)
    h0 = h0.index_copy(1, idx, h64)
    mask = torch.zeros(n, 1, dtype=torch.float64)
    mask[idx, 0] = 1.0
    pre = torch.nn.functional.linear(torch.cat([h0, mask.unsqueeze(0).expand(bsz, -1, -1)], dim=-1), w64, b64)
    (pre * (x0.cpu() > 0).double()).backward(dx.double())
    # non-sensor rows are exactly relu(bias)
    non = torch.ones(n, dtype=torch.bool)
    non[idx] = False
    assert torch.equal(x0[:, non.cuda(), :].cpu(), torch.relu(b).expand(bsz, int(non.sum()), d))
    if d not in (64, 128) or ds % 32:
        # widths the backward kernels are not built for: a loud error, not a library fallback
        with pytest.raises(ValueError, match="not built"):
            ops.node_init_bwd(h_s.cuda(), slot.cuda(), w.cuda(), dx.cuda(), x0, 1.0)
        return
    dhs, dw, db = ops.node_init_bwd(h_s.cuda(), slot.cuda(), w.cuda(), dx.cuda(), x0, 1.0)
    assert rel_err(dhs, h64.grad) <= TOL
    assert rel_err(dw, w64.grad) <= 5 * TOL
    assert rel_err(db, b64.grad) <= 5 * TOL
    a = ops.node_init_bwd(h_s.cuda(), slot.cuda(), w.cuda(), dx.cuda(), x0, 1.0)
    assert all(torch.equal(u, v) for u, v in zip(a, (dhs, dw, db)))  # deterministic


def test_node_init_dropout():
    n, s, ds, d, bsz = 661, 29, 64, 64, 200
    idx = torch.arange(0, n, 23)[:s]
    slot = torch.full((n,), -1, dtype=torch.int32)
    slot[idx] = torch.arange(s, dtype=torch.int32)
    h_s = torch.randn(bsz, s, ds).cuda()
    w = (torch.randn(d, ds + 1) * 0.2).cuda()
    b = (torch.rand(d) + 0.1).cuda()       # positive bias -> every non-sensor entry is live before dropout
    base = ops.node_init_fwd(h_s, slot.cuda(), n, w, b)
    dr = ops.node_init_fwd(h_s, slot.cuda(), n, w, b, 0.1, 99)
    live = base > 0
    frac = ((dr == 0) & live).float().sum().item() / live.float().sum().item()
    assert abs(frac - 0.1) < 3e-3
    kept = dr != 0
    p_eff = np.float32(round(0.1 * 65536) / 65536)                 # 16 random bits per element
    assert torch.equal(dr[kept], (base * np.float32(1.0 / np.float32(1.0 - p_eff)))[kept])
    assert torch.equal(dr, ops.node_init_fwd(h_s, slot.cuda(), n, w, b, 0.1, 99))


def _head_ref(x, ends, w1, b1, w2, b2):
    h_u, h_v = x[:, ends[:, 0], :], x[:, ends[:, 1], :]
    feat = torch.cat([h_u, h_v, (h_u - h_v).abs()], dim=-1)
    hid = torch.relu(torch.nn.functional.linear(feat, w1, b1))
    return torch.nn.functional.linear(hid, w2, b2).squeeze(-1), x.mean(dim=1)


@pytest.mark.parametrize("bsz,n,p", [(3, 661, 764), (1, 661, 2), (5, 785, 905), (40, 100, 37)])
def test_heads_fwd_bwd(bsz, n, p):
    gen = torch.Generator().manual_seed(bsz * 7 + p)
    x = torch.randn(bsz, n, 64, generator=gen).relu()
    ends = torch.randint(0, n, (p, 2), generator=gen)
    ends[0, 1] = ends[0, 0]  # a degenerate pipe: |h_u - h_v| = 0 and sign(0) = 0
    w1 = torch.randn(128, 192, generator=gen) * 0.1
    b1 = torch.randn(128, generator=gen) * 0.1
    w2 = torch.randn(1, 128, generator=gen) * 0.2
    b2 = torch.randn(1, generator=gen)
    dlog = torch.randn(bsz, p, generator=gen)
    dpool = torch.randn(bsz, 64, generator=gen)

    t = [v.double().requires_grad_(True) for v in (x, w1, b1, w2, b2)]
    lr, pr = _head_ref(t[0], ends, *t[1:])
    ((lr * dlog.double()).sum() + (pr * dpool.double()).sum()).backward()

    o = [v.cuda().requires_grad_(True) for v in (x, w1, b1, w2)]
    b2c = b2.cuda().requires_grad_(True)
    part, pooled = ops.heads(o[0], ends.to(torch.int32).cuda(), o[1], o[2], o[3], 0.1, False)
    logits = part.sum(0) + b2c
    assert rel_err(logits, lr) <= TOL and rel_err(pooled, pr) <= TOL
    ((logits * dlog.cuda()).sum() + (pooled * dpool.cuda()).sum()).backward()
    assert rel_err(o[0].grad, t[0].grad) <= TOL
    assert rel_err(o[1].grad, t[1].grad) <= TOL
    assert rel_err(o[2].grad, t[2].grad) <= TOL
    assert rel_err(o[3].grad, t[3].grad) <= TOL
    assert rel_err(b2c.grad, t[4].grad) <= TOL


@pytest.mark.parametrize("bsz,n,p", [(3, 661, 764), (1, 50, 2), (40, 300, 600), (2, 20000, 1000)])
def test_heads_wide_fwd_bwd(bsz, n, p):
    """Node width 128 (BASELINE configs[4]; detector.py:76-88,204-216 at node_hidden = 128): the composed pipe head of
    csrc/heads_wide.cu + the three-tap gathered-row GEMM, forward and every gradient against fp64.  40 x 600 rows = 188
    row tiles through the GEMM; 20 000 nodes with 1 000 class pipes = mostly nodes without an incident pipe."""
    gen = torch.Generator().manual_seed(bsz * 7 + p)
    x = torch.randn(bsz, n, 128, generator=gen).relu()
    ends = torch.randint(0, n, (p, 2), generator=gen)
    ends[0, 1] = ends[0, 0]  # a degenerate pipe: |h_u - h_v| = 0 and sign(0) = 0
    w1 = torch.randn(128, 384, generator=gen) * 0.1
    b1 = torch.randn(128, generator=gen) * 0.1
    w2 = torch.randn(1, 128, generator=gen) * 0.2
    b2 = torch.randn(1, generator=gen)
    dlog = torch.randn(bsz, p, generator=gen)
    dpool = torch.randn(bsz, 128, generator=gen)

    o = [v.cuda().requires_grad_(True) for v in (x, w1, b1, w2)]
    b2c = b2.cuda().requires_grad_(True)
    ops.DEBUG_CAPTURE = cap = {}
    try:
        part, pooled = ops.heads(o[0], ends.to(torch.int32).cuda(), o[1], o[2], o[3], 0.1, False)
    finally:
        ops.DEBUG_CAPTURE = None
    logits = part.sum(0) + b2c
    ((logits * dlog.cuda()).sum() + (pooled * dpool.cuda()).sum()).backward()

    # fp64 truth with the ReLU decisions the kernel took (a pre-activation within fp32 rounding of zero may flip in fp64)
    live = cap["head_live"].cpu().double().view(bsz, p, 128)
    t = [v.double().requires_grad_(True) for v in (x, w1, b1, w2, b2)]
    h_u, h_v = t[0][:, ends[:, 0], :], t[0][:, ends[:, 1], :]
    feat = torch.cat([h_u, h_v, (h_u - h_v).abs()], dim=-1)
    hid = torch.nn.functional.linear(feat, t[1], t[2]) * live
    lr, pr = torch.nn.functional.linear(hid, t[3], t[4]).squeeze(-1), t[0].mean(dim=1)
    ((lr * dlog.double()).sum() + (pr * dpool.double()).sum()).backward()
    assert rel_err(logits, lr) <= TOL and rel_err(pooled, pr) <= TOL
    for got, want in zip(o + [b2c], t):
        assert rel_err(got.grad, want.grad) <= TOL


def test_heads_wide_dropout_and_determinism():
    gen = torch.Generator().manual_seed(6)
    x = (torch.randn(8, 300, 128, generator=gen).relu()).cuda().requires_grad_(True)
    ends = torch.randint(0, 300, (500, 2), generator=gen).to(torch.int32).cuda()
    w1 = (torch.randn(128, 384, generator=gen) * 0.05).cuda().requires_grad_(True)
    b1 = (torch.rand(128, generator=gen) + 5.0).cuda().requires_grad_(True)   # all hidden units live: dropout is the only zero source
    w2 = (torch.rand(1, 128, generator=gen) + 0.5).cuda().requires_grad_(True)

    def run(train):
        for t in (x, w1, b1, w2):
            t.grad = None
        torch.manual_seed(1)
        part, pooled = ops.heads(x, ends, w1, b1, w2, 0.1, train)
        (part.sum() + pooled.sum()).backward()
        return part.detach().clone(), [t.grad.clone() for t in (x, w1, b1, w2)]

    part, grads = run(True)
    part2, grads2 = run(True)
    base, _ = run(False)
    assert torch.equal(part, part2) and all(torch.equal(a, b) for a, b in zip(grads, grads2))   # bitwise run to run
    assert not torch.equal(part, base)
    assert abs((part.sum(0) / base.sum(0)).mean().item() - 1.0) < 5e-3                            # unbiased


def test_heads_dropout_train_mode():
    gen = torch.Generator().manual_seed(5)
    x = (torch.randn(16, 661, 64, generator=gen).relu()).cuda()
    ends = torch.randint(0, 661, (764, 2), generator=gen).to(torch.int32).cuda()
    w1 = (torch.randn(128, 192, generator=gen) * 0.1).cuda()
    b1 = (torch.rand(128, generator=gen) + 5.0).cuda()     # all hidden units live: dropout is the only zero source
    w2 = (torch.rand(1, 128, generator=gen) + 0.5).cuda().requires_grad_(True)
    torch.manual_seed(1)
    part, _ = ops.heads(x, ends, w1, b1, w2, 0.1, True)
    base, _ = ops.heads(x, ends, w1, b1, w2, 0.1, False)
    torch.manual_seed(1)
    part2, _ = ops.heads(x, ends, w1, b1, w2, 0.1, True)
    assert torch.equal(part, part2)                         # torch.manual_seed reproduces the masks
    assert not torch.equal(part, base)
    ratio = (part.sum(0) / base.sum(0)).mean().item()
    assert abs(ratio - 1.0) < 5e-3                          # inverted dropout is unbiased
    # backward uses the saved post-dropout activations: d/dw2 = sum dlogit * hidden
    part.sum().backward()
    assert w2.grad.shape == (1, 128) and torch.isfinite(w2.grad).all()


@pytest.mark.parametrize("bsz,n,s,ds,d", [(7, 661, 29, 64, 64), (3, 785, 33, 64, 128), (2, 40, 5, 32, 64)])
def test_node_init_live_mask(bsz, n, s, ds, d):
    """node_init's 1-bit record of x0 > 0 gates the backward exactly like x0 itself."""
    gen = torch.Generator().manual_seed(11)
    idx = torch.randperm(n, generator=gen)[:s]
    slot = torch.full((n,), -1, dtype=torch.int32)
    slot[idx] = torch.arange(s, dtype=torch.int32)
    slot = slot.cuda()
    h_s = torch.randn(bsz, s, ds, generator=gen).cuda()
    w = (torch.randn(d, ds + 1, generator=gen) * 0.2).cuda()
    b = (torch.randn(d, generator=gen) * 0.2).cuda()
    live = ops.new_live_mask(bsz, n, d, h_s.device)
    live.fill_(-1)
    x0 = ops.node_init_fwd(h_s, slot, n, w, b, 0.2, 321, live_out=live)
    assert torch.equal(x0, ops.node_init_fwd(h_s, slot, n, w, b, 0.2, 321))
    assert torch.equal(ops.unpack_live_mask(live), x0 > 0)
    dx0 = torch.randn(bsz, n, d, generator=gen).cuda()
    want = ops.node_init_bwd(h_s, slot, w, dx0, x0, 1.25)
    got = ops.node_init_bwd(h_s, slot, w, dx0, x0, 1.25, live=live)
    for a, c in zip(got, want):
        assert torch.equal(a, c)
