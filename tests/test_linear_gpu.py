"""tcgen05 3xTF32 dense layer vs an fp64 matmul.  Tolerance: the fp32 bar of the north star,
max|a-b| <= 1e-5 * max|b|; in practice the error is at fp32-sgemm level (~1e-6 or better)."""
import pytest
import torch

from conftest import rel_err
from leak_det_gnn_b200 import lib as L
from leak_det_gnn_b200.ops import linear_tc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,k,n", [(128, 32, 16), (128, 64, 64), (1, 64, 64), (661, 64, 64), (1000, 128, 64),
                                   (84608, 64, 64), (300, 96, 32), (257, 64, 128), (4096, 160, 16)])
@pytest.mark.parametrize("relu,use_bias", [(False, False), (True, True)])
def test_linear_matches_fp64(m, k, n, relu, use_bias):
    gen = torch.Generator().manual_seed(m + k + n)
    x = torch.randn(m, k, generator=gen)
    w = torch.randn(n, k, generator=gen) * 0.2
    b = torch.randn(n, generator=gen) if use_bias else None
    want = x.double() @ w.double().t()
    if b is not None:
        want = want + b.double()
    if relu:
        want = want.relu()
    got = linear_tc(x.cuda(), w.cuda(), None if b is None else b.cuda(), relu)
    torch.cuda.synchronize()
    err = rel_err(got, want)
    fp32 = rel_err((x @ w.t() + (0 if b is None else b)).relu() if relu else x @ w.t() + (0 if b is None else b), want)
    assert err <= 1e-5, (err, fp32)
    assert err <= 20 * max(fp32, 1e-7), (err, fp32)  # same league as a CPU sgemm


def test_linear_structured_inputs():
    """Identity / one-hot operands expose any swizzle, descriptor or lane-mapping mistake exactly."""
    for k, n in ((64, 64), (128, 64), (32, 128)):
        x = torch.arange(300 * k, dtype=torch.float32).reshape(300, k) / 1024.0
        w = torch.zeros(n, k)
        for j in range(n):
            w[j, (7 * j + 3) % k] = 1.0
        got = linear_tc(x.cuda(), w.cuda()).cpu()
        want = x[:, [(7 * j + 3) % k for j in range(n)]]
        assert torch.equal(got, want), (k, n, (got - want).abs().max())


def test_linear_shape_errors():
    x = torch.zeros(4, 48, device="cuda")
    with pytest.raises(L.LtgnnError, match="multiple of 32"):
        linear_tc(x, torch.zeros(16, 48, device="cuda"))
    with pytest.raises(L.LtgnnError, match="multiple of 16"):
        linear_tc(torch.zeros(4, 64, device="cuda"), torch.zeros(8, 64, device="cuda"))


def test_linear_transposed_and_gate():
    gen = torch.Generator().manual_seed(11)
    g = torch.randn(1500, 64, generator=gen)
    w = torch.randn(64, 96, generator=gen) * 0.3   # [K, N]: dX = G @ W
    gate = torch.randn(1500, 96, generator=gen)
    want = (g.double() @ w.double()) * (gate > 0).double() * 1.25
    got = linear_tc(g.cuda(), w.cuda(), transposed=True, gate=gate.cuda(), gate_scale=1.25)
    assert rel_err(got, want) <= 1e-5
    assert bool(((got.cpu() == 0) | (gate > 0)).all())  # gated-off entries are exact zeros
