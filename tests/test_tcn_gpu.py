"""SURVEY 8f rank 1, native arm: the gathered-row convolution kernel of the frozen TCN predictor (csrc/tcn.cu,
ltgnn_tcn_conv) against an fp64 restatement of reference models/predictor.py:17-52 on the same gathered rows, and the
whole dependency cone (NormalPredictorTCN.forward_last on CUDA) against the golden minted by the reference's own code.
Tolerance: rel <= 1e-5 (fp32 path, 3xTF32 products)."""
import pytest
import torch

from conftest import GOLDEN, parity_log, rel_err
from leak_det_gnn_b200 import ops
from leak_det_gnn_b200.models import NormalPredictorTCN, build_residual_sequence_from_segment

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _conv_ref(x, src, w, bias, gamma, beta, eps, relu, res, res_row):
    """y[m] = [res[res_row[m]] +] relu(LN(bias + sum_t x[src[t, m]] @ w[t].T)) in fp64; src == -1 is the zero row."""
    xd = torch.cat([x.double(), x.new_zeros(1, x.shape[1]).double()])
    idx = torch.where(src < 0, torch.full_like(src, x.shape[0]), src).long()
    y = bias.double().unsqueeze(0).expand(src.shape[1], -1).clone()
    for t in range(src.shape[0]):
        y = y + xd[idx[t]] @ w[t].double().t()
    if gamma is not None:
        y = torch.nn.functional.layer_norm(y, (y.shape[1],), gamma.double(), beta.double(), eps)
    if relu:
        y = torch.relu(y)
    if res is not None:
        y = y + res.double()[res_row.long()]
    return y


@pytest.mark.parametrize("m,rows,taps,norm,relu,with_res", [
    (1, 7, 3, True, True, True),
    (127, 300, 3, True, True, False),
    (128, 128, 1, False, False, False),
    (129, 50, 2, True, False, True),
    (385, 1000, 3, True, True, True),
    (148 * 384 + 77, 40000, 3, True, True, True),
    (5000, 9000, 8, True, True, False),
])
def test_tcn_conv_matches_fp64(m, rows, taps, norm, relu, with_res):
    g = torch.Generator(device="cpu").manual_seed(m * 31 + taps)
    x = torch.randn(rows, 128, generator=g).cuda()
    src = torch.randint(-1, rows, (taps, m), generator=g, dtype=torch.int64).to(torch.int32).cuda()
    if m > 3:
        src[:, 2] = -1                                   # a row with only zero padding: y = LN(bias)
    w = (torch.randn(taps, 128, 128, generator=g) / (128 * taps) ** 0.5).cuda()
    bias = torch.randn(128, generator=g).cuda()
    gamma = (1 + 0.1 * torch.randn(128, generator=g)).cuda() if norm else None
    beta = (0.1 * torch.randn(128, generator=g)).cuda() if norm else None
    res = torch.randn(rows, 128, generator=g).cuda() if with_res else None
    res_row = torch.randint(0, rows, (m,), generator=g, dtype=torch.int64).to(torch.int32).cuda() if with_res else None
    y = ops.tcn_conv(x, src, w, bias, gamma, beta, 1e-5, relu, res, res_row)
    ref = _conv_ref(x, src, w, bias, gamma, beta, 1e-5, relu, res, res_row)
    err = rel_err(y, ref)
    parity_log(f"tcn_conv M={m} taps={taps} norm={norm} res={with_res}", {"y": err, "tol": TOL})
    assert y.shape == (m, 128) and err <= TOL
    y2 = ops.tcn_conv(x, src, w, bias, gamma, beta, 1e-5, relu, res, res_row)
    assert torch.equal(y, y2)                            # no atomics, fixed order


def test_tcn_conv_rejects_bad_arguments():
    x = torch.randn(8, 128, device="cuda")
    src = torch.zeros(3, 4, dtype=torch.int32, device="cuda")
    w = torch.randn(3, 128, 128, device="cuda")
    b = torch.zeros(128, device="cuda")
    with pytest.raises(ValueError):
        ops.tcn_conv(x, src.long(), w, b)
    with pytest.raises(ValueError):
        ops.tcn_conv(x, src, w[:2], b)
    with pytest.raises(ValueError):
        ops.tcn_conv(x, src, w, b, res=x)
    with pytest.raises(RuntimeError):
        ops.tcn_conv(x[:, :64].contiguous(), src, w[:, :64, :64].contiguous(), b[:64])     # C != 128
    assert ops.tcn_conv(x, src[:, :0].contiguous(), w, b).shape == (0, 128)


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN / "residual_TCN.pt", map_location="cpu")


def _model(g):
    m = NormalPredictorTCN(g["noisy_seg"].shape[-1], 9)
    m.load_state_dict(g["state_dict"], strict=True)
    return m.eval().cuda()


def test_forward_last_native_matches_reference_golden(gold):
    m = _model(gold)
    x, t = gold["noisy_seg"][:, :36].cuda(), gold["time_seg"][:, :36].cuda()
    assert m._native_ok(x)
    y = m.forward_last(x, t)
    err = rel_err(y, gold["y_hat_first_window"])
    parity_log("TCN cone (native) vs reference forward", {"y": err, "tol": TOL})
    assert err <= TOL
    # the torch-op cone is the cross-check of the same plan
    m.use_native = False
    try:
        y_t = m.forward_last(x, t)
    finally:
        m.use_native = True
    assert rel_err(y, y_t.double()) <= TOL


def test_residual_builder_native(gold):
    m = _model(gold)
    with torch.no_grad():
        r = build_residual_sequence_from_segment(m, gold["noisy_seg"], gold["time_seg"], gold["l_pred"], gold["l_det"],
                                                 device="cuda")
        r2 = build_residual_sequence_from_segment(m, gold["noisy_seg"], gold["time_seg"], gold["l_pred"], gold["l_det"],
                                                  device="cuda")
    err = rel_err(r, gold["residual64"])
    parity_log("residual builder (native TCN) vs reference fp64", {"y": err, "tol": TOL})
    assert err <= TOL and torch.equal(r, r2)


def test_forward_last_native_large_batch():
    """ltown_dp256's shape: 256 segments x 36 windows of 36 steps, 29 sensors -- against the dense fp64 module."""
    torch.manual_seed(5)
    m = NormalPredictorTCN(29, 9).eval().cuda()
    x = torch.randn(2048, 36, 29, device="cuda")
    t = torch.randn(2048, 36, 9, device="cuda")
    y = m.forward_last(x, t)
    m64 = NormalPredictorTCN(29, 9).double().eval().cuda()
    m64.load_state_dict({k: v.double() for k, v in m.state_dict().items()})
    with torch.no_grad():
        ref = m64(x.double(), t.double())
    err = rel_err(y, ref)
    parity_log("TCN cone (native) B=2048 vs dense fp64", {"y": err, "tol": TOL})
    assert err <= TOL
