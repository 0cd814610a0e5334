"""the frozen TCN predictor and the residual builder (reference models/predictor.py:17-81,
models/utils.py:169-216) against a golden minted by the reference's OWN code (tests/golden/make_goldens.py imports
both reference modules unmodified) -- this row's parity is pinned by reference code, not by a restatement."""
import pytest
import torch

from conftest import GOLDEN, rel_err
from leak_det_gnn_b200.models import NormalPredictorTCN, build_residual_sequence_from_segment
from leak_det_gnn_b200.models.predictor import cone_positions

TOL = 1e-5


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLDEN / "residual_TCN.pt", map_location="cpu")


def _model(g):
    m = NormalPredictorTCN(g["noisy_seg"].shape[-1], 9)
    assert list(m.state_dict().keys()) == list(g["state_dict"].keys())      # checkpoint layout == reference
    m.load_state_dict(g["state_dict"], strict=True)
    return m.eval()


def test_cone_positions():
    plan = cone_positions(36, 3, [1, 2, 4, 8])
    assert [(len(a), len(b)) for a, b in plan] == [(36, 18), (18, 9), (9, 5), (3, 1)]
    assert plan[-1][1] == [35] and plan[0][0] == list(range(36))
    assert cone_positions(5, 3, [1]) == [([2, 3, 4], [4])]


def test_dense_forward_equals_reference(gold):
    m = _model(gold)
    with torch.no_grad():
        y = m(gold["noisy_seg"][:, :36], gold["time_seg"][:, :36])
    assert rel_err(y, gold["y_hat_first_window"]) <= 1e-6                    # same torch ops (thread count may differ)


def test_residual_builder_cpu(gold):
    m = _model(gold)
    with torch.no_grad():
        r = build_residual_sequence_from_segment(m, gold["noisy_seg"], gold["time_seg"], gold["l_pred"], gold["l_det"])
        one = build_residual_sequence_from_segment(m, gold["noisy_seg"][2], gold["time_seg"][2], 36, 36)
    assert r.shape == gold["residual"].shape and one.shape == gold["residual"].shape[1:]
    assert rel_err(r, gold["residual64"]) <= TOL and rel_err(r, gold["residual"]) <= TOL
    assert rel_err(one, gold["residual64"][2]) <= TOL
    # with gradients enabled (or a module without forward_last) the builder takes the reference's dense route
    r2 = build_residual_sequence_from_segment(m, gold["noisy_seg"], gold["time_seg"], 36, 36)
    assert torch.equal(r2.detach(), gold["residual"])


@pytest.mark.gpu
def test_residual_builder_gpu(gold):
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        m = _model(gold).cuda()
        with torch.no_grad():
            r = build_residual_sequence_from_segment(m, gold["noisy_seg"], gold["time_seg"], 36, 36, device="cuda")
        assert r.is_cuda and rel_err(r, gold["residual64"]) <= TOL
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
