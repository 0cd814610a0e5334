"""Batched scenario windows and the CUDA-graph detector against the reference's loop shape: one B = 1 detector call per
window (eval/event_evaluator.py:478-492)."""
import numpy as np
import pytest
import torch

from conftest import TOPO, rel_err
from leak_det_gnn_b200.event_windows import GraphedDetector, scenario_window_logits
from leak_det_gnn_b200.models import LeakDetector, NormalPredictorTCN, build_residual_sequence_from_segment

pytestmark = pytest.mark.gpu


def _models(graph_golden, n_pipes=60):
    g = graph_golden("LTA")
    sensors, pipes = [str(s) for s in g["sensor_node_ids"]], [str(p) for p in g["pipe_ids"]][:n_pipes]
    torch.manual_seed(11)
    det = LeakDetector(TOPO["LTA"], sensors, pipes).cuda().eval()
    pred = NormalPredictorTCN(29, 9).cuda().eval()
    return pred, det


def test_batched_windows_equal_one_call_per_window(graph_golden):
    pred, det = _models(graph_golden)
    rng = np.random.default_rng(0)
    T, l_pred, l_det, stride = 200, 36, 36, 3
    pressure, tfeat = rng.standard_normal((T, 29)).astype(np.float32), rng.standard_normal((T, 9)).astype(np.float32)
    ends, logits = scenario_window_logits(pred, det, pressure, tfeat, l_pred, l_det, stride, "cuda", batch=16)
    want_ends, want = [], []
    with torch.no_grad():
        for t0 in range(l_pred, T - l_det + 1, stride):          # the reference's loop, one window at a time
            seg_p = torch.from_numpy(pressure[t0 - l_pred:t0 + l_det]).cuda()
            seg_f = torch.from_numpy(tfeat[t0 - l_pred:t0 + l_det]).cuda()
            res = build_residual_sequence_from_segment(pred, seg_p, seg_f, l_pred, l_det)
            want.append(det(res.unsqueeze(0), seg_f[l_pred:].unsqueeze(0))[0].cpu())
            want_ends.append(t0 + l_det)
    assert ends == want_ends and logits.shape == (len(want), 61)
    assert rel_err(logits, torch.stack(want)) <= 1e-5


@pytest.mark.parametrize("batch", [1, 8])
def test_graphed_detector_replays_the_forward(graph_golden, batch):
    _, det = _models(graph_golden)
    g = GraphedDetector(det, batch, 36)
    gen = torch.Generator().manual_seed(batch)
    for _ in range(3):
        r, t = torch.randn(batch, 36, 29, generator=gen).cuda(), torch.randn(batch, 36, 9, generator=gen).cuda()
        with torch.no_grad():
            want = det(r, t)
        assert torch.equal(g(r, t), want)
