"""Pins the PyG restatement (oracle/pyg_restatement.py) against the REAL torch_geometric wherever it is importable.
The reference delegates its graph arithmetic to ``torch_geometric.nn.GCNConv`` / ``global_mean_pool``
(models/detector.py:23,162-164,199,215), un-vendored and un-pinned; the build image has no PyG, so these tests skip there
and the oracle's header says "parity unpinned".  On a machine with PyG they turn that into a pinned check and record the
version (conftest.parity_log)."""
import pytest
import torch

from conftest import parity_log, rel_err
from oracle import pyg_restatement as pyg

tg = pytest.importorskip("torch_geometric")
from torch_geometric.nn import GCNConv, global_mean_pool  # noqa: E402


def test_gcnconv_restatement_equals_pyg(graph_golden):
    g0 = graph_golden("LTA")
    ei = torch.from_numpy(g0["edge_index"])
    n, b = len(g0["node_names"]), 3
    torch.manual_seed(0)
    real = GCNConv(64, 64, add_self_loops=True, normalize=True)
    torch.manual_seed(0)
    ours = pyg.GCNConv(64, 64)
    same_init = torch.equal(real.lin.weight, ours.lin.weight)      # glorot drawn twice, bias zero
    ours.load_state_dict(real.state_dict())
    big = ei.repeat(1, b) + (torch.arange(b).repeat_interleave(ei.size(1)) * n).unsqueeze(0)
    x = torch.randn(b * n, 64)
    y_real, y_ours = real(x, big), ours(x, big)
    parity_log("pyg_pin GCNConv", {"torch_geometric": tg.__version__, "same_seeded_init": same_init,
                                   "forward_rel_err": rel_err(y_ours, y_real)})
    assert rel_err(y_ours, y_real) <= 1e-6
    assert list(real.state_dict().keys()) == list(ours.state_dict().keys()) == ["bias", "lin.weight"]


def test_global_mean_pool_restatement_equals_pyg():
    x = torch.randn(5 * 661, 64)
    batch = torch.arange(5).repeat_interleave(661)
    assert rel_err(pyg.global_mean_pool(x, batch), global_mean_pool(x, batch)) <= 1e-6
