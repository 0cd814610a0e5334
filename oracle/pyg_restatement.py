"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (plain torch ops, any float dtype) of the two third-party operators the
reference's hot path delegates to:

    torch_geometric.nn.GCNConv(in, out, add_self_loops=True, normalize=True)
    torch_geometric.nn.global_mean_pool

Reference call sites: models/detector.py:23 (import), :162-164 (construction),
:199 (``x = conv(x, edge_index)``), :215 (``global_mean_pool(x, batch)``).

The package is a dependency that is NOT vendored in /root/reference and NOT pinned by
it (no requirements / lock file; README.md:1-2 only names Python 3.11), and it is not
installed in the build image.  The algorithm below is restated from PyG 2.x's published
source (torch_geometric/nn/conv/gcn_conv.py ``gcn_norm`` + ``GCNConv.forward``;
torch_geometric/utils/loop.py ``add_remaining_self_loops``;
torch_geometric/nn/pool/glob.py ``global_mean_pool``; torch_geometric/nn/dense/linear.py
``Linear`` + torch_geometric/nn/inits.py ``glorot``).

PARITY UNPINNED: the reference ships no tests, golden vectors or known-answer data for
this path, and PyG cannot be imported here to cross-check, so the
operator semantics rest on the published algorithm.  What IS pinned: everything around
the two operators runs as the reference's own unmodified code when the goldens are
minted (tests/golden/make_goldens.py imports /root/reference/models/detector.py with
these two names injected as ``torch_geometric.nn``).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn


def add_remaining_self_loops(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """loop.py: drop existing self loops, append (i,i) for every node after all edges."""
    keep = edge_index[0] != edge_index[1]
    loops = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    loops = loops.unsqueeze(0).repeat(2, 1)
    return torch.cat([edge_index[:, keep], loops], dim=1)


def gcn_norm(edge_index: torch.Tensor, num_nodes: int, dtype=torch.float32) -> Tuple[torch.Tensor, torch.Tensor]:
    """gcn_conv.py ``gcn_norm`` with edge_weight=None, improved=False, add_self_loops=True,
    flow='source_to_target'."""
    edge_index = add_remaining_self_loops(edge_index, num_nodes)
    w = torch.ones(edge_index.size(1), dtype=dtype, device=edge_index.device)
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(num_nodes, dtype=dtype, device=edge_index.device).scatter_add_(0, col, w)
    dis = deg.pow_(-0.5)
    dis.masked_fill_(dis == float("inf"), 0)
    return edge_index, dis[row] * w * dis[col]


def propagate_add(edge_index: torch.Tensor, norm: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """message = norm * x_j (j = edge source, row 0); aggr='add' at the target (row 1).

    ``index_add_`` on CPU walks the index sequentially, i.e. each target row receives its
    messages in edge-list order with the self loop last.
    """
    row, col = edge_index[0], edge_index[1]
    out = torch.zeros(x.size(0), x.size(1), dtype=x.dtype, device=x.device)
    return out.index_add_(0, col, norm.unsqueeze(-1) * x.index_select(0, row))


class _Lin(nn.Module):
    """PyG ``Linear(in, out, bias=False, weight_initializer='glorot')``: parameter name
    ``weight`` of shape (out, in); draws its glorot init inside ``__init__``."""

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        a = math.sqrt(6.0 / (self.weight.size(-2) + self.weight.size(-1)))
        self.weight.data.uniform_(-a, a)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return torch.nn.functional.linear(x, self.weight)


class GCNConv(nn.Module):
    """Restatement of ``GCNConv.forward`` for the configuration the reference uses.

    ``state_dict`` keys: ``bias``, ``lin.weight`` (reference checkpoint layout,
    SURVEY.md section 8b).  ``cached=False``: the normalisation is recomputed on every
    call exactly like the reference run does.
    """

    def __init__(self, in_channels: int, out_channels: int, add_self_loops: bool = True, normalize: bool = True) -> None:
        super().__init__()
        if not (add_self_loops and normalize):
            raise NotImplementedError("oracle covers the reference configuration only")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.lin = _Lin(in_channels, out_channels)  # draw #1 (Linear.__init__)
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()  # draw #2 (GCNConv.reset_parameters)

    def reset_parameters(self) -> None:
        self.lin.reset_parameters()
        nn.init.zeros_(self.bias)

    def forward(self, x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
        ei, norm = gcn_norm(edge_index, x.size(0), dtype=x.dtype)
        x = self.lin(x)
        out = propagate_add(ei, norm, x)
        return out + self.bias


def global_mean_pool(x: torch.Tensor, batch: Optional[torch.Tensor], size: Optional[int] = None) -> torch.Tensor:
    """glob.py: ``scatter(x, batch, dim=0, dim_size=size, reduce='mean')``."""
    if batch is None:
        return x.mean(dim=0, keepdim=True)
    if size is None:
        size = int(batch.max()) + 1
    out = torch.zeros(size, x.size(1), dtype=x.dtype, device=x.device).index_add_(0, batch, x)
    cnt = torch.zeros(size, dtype=x.dtype, device=x.device).index_add_(0, batch, torch.ones_like(batch, dtype=x.dtype))
    return out / cnt.clamp_(min=1).unsqueeze(-1)
