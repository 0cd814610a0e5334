"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement of the reference leak detector's forward pass
(reference models/detector.py:28-73 sensor GRU encoder, :76-102 heads, :105-114
``_batchify_edge_index``, :123-218 ``LeakDetector``), built on the PyG restatement in
``oracle/pyg_restatement.py``.  Works in fp32 (the parity yardstick) and fp64 (the error
yardstick).  Parameter names / shapes equal the reference ``state_dict`` (18 tensors,
SURVEY.md section 8b) so one ``state_dict`` drives the oracle, the reference module
(when it can be imported) and the CUDA drop-in.

PARITY UNPINNED for the two PyG operators (see pyg_restatement.py header); the rest of
this file is cross-checked against the reference's own ``LeakDetector`` code in
tests/golden/make_goldens.py (bit-identical logits on CPU, recorded in the golden file).

The graph is passed in as plain tensors so nothing here touches the product package.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .pyg_restatement import GCNConv, global_mean_pool


class _SensorGRU(nn.Module):
    """detector.py:28-73 -- one shared single-layer GRU over every sensor's series."""

    def __init__(self, hidden: int, use_time: bool, time_dim: int = 9) -> None:
        super().__init__()
        self.use_time = use_time
        self.gru = nn.GRU(input_size=1 + (time_dim if use_time else 0), hidden_size=hidden,
                          num_layers=1, batch_first=True, dropout=0.0)

    def forward(self, r: torch.Tensor, tfeat: Optional[torch.Tensor]) -> torch.Tensor:
        b, l, s = r.shape
        seq = r.transpose(1, 2).contiguous().view(b * s, l, 1)
        if self.use_time:
            if tfeat is None:
                raise ValueError("tfeat required when use_time=True")
            t = tfeat.unsqueeze(1).repeat(1, s, 1, 1).contiguous().view(b * s, l, -1)
            seq = torch.cat([seq, t], dim=-1)
        out, _ = self.gru(seq)
        return out[:, -1, :].view(b, s, -1)


class _Head(nn.Module):
    """detector.py:76-102 -- Linear -> ReLU -> Dropout -> Linear(.,1); ``mlp.0``/``mlp.3``."""

    def __init__(self, in_dim: int, hidden: int, dropout: float) -> None:
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Dropout(dropout), nn.Linear(hidden, 1))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.mlp(x).squeeze(-1)


class OracleLeakDetector(nn.Module):
    def __init__(self, num_nodes: int, edge_index: torch.Tensor, pipe_ends: torch.Tensor,
                 sensor_node_idx: Sequence[int] | torch.Tensor, sensor_hidden: int = 64,
                 node_hidden: int = 64, gnn_layers: int = 2, dropout: float = 0.1,
                 use_time: bool = True) -> None:
        super().__init__()
        self.num_nodes = int(num_nodes)
        self.edge_index_single = edge_index.to(torch.long)
        self.pipe_ends = torch.as_tensor(pipe_ends, dtype=torch.long)
        self.sensor_node_idx = torch.as_tensor(sensor_node_idx, dtype=torch.long)
        # construction order == reference (detector.py:158-168) so a seeded init draws
        # the same random stream
        self.sensor_encoder = _SensorGRU(sensor_hidden, use_time)
        self.sensor_to_node = nn.Linear(sensor_hidden + 1, node_hidden)
        self.convs = nn.ModuleList([GCNConv(node_hidden, node_hidden) for _ in range(gnn_layers)])
        self.dropout = nn.Dropout(dropout)
        self.edge_head = _Head(node_hidden * 3, 128, dropout)
        self.noleak_head = _Head(node_hidden, 128, dropout)

    def forward(self, residual: torch.Tensor, tfeat: Optional[torch.Tensor] = None,
                return_intermediates: bool = False):
        return self.gnn_stack(self.sensor_encoder(residual, tfeat), return_intermediates)

    def gnn_stack(self, h_s: torch.Tensor, return_intermediates: bool = False):
        """detector.py:178-218 from the sensor embeddings on (the message-passing hot path)."""
        b = h_s.shape[0]
        n = self.num_nodes

        # detector.py:179-190 : zero-padded node tensor + sensor mask -> Linear -> ReLU
        h0 = torch.zeros(b, n, h_s.shape[-1], dtype=h_s.dtype)
        h0[:, self.sensor_node_idx, :] = h_s
        mask = torch.zeros(n, 1, dtype=h_s.dtype)
        mask[self.sensor_node_idx, 0] = 1.0
        h = torch.cat([h0, mask.unsqueeze(0).expand(b, -1, -1)], dim=-1)
        h = self.dropout(F.relu(self.sensor_to_node(h)))
        x = h.reshape(b * n, -1)
        inter = {"h_s": h_s, "x0": x.view(b, n, -1)}

        # detector.py:105-114,195-201 : B-times replicated graph, conv -> relu -> dropout
        e = self.edge_index_single.size(1)
        off = torch.arange(b).repeat_interleave(e) * n
        edge_index = self.edge_index_single.repeat(1, b) + off.unsqueeze(0)
        for li, conv in enumerate(self.convs):
            x = self.dropout(F.relu(conv(x, edge_index)))
            inter[f"x{li + 1}"] = x.view(b, n, -1)

        # detector.py:204-218 : per-pipe head on [h_u, h_v, |h_u-h_v|], pooled no-leak head
        h_nodes = x.view(b, n, -1)
        h_u = h_nodes[:, self.pipe_ends[:, 0], :]
        h_v = h_nodes[:, self.pipe_ends[:, 1], :]
        pipe_logits = self.edge_head(torch.cat([h_u, h_v, (h_u - h_v).abs()], dim=-1))
        batch = torch.arange(b).repeat_interleave(n)
        pooled = global_mean_pool(x, batch)
        noleak = self.noleak_head(pooled).unsqueeze(-1)
        logits = torch.cat([pipe_logits, noleak], dim=-1)
        if return_intermediates:
            inter["pooled"] = pooled
            return logits, inter
        return logits
