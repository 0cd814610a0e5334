"""Normal-state predictor, API- and checkpoint-compatible with reference models/predictor.py.

The predictor feeds the hot path: detector training runs it frozen on ``l_det`` overlapping windows of every
segment to form the residual sequence (models/utils.py:169-216, train_detector.py:302-307; SURVEY.md section 8f
rank 1).  It contains no graph layers.  ``state_dict`` keys equal the reference's
(``input_proj.{weight,bias}``, ``tcn.{i}.conv{1,2}.conv.{weight,bias}``, ``tcn.{i}.norm{1,2}.{weight,bias}``,
``head.{weight,bias}``), so ``predictor_best.ckpt`` files load unchanged.

``forward`` is the reference's dense evaluation (every time step of every layer); ``forward_last`` evaluates only the
dependency cone of the last time step -- the only one the model reads (predictor.py:80): 99 instead of 288
convolution evaluations per window -- and is what the residual builder (models/utils.py) calls.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class CausalConv1d(nn.Module):
    """Left-padded dilated convolution: output t sees inputs t, t - d, t - 2 d (predictor.py:17-28)."""

    def __init__(self, in_ch: int, out_ch: int, kernel_size: int, dilation: int = 1) -> None:
        super().__init__()
        self.pad = (kernel_size - 1) * dilation
        self.conv = nn.Conv1d(in_ch, out_ch, kernel_size, dilation=dilation, padding=self.pad)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = self.conv(x)
        return y[..., : y.shape[-1] - self.pad] if self.pad > 0 else y


class TCNBlock(nn.Module):
    """x + drop(relu(norm2(conv2(drop(relu(norm1(conv1(x)))))))), LayerNorm over channels (predictor.py:31-52)."""

    def __init__(self, channels: int, kernel_size: int, dilation: int, dropout: float) -> None:
        super().__init__()
        self.conv1 = CausalConv1d(channels, channels, kernel_size, dilation=dilation)
        self.conv2 = CausalConv1d(channels, channels, kernel_size, dilation=dilation)
        self.dropout = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(channels)
        self.norm2 = nn.LayerNorm(channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = x
        for conv, norm in ((self.conv1, self.norm1), (self.conv2, self.norm2)):
            y = self.dropout(F.relu(norm(conv(y).transpose(1, 2)))).transpose(1, 2)
        return x + y


class NormalPredictorTCN(nn.Module):
    """(B, L, S) noisy pressures + (B, L, 9) time features -> (B, S) next-step prediction (predictor.py:55-81)."""

    def __init__(self, num_sensors: int, time_dim: int = 9, hidden_channels: int = 128, kernel_size: int = 3,
                 num_blocks: int = 4, dropout: float = 0.1) -> None:
        super().__init__()
        self.num_sensors = int(num_sensors)
        self.time_dim = int(time_dim)
        self.input_proj = nn.Conv1d(self.num_sensors + self.time_dim, hidden_channels, kernel_size=1)
        self.tcn = nn.Sequential(*[TCNBlock(hidden_channels, kernel_size, 2 ** i, dropout) for i in range(num_blocks)])
        self.head = nn.Linear(hidden_channels, self.num_sensors)

    def forward(self, x: torch.Tensor, x_time: torch.Tensor) -> torch.Tensor:
        h = self.input_proj(torch.cat([x, x_time], dim=-1).transpose(1, 2))
        return self.head(self.tcn(h)[:, :, -1])

    @torch.no_grad()
    def forward_last(self, x: torch.Tensor, x_time: torch.Tensor) -> torch.Tensor:
        """Same result as ``forward`` in eval mode, computing only what the last time step depends on.
        Every convolution becomes one (rows x 3C) @ (3C x C) product over the needed (window, position) rows, with
        zero rows where a tap reaches before the window start (the left padding of CausalConv1d)."""
        if self.training:
            raise RuntimeError("forward_last is the inference path of the frozen predictor (call .eval() first)")
        b, length, _ = x.shape
        gather = self._cone_plan(length, x.device)
        h = F.linear(torch.cat([x, x_time], dim=-1), self.input_proj.weight[:, :, 0], self.input_proj.bias)  # (B, L, C)
        zero = h.new_zeros(b, 1, h.shape[-1])
        for blk, (taps1, taps2, keep) in zip(self.tcn, gather):
            y = h
            for conv, norm, taps in ((blk.conv1, blk.norm1, taps1), (blk.conv2, blk.norm2, taps2)):
                w = conv.conv.weight.permute(0, 2, 1).reshape(conv.conv.weight.shape[0], -1)   # (C, ks * C), tap-major
                padded = torch.cat([y, zero], dim=1)            # last index = the zero row (left padding)
                rows = padded.index_select(1, taps).view(b, -1, w.shape[1])    # (B, n_out, ks * C)
                y = F.relu(norm(F.linear(rows, w, conv.conv.bias)))
            h = h.index_select(1, keep) + y
        return self.head(h[:, 0, :])

    def _cone_plan(self, length: int, device):
        """Per block: flat tap indices of conv1 / conv2 (oldest tap first, `n_in` = the zero row) and the positions of
        the block input that survive as residual inputs; built once per (length, device)."""
        cache = self.__dict__.setdefault("_cone_cache", {})
        key = (length, str(device))
        if key in cache:
            return cache[key]
        dil = [blk.conv1.conv.dilation[0] for blk in self.tcn]
        ks = self.tcn[0].conv1.conv.kernel_size[0]
        plan = cone_positions(length, ks, dil)
        pos = list(range(length))
        out = []
        for d, (p1, p2) in zip(dil, plan):
            blocks = []
            src = pos
            for out_pos in (p1, p2):
                idx = {t: i for i, t in enumerate(src)}
                sel = [idx.get(t - j * d, len(src)) if t - j * d >= 0 else len(src) for t in out_pos
                       for j in range(ks - 1, -1, -1)]          # weight[:, :, 0] multiplies the OLDEST tap
                blocks.append(torch.tensor(sel, dtype=torch.long, device=device))
                src = out_pos
            idx = {t: i for i, t in enumerate(pos)}
            keep = torch.tensor([idx[t] for t in p2], dtype=torch.long, device=device)
            out.append((blocks[0], blocks[1], keep))
            pos = p2
        assert pos == [length - 1]
        cache[key] = out
        return out


class NormalPredictorGRU(nn.Module):
    """Two-layer GRU baseline (predictor.py:84-111); cuDNN, not a kernel target."""

    def __init__(self, num_sensors: int, time_dim: int = 9, hidden_size: int = 128, num_layers: int = 2,
                 dropout: float = 0.1) -> None:
        super().__init__()
        self.num_sensors = int(num_sensors)
        self.time_dim = int(time_dim)
        self.gru = nn.GRU(input_size=self.num_sensors + self.time_dim, hidden_size=hidden_size, num_layers=num_layers,
                          batch_first=True, dropout=dropout if num_layers > 1 else 0.0)
        self.head = nn.Linear(hidden_size, self.num_sensors)

    def forward(self, x: torch.Tensor, x_time: torch.Tensor) -> torch.Tensor:
        out, _ = self.gru(torch.cat([x, x_time], dim=-1))
        return self.head(out[:, -1, :])


def cone_positions(length: int, kernel_size: int, dilations: List[int]) -> List[Tuple[List[int], List[int]]]:
    """Time steps whose values the LAST step depends on, per convolution (two per block), walking backwards:
    element i = (positions needed of conv1's output, positions needed of conv2's output) of block i.
    For L = 36, k = 3, d = 1, 2, 4, 8 that is 36, 18, 18, 9, 9, 5, 3, 1 positions (99) instead of 8 x 36 (288)."""
    need = {length - 1}
    out: List[Tuple[List[int], List[int]]] = []
    for d in reversed(dilations):
        conv2 = sorted(need)                                   # block output positions = conv2 output positions
        conv1 = sorted({t - j * d for t in need for j in range(kernel_size) if t - j * d >= 0})
        need = set(conv1) | need                               # the residual x + y needs x at the block's outputs too
        need = {t - j * d for t in conv1 for j in range(kernel_size) if t - j * d >= 0} | need
        out.append((conv1, conv2))
    return list(reversed(out))
