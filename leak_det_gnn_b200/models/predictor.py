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

    # use the hand-written sm_100a kernels for ``forward_last`` on CUDA (csrc/tcn.cu + csrc/linear.cu) where the shapes
    # allow; False = the same cone through torch ops (cuBLAS), kept as the cross-check
    use_native = True

    def _native_ok(self, x: torch.Tensor) -> bool:
        blk = self.tcn[0]
        return (self.use_native and x.is_cuda and x.dtype == torch.float32 and self.input_proj.out_channels == 128
                and blk.conv1.conv.kernel_size[0] <= 8 and self.num_sensors + self.time_dim <= 64 and self.num_sensors <= 32)

    @torch.no_grad()
    def forward_last(self, x: torch.Tensor, x_time: torch.Tensor) -> torch.Tensor:
        """Same result as ``forward`` in eval mode, computing only what the last time step depends on.
        Every convolution becomes one (rows x 3C) @ (3C x C) product over the needed (window, position) rows, with
        zero rows where a tap reaches before the window start (the left padding of CausalConv1d)."""
        if self.training:
            raise RuntimeError("forward_last is the inference path of the frozen predictor (call .eval() first)")
        if self._native_ok(x):
            return self._forward_last_native(x, x_time)
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

    def _forward_last_native(self, x: torch.Tensor, x_time: torch.Tensor) -> torch.Tensor:
        """The cone on the tensor cores: input projection and output head through ``ops.linear_tc`` (zero-padded to the
        kernel's K / N granularity), every convolution + LayerNorm + ReLU (+ residual) through ``ops.tcn_conv``."""
        from .. import ops

        b, length, _ = x.shape
        dev = x.device
        prep = self._native_weights(dev)
        rows = self._native_rows(length, b, dev)
        c_in = self.num_sensors + self.time_dim
        inp = torch.zeros(b, length, 64, device=dev, dtype=torch.float32)
        inp[:, :, :self.num_sensors] = x
        inp[:, :, self.num_sensors:c_in] = x_time
        h = ops.linear_tc(inp.view(b * length, 64), prep["w_in"], prep["b_in"])                 # (B * L, 128)
        for blk, (w1, w2), (src1, src2, keep_rows) in zip(self.tcn, prep["convs"], rows):
            y = ops.tcn_conv(h, src1, w1, blk.conv1.conv.bias, blk.norm1.weight, blk.norm1.bias, blk.norm1.eps, True)
            h = ops.tcn_conv(y, src2, w2, blk.conv2.conv.bias, blk.norm2.weight, blk.norm2.bias, blk.norm2.eps, True,
                             res=h, res_row=keep_rows)
        return ops.linear_tc(h, prep["w_out"], prep["b_out"])[:, :self.num_sensors]

    def _native_weights(self, device):
        """Weights in the kernels' layouts, rebuilt when a parameter changes (the predictor is frozen in practice)."""
        cache = self.__dict__.setdefault("_native_w", {})
        ver = tuple(p._version for p in self.parameters()) + (str(device),)
        if cache.get("ver") == ver:
            return cache["prep"]
        c_in = self.num_sensors + self.time_dim
        w_in = torch.zeros(128, 64, device=device)
        w_in[:, :c_in] = self.input_proj.weight[:, :, 0]
        w_out = torch.zeros(32, 128, device=device)
        w_out[: self.num_sensors] = self.head.weight
        b_out = torch.zeros(32, device=device)
        b_out[: self.num_sensors] = self.head.bias
        convs = [(blk.conv1.conv.weight.permute(2, 0, 1).contiguous(), blk.conv2.conv.weight.permute(2, 0, 1).contiguous())
                 for blk in self.tcn]                       # (taps, C_out, C_in): tap 0 = the oldest input
        prep = {"w_in": w_in, "b_in": self.input_proj.bias.detach().contiguous(), "w_out": w_out, "b_out": b_out,
                "convs": convs}
        cache.update(ver=ver, prep=prep)
        return prep

    def _native_rows(self, length: int, batch: int, device):
        """Global row indices of the gathered convolutions for ``batch`` windows: per block (src1, src2, keep_rows) with
        src int32 (taps, batch * n_out), -1 = zero padding; activations are laid out (window, position, channel)."""
        cache = self.__dict__.setdefault("_native_r", {})
        key = (length, batch, str(device))
        if key in cache:
            return cache[key]
        ks = self.tcn[0].conv1.conv.kernel_size[0]
        win = torch.arange(batch, device=device, dtype=torch.int64).unsqueeze(1)
        out = []
        n_in = length
        for taps1, taps2, keep in self._cone_plan(length, device):
            rows = []
            n_src = n_in
            for taps in (taps1, taps2):
                rel = taps.view(-1, ks).t()                                   # (ks, n_out), value n_src = the zero row
                n_out = rel.shape[1]
                glob = torch.where(rel.unsqueeze(1) == n_src, torch.full((), -1, device=device, dtype=torch.int64),
                                   win.unsqueeze(0) * n_src + rel.unsqueeze(1))   # (ks, batch, n_out)
                rows.append(glob.reshape(ks, batch * n_out).to(torch.int32).contiguous())
                n_src = n_out
            keep_rows = (win * n_in + keep.unsqueeze(0)).reshape(-1).to(torch.int32).contiguous()
            out.append((rows[0], rows[1], keep_rows))
            n_in = keep.numel()
        if len(cache) > 8:
            cache.clear()
        cache[key] = out
        return out

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
