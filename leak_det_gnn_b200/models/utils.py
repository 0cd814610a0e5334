"""Mirror of reference models/utils.py: the graph builder (lines 18-166) and the residual builder (lines 169-216)."""
from __future__ import annotations

from typing import Any, Optional

import torch

from ..graph import WDNGraph, build_wdn_graph_from_inp, parse_epanet_inp  # noqa: F401

__all__ = ["WDNGraph", "build_wdn_graph_from_inp", "parse_epanet_inp", "build_residual_sequence_from_segment"]


def build_residual_sequence_from_segment(predictor: Any, noisy_seg: torch.Tensor, time_seg: torch.Tensor, l_pred: int,
                                         l_det: int, device: Optional[Any] = None) -> torch.Tensor:
    """Residual sequence for the detector from a (l_pred + l_det)-step segment: same arguments, result and window
    order as reference models/utils.py:169-216 (window kk covers steps [kk, kk + l_pred) and predicts step
    l_pred + kk; the predictor sees the l_det * B windows kk-major; ``residual[b, kk] = noisy[b, l_pred + kk] - y_hat``).

    The reference slices and concatenates l_det copies of every window on the host side of the call; here the windows
    are one strided view of the segment, and a frozen TCN predictor is evaluated only on the dependency cone of its
    last time step (``NormalPredictorTCN.forward_last``: 99 instead of 288 convolution evaluations per window)."""
    squeeze_back = noisy_seg.dim() == 2
    if squeeze_back:
        noisy_seg, time_seg = noisy_seg.unsqueeze(0), time_seg.unsqueeze(0)
    b, seg_len, s = noisy_seg.shape
    assert seg_len == l_pred + l_det, (seg_len, l_pred, l_det)
    if device is not None:
        noisy_seg, time_seg = noisy_seg.to(device), time_seg.to(device)

    def windows(seg: torch.Tensor) -> torch.Tensor:  # (B, l_pred + l_det, C) -> (l_det * B, l_pred, C), kk-major
        w = seg.unfold(1, l_pred, 1)[:, :l_det]      # (B, l_det, C, l_pred) view
        return w.permute(1, 0, 3, 2).reshape(l_det * b, l_pred, seg.shape[-1])

    x, t = windows(noisy_seg), windows(time_seg)
    fast = getattr(predictor, "forward_last", None)
    if fast is not None and not predictor.training and not torch.is_grad_enabled():
        y_hat = fast(x, t)
    else:
        y_hat = predictor(x, t)
    if y_hat.dim() == 3:
        y_hat = y_hat[:, -1, :]
    target = noisy_seg[:, l_pred:, :].transpose(0, 1).reshape(l_det * b, s)
    residual = (target - y_hat).view(l_det, b, s).transpose(0, 1).contiguous()
    return residual.squeeze(0) if squeeze_back else residual
