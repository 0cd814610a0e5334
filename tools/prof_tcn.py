#!/usr/bin/env python
"""The frozen TCN's dependency cone at ltown_dp256's shape (256 segments x 36 windows), for ncu:  python tools/prof_tcn.py"""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(REPO))
from leak_det_gnn_b200.models import NormalPredictorTCN  # noqa: E402

torch.manual_seed(5)
m = NormalPredictorTCN(29, 9).eval().cuda()
x = torch.randn(9216, 36, 29, device="cuda")
t = torch.randn(9216, 36, 9, device="cuda")
for _ in range(2):
    y = m.forward_last(x, t)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
