"""Launch accounting for bench.py: how many libltgnn kernels ran, and (opt-in) CUDA-event
timing of each named kernel on the stream it was launched on."""
from __future__ import annotations

from collections import defaultdict
from typing import Dict, List, Tuple

import torch

launches: int = 0
_timing: bool = False
_events: Dict[str, List[Tuple[torch.cuda.Event, torch.cuda.Event]]] = defaultdict(list)


def reset(timing: bool = False) -> None:
    global launches, _timing
    launches = 0
    _timing = timing
    _events.clear()


def begin(name: str):
    """Call right before a kernel launch; returns a token for :func:`end`."""
    global launches
    launches += 1
    if not _timing:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return (name, e0)


def end(token) -> None:
    if token is None:
        return
    e1 = torch.cuda.Event(enable_timing=True)
    e1.record()
    _events[token[0]].append((token[1], e1))


def summary() -> Dict[str, Dict[str, float]]:
    """name -> {count, total_ms, mean_ms}; call after torch.cuda.synchronize()."""
    out = {}
    for name, pairs in _events.items():
        ms = [a.elapsed_time(b) for a, b in pairs]
        out[name] = {"count": len(ms), "total_ms": sum(ms), "mean_ms": sum(ms) / max(len(ms), 1)}
    return out
