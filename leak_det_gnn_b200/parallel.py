"""Data parallelism over window batches: one process per GPU, one flat-bucket gradient all-reduce.

The reference is single-process (models/train_detector.py:161); the path shards naturally because every
window-graph is independent.  Each rank holds the full model (60 418 parameters =
242 KB of gradients) and the CSR/CSC of the pipe graph; the only exchange is the gradient average, done as
ONE collective on a contiguous bucket: every ``p.grad`` is a view into the bucket, so backward kernels write
their results straight into it and no pack/unpack copy runs.  ``clip_grad_norm_`` must be applied AFTER
``allreduce()`` (reference order: backward -> clip -> step, train_detector.py:314-317, on the averaged grads).

Works with any ``torch.distributed`` backend: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist

__all__ = ["FlatGradBucket", "shard_indices", "broadcast_parameters"]


class FlatGradBucket:
    """Owns one contiguous fp32 buffer; ``param.grad`` of every parameter is a view into it."""

    def __init__(self, params: Iterable[torch.nn.Parameter]) -> None:
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev, dt = self.params[0].device, self.params[0].dtype
        if any(p.device != dev or p.dtype != dt for p in self.params):
            raise ValueError("all parameters must share device and dtype")
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self) -> None:
        """Use instead of ``optimizer.zero_grad(set_to_none=True)``: the views must stay attached."""
        self.flat.zero_()

    def attached(self) -> bool:
        base = self.flat.untyped_storage().data_ptr()
        return all(p.grad is not None and p.grad.untyped_storage().data_ptr() == base for p in self.params)

    def allreduce(self, group=None) -> None:
        """Average the gradients over all ranks with a single collective (no-op for world size 1)."""
        if not self.attached():
            raise RuntimeError("a parameter's .grad was replaced (zero_grad(set_to_none=True)?); use bucket.zero()")
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)
        else:  # gloo has no AVG
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))

    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()


def shard_indices(n_items: int, rank: int, world_size: int) -> range:
    """Indices of the deterministic ``seed + idx`` datasets (models/datasets.py:236,490) owned by ``rank``:
    rank r takes r, r + W, r + 2W, ... so W ranks with local batch B reproduce one process with batch W*B."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return range(rank, n_items, world_size)


def broadcast_parameters(params: Sequence[torch.Tensor], src: int = 0, group=None) -> None:
    """Make every rank start from rank ``src``'s weights (one flat broadcast)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([p.detach().reshape(-1) for p in params])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    with torch.no_grad():
        for p in params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n
