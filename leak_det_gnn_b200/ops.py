"""torch.autograd bindings of the libltgnn kernels.

Everything here runs on CUDA tensors only and calls straight into the C ABI with raw
device pointers and the current torch stream.  Non-CUDA / non-fp32 inputs are a
``ValueError``; a missing extension is a ``RuntimeError`` -- there is no fallback path.
"""
from __future__ import annotations

import ctypes
import threading
from typing import Dict, Optional

import numpy as np
import torch

from . import instrument as _inst
from . import lib as _lib
from .graph import GCNCsr, build_gcn_csr

__all__ = ["PipeGraph", "spmm", "aggregate"]


def _ptr(a: np.ndarray) -> ctypes.c_void_p:
    return a.ctypes.data_as(ctypes.c_void_p)


class PipeGraph:
    """The normalised adjacency of ONE pipe network, uploaded once per device.

    Stands in for what the reference rebuilds on every forward: the B-times replicated
    ``edge_index`` (models/detector.py:105-114,195-196) and PyG's ``gcn_norm`` inside each
    ``GCNConv`` call (models/detector.py:199).
    """

    def __init__(self, edge_index: torch.Tensor, num_nodes: int) -> None:
        self.num_nodes = int(num_nodes)
        self.csr: GCNCsr = build_gcn_csr(edge_index, self.num_nodes)
        self._handles: Dict[int, ctypes.c_void_p] = {}
        self._lock = threading.Lock()

    @property
    def nnz(self) -> int:
        return self.csr.nnz

    def handle(self, device: torch.device) -> ctypes.c_void_p:
        if device.type != "cuda":
            raise ValueError(f"PipeGraph needs a CUDA device, got {device}")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            with self._lock:
                h = self._handles.get(idx)
                if h is None:
                    c = self.csr
                    out = ctypes.c_void_p()
                    L = _lib.load()
                    _lib.check(L.ltgnn_graph_create(idx, c.num_nodes, c.nnz, _ptr(c.rowptr), _ptr(c.col), _ptr(c.val),
                                                    _ptr(c.t_rowptr), _ptr(c.t_col), _ptr(c.t_val), ctypes.byref(out)))
                    self._handles[idx] = h = out
        return h

    def __del__(self) -> None:  # best effort; the handle only owns ~40 KB of device memory
        try:
            L = _lib.load()
            for h in self._handles.values():
                L.ltgnn_graph_destroy(h)
        except Exception:
            pass


def _check_act(x: torch.Tensor, name: str) -> None:
    if not x.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (libltgnn has no CPU path), got {x.device}")
    if x.dtype != torch.float32:
        raise ValueError(f"{name} must be float32, got {x.dtype}")
    if not x.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def _stream(x: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)


def spmm(graph: PipeGraph, x: torch.Tensor, transpose: bool = False, algo: int = _lib.SPMM_AUTO,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Raw (non-differentiable) batched aggregation.  x: (B, N, D) or (B*N, D) fp32 CUDA."""
    _check_act(x, "x")
    n = graph.num_nodes
    d = x.shape[-1]
    if x.numel() % (n * d) != 0:
        raise ValueError(f"x with shape {tuple(x.shape)} is not a whole number of {n}-node graphs")
    b = x.numel() // (n * d)
    y = torch.empty_like(x) if out is None else out
    if out is not None:
        _check_act(out, "out")
        if out.shape != x.shape:
            raise ValueError("out shape mismatch")
    L = _lib.load()
    h = graph.handle(x.device)
    tok = _inst.begin("spmm_bwd" if transpose else "spmm_fwd")
    _lib.check(L.ltgnn_spmm(h, int(bool(transpose)), b, d, x.data_ptr(), y.data_ptr(), int(algo), _stream(x)))
    _inst.end(tok)
    return y


class _Aggregate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, graph: PipeGraph) -> torch.Tensor:
        ctx.graph = graph
        return spmm(graph, x.contiguous(), transpose=False)

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        return spmm(ctx.graph, dy.contiguous(), transpose=True), None


def aggregate(x: torch.Tensor, graph: PipeGraph) -> torch.Tensor:
    """Differentiable ``A_hat @ x`` per window; backward is the CSC-transpose gather."""
    return _Aggregate.apply(x, graph)


def gcn_conv(x: torch.Tensor, graph: PipeGraph, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """``A_hat (x W^T) + b`` -- the operator ``torch_geometric.nn.GCNConv.forward`` computes
    (reference call site models/detector.py:199).  x: (B, N, Din) or (B*N, Din)."""
    n = graph.num_nodes
    xw = torch.nn.functional.linear(x, weight)
    y = aggregate(xw.reshape(-1, n, xw.shape[-1]), graph).view(*x.shape[:-1], xw.shape[-1])
    return y if bias is None else y + bias


def mean_pool(x: torch.Tensor) -> torch.Tensor:
    """(B, N, D) -> (B, D): ``global_mean_pool`` for equal-sized graphs (detector.py:214-215)."""
    return x.mean(dim=1)


def node_init(h_s: torch.Tensor, sensor_idx: torch.Tensor, num_nodes: int, weight: torch.Tensor,
              bias: torch.Tensor) -> torch.Tensor:
    """Node-feature initialisation, reference models/detector.py:178-189:
    ``relu(Linear([h0 | mask]))`` with ``h0`` = zeros except the sensor rows (= h_s) and ``mask`` the
    sensor indicator.  Never materialises the zero-padded (B, N, d_s+1) tensor: a non-sensor row is
    the batch-independent constant ``relu(bias)``; a sensor row is
    ``relu(W[:, :d_s] h_s + W[:, d_s] + bias)``.  Returns (B, N, D)."""
    b, s, ds = h_s.shape
    base = torch.relu(bias)
    x = base.expand(b, num_nodes, base.shape[0]).contiguous()
    sens = torch.relu(torch.nn.functional.linear(h_s, weight[:, :ds], weight[:, ds] + bias))
    x[:, sensor_idx, :] = sens
    return x


def _dev_index(x: torch.Tensor) -> int:
    return x.device.index if x.device.index is not None else torch.cuda.current_device()


def linear_tc(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, relu: bool = False,
              transposed: bool = False, gate: Optional[torch.Tensor] = None, gate_scale: float = 1.0) -> torch.Tensor:
    """Raw (non-differentiable) tensor-core dense layer (tcgen05, 3xTF32):
    ``gate(act(x @ weight.T + bias))``, or ``x @ weight`` when ``transposed``; see ltgnn_linear."""
    _check_act(x, "x")
    _check_act(weight, "weight")
    k = x.shape[-1]
    n = weight.shape[1] if transposed else weight.shape[0]
    if (weight.shape[0] if transposed else weight.shape[1]) != k:
        raise ValueError(f"weight {tuple(weight.shape)} does not match x[..., {k}] (transposed={transposed})")
    m = x.numel() // k
    y = torch.empty(*x.shape[:-1], n, device=x.device, dtype=torch.float32)
    if gate is not None:
        _check_act(gate, "gate")
        if gate.numel() != y.numel():
            raise ValueError("gate must have the shape of the output")
    L = _lib.load()
    tok = _inst.begin("linear_tc")
    _lib.check(L.ltgnn_linear(_dev_index(x), m, k, n, x.data_ptr(), weight.data_ptr(), int(transposed),
                              None if bias is None else bias.data_ptr(), int(relu),
                              None if gate is None else gate.data_ptr(), float(gate_scale), y.data_ptr(), _stream(x)))
    _inst.end(tok)
    return y


def new_dropout_seed() -> int:
    """A fresh 63-bit key for one in-kernel dropout stream, drawn from torch's CPU generator so that
    ``torch.manual_seed`` makes training runs reproducible."""
    return int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())


def spmm_fused(graph: PipeGraph, x: torch.Tensor, transpose: bool = False, bias: Optional[torch.Tensor] = None,
               relu: bool = False, drop_p: float = 0.0, drop_seed: int = 0, gate: Optional[torch.Tensor] = None,
               gate_scale: float = 1.0, want_colsum: bool = False):
    """Raw fused aggregation (see ltgnn_spmm_fused): ``dropout(relu(A (x * gatemask) + bias))``.
    Returns ``y`` or ``(y, colsum)`` when ``want_colsum`` (column sums of the gated input)."""
    _check_act(x, "x")
    n, d = graph.num_nodes, x.shape[-1]
    if x.numel() % (n * d) != 0:
        raise ValueError(f"x with shape {tuple(x.shape)} is not a whole number of {n}-node graphs")
    b = x.numel() // (n * d)
    if gate is not None:
        _check_act(gate, "gate")
        if gate.shape != x.shape:
            raise ValueError("gate must have the shape of x")
    if bias is not None:
        _check_act(bias, "bias")
    y = torch.empty_like(x)
    L = _lib.load()
    h = graph.handle(x.device)
    colsum = ws = None
    if want_colsum:
        if gate is None:
            raise ValueError("want_colsum needs a gate")
        colsum = torch.empty(d, device=x.device, dtype=torch.float32)
        ws = torch.empty(int(L.ltgnn_spmm_ws_floats(h)), device=x.device, dtype=torch.float32)
    tok = _inst.begin("spmm_fused_bwd" if transpose else "spmm_fused_fwd")
    _lib.check(L.ltgnn_spmm_fused(h, int(bool(transpose)), b, d, x.data_ptr(), y.data_ptr(),
                                  None if bias is None else bias.data_ptr(), int(relu), float(drop_p),
                                  int(drop_seed) & (2**64 - 1), None if gate is None else gate.data_ptr(),
                                  float(gate_scale), None if colsum is None else colsum.data_ptr(),
                                  None if ws is None else ws.data_ptr(), _stream(x)))
    _inst.end(tok)
    return (y, colsum) if want_colsum else y
