"""ORACLE -- test infrastructure only.  ctypes front for oracle/spmm_ref.c."""
from __future__ import annotations

import ctypes
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libltgnn_oracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "spmm_ref.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B", "_build/libltgnn_oracle.so"], check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


def _lib() -> ctypes.CDLL:
    lib = ctypes.CDLL(str(build()))
    lib.ltgnn_oracle_spmm_f32.restype = None
    lib.ltgnn_oracle_scatter_add_f32.restype = None
    return lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


def spmm(rowptr: np.ndarray, col: np.ndarray, val: np.ndarray, x: np.ndarray) -> np.ndarray:
    """x: (B, N, D) float32 -> (B, N, D) float32, sequential fp32 mul+add per entry."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b, n, d = x.shape
    y = np.empty_like(x)
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int32)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float32)
    _lib().ltgnn_oracle_spmm_f32(ctypes.c_int64(b), ctypes.c_int32(n), ctypes.c_int32(d), _p(rowptr), _p(col),
                                 _p(val), _p(x), _p(y))
    return y


def scatter_add(row: np.ndarray, col: np.ndarray, norm: np.ndarray, x: np.ndarray) -> np.ndarray:
    """x: (n_nodes, D) -> (n_nodes, D); literal edge-order scatter-add."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n, d = x.shape
    y = np.empty_like(x)
    row = np.ascontiguousarray(row, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int64)
    norm = np.ascontiguousarray(norm, dtype=np.float32)
    _lib().ltgnn_oracle_scatter_add_f32(ctypes.c_int64(row.shape[0]), ctypes.c_int32(n), ctypes.c_int32(d), _p(row),
                                        _p(col), _p(norm), _p(x), _p(y))
    return y
