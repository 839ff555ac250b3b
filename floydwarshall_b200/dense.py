"""Dense (rate, next) entry points over the C ABI -- numpy host buffers or
torch CUDA tensors.  Thin: all arithmetic happens in libfwgpu.so."""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib


@dataclass
class DenseResult:
    rate: np.ndarray
    next: np.ndarray
    mid: Optional[np.ndarray] = None
    csT: Optional[np.ndarray] = None
    rs: Optional[np.ndarray] = None


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def solve(rate: np.ndarray, nxt: np.ndarray, *, paths: bool = False,
          ctx: Optional[_lib.Context] = None) -> DenseResult:
    """fw_solve on copies of the inputs (host buffers; H2D/D2H inside the call)."""
    n = rate.shape[0]
    assert rate.shape == (n, n) and nxt.shape == (n, n)
    r = np.ascontiguousarray(rate, dtype=np.float64).copy()
    x = np.ascontiguousarray(nxt, dtype=np.int32).copy()
    mid = np.empty((n, n), dtype=np.int32) if paths else None
    csT = np.empty((n, n), dtype=np.int32) if paths else None
    rs = np.empty((n, n), dtype=np.int32) if paths else None
    L = _lib.load()
    _lib.check(L.fw_solve(ctx.handle if ctx else None, n, _vp(r), _vp(x), _vp(mid), _vp(csT), _vp(rs)))
    return DenseResult(r, x, mid, csT, rs)


def solve_inplace(rate: np.ndarray, nxt: np.ndarray, ctx: Optional[_lib.Context] = None):
    """fw_solve directly on caller-owned C-contiguous host buffers (e.g. pinned)."""
    n = rate.shape[0]
    assert rate.flags.c_contiguous and nxt.flags.c_contiguous
    assert rate.dtype == np.float64 and nxt.dtype == np.int32
    L = _lib.load()
    _lib.check(L.fw_solve(ctx.handle if ctx else None, n, _vp(rate), _vp(nxt), None, None, None))


def solve_batched(rate: np.ndarray, nxt: np.ndarray, *, paths: bool = False,
                  ctx: Optional[_lib.Context] = None) -> DenseResult:
    b, n = rate.shape[0], rate.shape[1]
    assert rate.shape == (b, n, n) and nxt.shape == (b, n, n)
    r = np.ascontiguousarray(rate, dtype=np.float64).copy()
    x = np.ascontiguousarray(nxt, dtype=np.int32).copy()
    mid = np.empty((b, n, n), dtype=np.int32) if paths else None
    csT = np.empty((b, n, n), dtype=np.int32) if paths else None
    rs = np.empty((b, n, n), dtype=np.int32) if paths else None
    L = _lib.load()
    _lib.check(L.fw_solve_batched(ctx.handle if ctx else None, b, n, _vp(r), _vp(x), _vp(mid), _vp(csT),
                                  _vp(rs)))
    return DenseResult(r, x, mid, csT, rs)


# ---- torch CUDA tensors (device-resident; plumbing only) -------------------
def solve_device(ctx: _lib.Context, rate_t, next_t, mid_t=None, csT_t=None, rs_t=None):
    """fw_solve_device on torch CUDA tensors, in place, async on ctx's stream."""
    n = rate_t.shape[0]
    assert rate_t.is_cuda and rate_t.is_contiguous() and next_t.is_contiguous()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    L = _lib.load()
    _lib.check(L.fw_solve_device(ctx.handle, n, rate_t.stride(0), p(rate_t), p(next_t), p(mid_t), p(csT_t),
                                 p(rs_t)))


def solve_batched_device(ctx: _lib.Context, rate_t, next_t, mid_t=None, csT_t=None, rs_t=None):
    b, n = rate_t.shape[0], rate_t.shape[1]
    assert rate_t.is_cuda and rate_t.is_contiguous() and next_t.is_contiguous()
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    L = _lib.load()
    _lib.check(L.fw_solve_batched_device(ctx.handle, b, n, p(rate_t), p(next_t), p(mid_t), p(csT_t), p(rs_t)))
