"""ctypes wrapper of fw_paths: exact `_path` expansion on the device."""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib


def expand(init_next: np.ndarray, mid: np.ndarray, csT: np.ndarray, rs: np.ndarray,
           pairs: Sequence[Tuple[int, int]], ctx=None, cap: int = 0) -> List[List[int]]:
    """Index paths (start excluded, destination included) for (src, dst) pairs."""
    n = init_next.shape[0]
    nq = len(pairs)
    if nq == 0:
        return []
    q = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(nq, 2))
    tabs = [np.ascontiguousarray(t, dtype=np.int32) for t in (init_next, mid, csT, rs)]
    offsets = np.zeros(nq + 1, dtype=np.int64)
    cap = cap or max(64, 16 * nq)
    L = _lib.load()
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    for _ in range(2):
        verts = np.empty(cap, dtype=np.int32)
        rc = L.fw_paths(ctx.handle if ctx else None, n, vp(tabs[0]), vp(tabs[1]), vp(tabs[2]), vp(tabs[3]),
                        nq, vp(q), vp(offsets), vp(verts), cap)
        if rc == _lib.FW_ERR_CAP and offsets[-1] > cap:
            cap = int(offsets[-1])        # offsets carry the needed total: retry once with room
            continue
        _lib.check(rc)
        return [verts[offsets[i]:offsets[i + 1]].tolist() for i in range(nq)]
    _lib.check(rc)
    return []


class DeviceTables:
    """The four tables of one optimised matrix, uploaded ONCE (fw_tables_create) and queried many times
    (fw_tables_paths): what a `Matrix RateEntry` with lazily expanded `_path` fields needs."""

    def __init__(self, init_next: np.ndarray, mid: np.ndarray, csT: np.ndarray, rs: np.ndarray, ctx=None):
        self.n = init_next.shape[0]
        tabs = [np.ascontiguousarray(t, dtype=np.int32) for t in (init_next, mid, csT, rs)]
        L = _lib.load()
        self._ctxh = ctx.handle if ctx else None
        h = ctypes.c_void_p()
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        _lib.check(L.fw_tables_create(self._ctxh, self.n, vp(tabs[0]), vp(tabs[1]), vp(tabs[2]), vp(tabs[3]),
                                      ctypes.byref(h)), self._ctxh)
        self._h = h

    def expand(self, pairs: Sequence[Tuple[int, int]], cap: int = 0) -> List[List[int]]:
        nq = len(pairs)
        if nq == 0:
            return []
        q = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32).reshape(nq, 2))
        offsets = np.zeros(nq + 1, dtype=np.int64)
        cap = cap or max(64, 16 * nq)
        L = _lib.load()
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        for _ in range(2):
            verts = np.empty(cap, dtype=np.int32)
            rc = L.fw_tables_paths(self._h, nq, vp(q), vp(offsets), vp(verts), cap)
            if rc == _lib.FW_ERR_CAP and offsets[-1] > cap:
                cap = int(offsets[-1])
                continue
            _lib.check(rc, self._ctxh)
            return [verts[offsets[i]:offsets[i + 1]].tolist() for i in range(nq)]
        _lib.check(rc, self._ctxh)
        return []

    def close(self):
        if self._h:
            _lib.load().fw_tables_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
