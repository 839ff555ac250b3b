"""Row-sharded solve across the GPUs of one box (config C5) -- thin ctypes wrapper.

Everything that matters lives behind the C ABI (include/fwgpu.h, fw_multi_*; csrc/fw_multi.cuh): the k-block
schedule (csrc/fw_plan.hpp), the two stream lanes per GPU, the pivot-panel transport (copy engines or NCCL),
buildMatrix per shard, the resident sharded state and the `optimum` read-out across shards.  A Haskell caller
binds the same entry points; this module only adapts them to numpy / torch for tests and
bench.py.

Two ways to run, as in the C ABI:
  MultiSolver(devices=[0, 1, ...])          one process drives all GPUs (the reference's in-process call)
  MultiSolver.for_rank(device, rank, world) one process per GPU (torchrun); the NCCL id is shared through
                                            torch.distributed by `for_torchrun()`.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from . import _lib

B = _lib.FW_TILE


def _vp(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


class MultiSolver:
    """Owns one fw_multi object."""

    def __init__(self, devices: Optional[Sequence[int]] = None, ndev: Optional[int] = None, _handle=None):
        self.L = _lib.load()
        if _handle is not None:
            self._h = _handle
            return
        if devices is None:
            ndev = ndev or max(1, self.L.fw_device_count())   # no device: let the library say so (FW_ERR_CUDA, no fallback)
            arr = None
        else:
            ndev = len(devices)
            arr = (ctypes.c_int32 * ndev)(*devices)
        h = ctypes.c_void_p()
        _lib.check(self.L.fw_multi_create(ndev, arr, ctypes.byref(h)))
        self._h = h

    # ---- construction, one process per GPU
    @classmethod
    def for_rank(cls, device: int, rank: int, world: int, nccl_id: bytes) -> "MultiSolver":
        L = _lib.load()
        h = ctypes.c_void_p()
        buf = ctypes.create_string_buffer(nccl_id, 128) if world > 1 else None
        _lib.check(L.fw_multi_create_rank(device, rank, world, buf, ctypes.byref(h)))
        return cls(_handle=h)

    @classmethod
    def for_torchrun(cls, device: int) -> "MultiSolver":
        """Rank / world from torch.distributed (already initialised); rank 0's NCCL id goes round by broadcast."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0 and world > 1:
            raw = ctypes.create_string_buffer(128)
            _lib.check(_lib.load().fw_multi_unique_id(raw))
            ident = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        if world > 1:
            t = ident.to(torch.device("cuda", device))
            dist.broadcast(t, src=0)
            ident = t.cpu()
        return cls.for_rank(device, rank, world, bytes(ident.numpy().tobytes()))

    def _check(self, rc: int):
        if rc != _lib.FW_OK:
            msg = self.L.fw_multi_last_error(self._h) or self.L.fw_last_error()
            raise _lib.FwError(rc, (msg or b"").decode("utf-8", "replace"))

    # ---- floydWarshall (Algorithms.hs:19-20) on all GPUs
    def sync(self, n: int, ccy, src, dst, val, paths: bool = True):
        """syncMatrix on an OutSync state: map in COO form in, optimised matrix stays sharded in HBM."""
        ccy = np.ascontiguousarray(ccy, dtype=np.int32)
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        self._check(self.L.fw_multi_sync(self._h, n, _vp(ccy), len(src), _vp(src), _vp(dst), _vp(val), 1 if paths else 0))
        self.n = n

    def resolve(self):
        """The last sync() again from the COO still on the devices (buildMatrix + solve, no PCIe traffic)."""
        self._check(self.L.fw_multi_resolve(self._h))

    def download_local(self, i: int, rate_ptr: int, next_ptr: int):
        """Local shard i (rows x n_padded, local row order) into host buffers given by address."""
        self._check(self.L.fw_multi_download_local(self._h, i, ctypes.c_void_p(rate_ptr) if rate_ptr else None,
                                                   ctypes.c_void_p(next_ptr) if next_ptr else None))

    def download_locals(self, rate_ptrs, next_ptrs):
        """Every local shard into its host buffer (addresses), all devices copying at once."""
        k = len(rate_ptrs)
        ra = (ctypes.c_void_p * k)(*rate_ptrs)
        xa = (ctypes.c_void_p * k)(*next_ptrs)
        self._check(self.L.fw_multi_download_locals(self._h, ra, xa))

    def transport(self) -> str:
        return (self.L.fw_multi_transport(self._h) or b"").decode()

    def optimum(self, src: int, dst: int, cap: int = 4096):
        """(rate, index path) of one entry, read across the shards (Algorithms.hs:74-75)."""
        rate = ctypes.c_double()
        plen = ctypes.c_int32()
        while True:
            path = np.empty(cap, dtype=np.int32)
            rc = self.L.fw_multi_optimum(self._h, src, dst, ctypes.byref(rate), _vp(path), cap, ctypes.byref(plen))
            if rc == _lib.FW_ERR_CAP and plen.value > cap:
                cap = plen.value
                continue
            self._check(rc)
            return rate.value, path[:plen.value].tolist()

    def solve_dense(self, rate: np.ndarray, nxt: np.ndarray, paths: bool = False):
        """runAlgo on host matrices (copies): rows go to their shards, the solve runs, rows come back."""
        from .dense import DenseResult
        n = rate.shape[0]
        r = np.ascontiguousarray(rate, dtype=np.float64).copy()
        x = np.ascontiguousarray(nxt, dtype=np.int32).copy()
        mid, csT, rs = ((np.empty((n, n), dtype=np.int32) for _ in range(3)) if paths else (None, None, None))
        self._check(self.L.fw_multi_solve(self._h, n, _vp(r), _vp(x), _vp(mid), _vp(csT), _vp(rs)))
        self.n = n
        return DenseResult(r, x, mid, csT, rs)

    def solve_edges(self, n: int, ccy, src, dst, val, paths: bool = False, out_rate=None, out_next=None):
        """Map in COO form in, dense host matrices out (fw_multi_solve_edges)."""
        from .dense import DenseResult
        ccy = np.ascontiguousarray(ccy, dtype=np.int32)
        src = np.ascontiguousarray(src, dtype=np.int32)
        dst = np.ascontiguousarray(dst, dtype=np.int32)
        val = np.ascontiguousarray(val, dtype=np.float64)
        r = out_rate if out_rate is not None else np.empty((n, n), dtype=np.float64)
        x = out_next if out_next is not None else np.empty((n, n), dtype=np.int32)
        ini, mid, csT, rs = ((np.empty((n, n), dtype=np.int32) for _ in range(4)) if paths else (None,) * 4)
        self._check(self.L.fw_multi_solve_edges(self._h, n, _vp(ccy), len(src), _vp(src), _vp(dst), _vp(val),
                                                _vp(r), _vp(x), _vp(ini), _vp(mid), _vp(csT), _vp(rs)))
        self.n = n
        res = DenseResult(r, x, mid, csT, rs)
        res.init_next = ini
        return res

    # ---- resident workflow (bench, tests)
    def alloc(self, n: int, paths: bool = False):
        self._check(self.L.fw_multi_alloc(self._h, n, 1 if paths else 0))
        self.n = n

    def upload(self, row0: int, rate_rows: np.ndarray, next_rows: np.ndarray):
        assert rate_rows.flags.c_contiguous and next_rows.flags.c_contiguous
        self._check(self.L.fw_multi_upload(self._h, row0, rate_rows.shape[0], _vp(rate_rows), _vp(next_rows)))

    def solve_resident(self):
        self._check(self.L.fw_multi_solve_resident(self._h))

    def download(self, row0: int, rows: int, want=("rate", "next")):
        n = self.n
        out = {}
        for k in ("rate", "next", "init_next", "mid", "csT", "rs"):
            out[k] = np.empty((rows, n), dtype=np.float64 if k == "rate" else np.int32) if k in want else None
        self._check(self.L.fw_multi_download(self._h, row0, rows, _vp(out["rate"]), _vp(out["next"]),
                                             _vp(out["init_next"]), _vp(out["mid"]), _vp(out["csT"]), _vp(out["rs"])))
        return out

    def download_into(self, row0: int, rows: int, rate_ptr: int, next_ptr: int):
        """Row range into caller-owned (e.g. pinned) host buffers given by address."""
        self._check(self.L.fw_multi_download(self._h, row0, rows, ctypes.c_void_p(rate_ptr), ctypes.c_void_p(next_ptr),
                                             None, None, None, None))

    def shards(self) -> List[_lib.ShardInfo]:
        out = []
        for i in range(int(self.L.fw_multi_local_shards(self._h))):
            info = _lib.ShardInfo()
            self._check(self.L.fw_multi_shard(self._h, i, ctypes.byref(info)))
            out.append(info)
        return out

    def last_solve(self):
        """(device ms of the last solve: CUDA events, max over the local shards; kernel launches issued)."""
        ms = ctypes.c_double()
        ln = ctypes.c_int64()
        self._check(self.L.fw_multi_last_solve_ms(self._h, ctypes.byref(ms), ctypes.byref(ln)))
        return ms.value, int(ln.value)

    def set_profiling(self, on: bool):
        self._check(self.L.fw_multi_set_profiling(self._h, 1 if on else 0))

    def phase_ms(self):
        ms = (ctypes.c_double * 4)()
        cnt = (ctypes.c_int64 * 4)()
        self._check(self.L.fw_multi_phase_ms(self._h, ms, cnt))
        return list(ms), list(cnt)

    def record_row_snapshots(self, on: bool):
        self._check(self.L.fw_multi_record_row_snapshots(self._h, 1 if on else 0))

    def download_sink(self, row0: int, rows: int) -> np.ndarray:
        out = np.empty((rows, self.n), dtype=np.float64)
        self._check(self.L.fw_multi_download_sink(self._h, row0, rows, _vp(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self.L.fw_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def global_rows(info: _lib.ShardInfo) -> np.ndarray:
    """Global row index of every local row of a shard (cyclic blocks: include/fwgpu.h, fw_shard_info)."""
    l = np.arange(info.rows, dtype=np.int64)
    return ((l // info.cyclic_rows) * info.world + info.rank) * info.cyclic_rows + l % info.cyclic_rows
