"""ctypes loader for libfwgpu.so (C ABI declared in include/fwgpu.h).

The library is the product; there is no Python or CPU fallback.  Importing
this module never needs a GPU (the .so loads and exports its symbols on a
CPU-only box); every compute entry point fails loudly with FwError when no
CUDA device is present or the library has not been built.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FWGPU_LIB") or os.path.join(_HERE, "libfwgpu.so")   # FWGPU_LIB: experiment builds

FW_OK = 0
FW_ERR_INVALID = -1
FW_ERR_CUDA = -2
FW_ERR_DOMAIN = -3
FW_ERR_NOMEM = -4
FW_ERR_CAP = -5
FW_TILE = 128

# every symbol include/fwgpu.h declares (tests check the .so exports them all)
SYMBOLS = (
    "fw_version", "fw_last_error", "fw_device_count", "fw_ctx_create", "fw_ctx_destroy",
    "fw_ctx_set_stream", "fw_ctx_last_launches", "fw_solve", "fw_solve_device", "fw_solve_batched",
    "fw_solve_batched_device", "fw_ctx_synchronize", "fw_ctx_set_profiling", "fw_ctx_phase_ms",
    "fw_ctx_phase_spans",
    "fw_multi_create", "fw_multi_unique_id", "fw_multi_create_rank", "fw_multi_destroy", "fw_multi_last_error",
    "fw_multi_sync", "fw_multi_optimum", "fw_multi_download", "fw_multi_solve_edges", "fw_multi_solve",
    "fw_multi_alloc", "fw_multi_upload", "fw_multi_solve_resident", "fw_multi_local_shards", "fw_multi_shard",
    "fw_multi_last_solve_ms", "fw_multi_set_profiling", "fw_multi_phase_ms", "fw_multi_record_row_snapshots",
    "fw_multi_download_sink", "fw_multi_plan", "fw_multi_resolve", "fw_multi_download_local",
    "fw_multi_download_locals", "fw_multi_transport",
    "fw_paths", "fw_paths_device", "fw_build_matrix_device", "fw_state_create", "fw_state_destroy",
    "fw_state_sync", "fw_state_optimum", "fw_state_download", "fw_solve_edges",
    "fw_ctx_last_error", "fw_solve_device_range", "fw_ctx_set_row_snapshot_sink",
    "fw_tables_create", "fw_tables_destroy", "fw_tables_paths",
)


class PlanOp(ctypes.Structure):
    """fw_plan_op (include/fwgpu.h): one operation of the multi-GPU schedule."""
    _fields_ = [(f, ctypes.c_int32) for f in ("kind", "rank", "lane", "b0", "nb", "buf", "row_lo", "row_n",
                                              "ex_lo", "ex_n", "grp_lo")]


class ShardInfo(ctypes.Structure):
    """fw_shard_info (include/fwgpu.h)."""
    _fields_ = [(f, ctypes.c_int32) for f in ("device", "rank", "world", "rows", "n_padded", "cyclic_rows", "group")] + \
               [("ld", ctypes.c_int64), ("d_rate", ctypes.c_void_p), ("d_next", ctypes.c_void_p)]


OP_PIVOT, OP_APPLY, OP_BCAST, OP_A_DONE, OP_WAIT_A, OP_B_DONE, OP_WAIT_B = 1, 2, 3, 4, 5, 6, 7


class FwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fwgpu error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def load():
    """Load libfwgpu.so; raises FwError if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FwError(FW_ERR_CUDA, f"{LIB_PATH} is missing: run __graft_entry__.build() "
                                   "(make -C floydwarshall_b200/csrc); there is no CPU fallback")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64
    L.fw_version.restype = ctypes.c_char_p
    L.fw_last_error.restype = ctypes.c_char_p
    L.fw_device_count.restype = ctypes.c_int
    L.fw_ctx_create.restype = ctypes.c_int
    L.fw_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.fw_ctx_destroy.restype = None
    L.fw_ctx_destroy.argtypes = [vp]
    L.fw_ctx_set_stream.restype = ctypes.c_int
    L.fw_ctx_set_stream.argtypes = [vp, vp, ctypes.c_int]
    L.fw_ctx_last_launches.restype = i64
    L.fw_ctx_last_launches.argtypes = [vp]
    L.fw_ctx_synchronize.restype = ctypes.c_int
    L.fw_ctx_synchronize.argtypes = [vp]
    L.fw_ctx_set_profiling.restype = ctypes.c_int
    L.fw_ctx_set_profiling.argtypes = [vp, ctypes.c_int]
    L.fw_ctx_phase_ms.restype = ctypes.c_int
    L.fw_ctx_phase_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    L.fw_ctx_phase_spans.restype = i64
    L.fw_ctx_phase_spans.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double), i64]
    L.fw_multi_create.restype = ctypes.c_int
    L.fw_multi_create.argtypes = [i32, vp, ctypes.POINTER(vp)]
    L.fw_multi_unique_id.restype = ctypes.c_int
    L.fw_multi_unique_id.argtypes = [vp]
    L.fw_multi_create_rank.restype = ctypes.c_int
    L.fw_multi_create_rank.argtypes = [i32, i32, i32, vp, ctypes.POINTER(vp)]
    L.fw_multi_destroy.restype = None
    L.fw_multi_destroy.argtypes = [vp]
    L.fw_multi_last_error.restype = ctypes.c_char_p
    L.fw_multi_last_error.argtypes = [vp]
    L.fw_multi_sync.restype = ctypes.c_int
    L.fw_multi_sync.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32]
    L.fw_multi_optimum.restype = ctypes.c_int
    L.fw_multi_optimum.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double), vp, i32, ctypes.POINTER(i32)]
    L.fw_multi_download.restype = ctypes.c_int
    L.fw_multi_download.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.fw_multi_solve_edges.restype = ctypes.c_int
    L.fw_multi_solve_edges.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fw_multi_solve.restype = ctypes.c_int
    L.fw_multi_solve.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.fw_multi_alloc.restype = ctypes.c_int
    L.fw_multi_alloc.argtypes = [vp, i32, i32]
    L.fw_multi_upload.restype = ctypes.c_int
    L.fw_multi_upload.argtypes = [vp, i32, i32, vp, vp]
    L.fw_multi_solve_resident.restype = ctypes.c_int
    L.fw_multi_solve_resident.argtypes = [vp]
    L.fw_multi_local_shards.restype = i32
    L.fw_multi_local_shards.argtypes = [vp]
    L.fw_multi_shard.restype = ctypes.c_int
    L.fw_multi_shard.argtypes = [vp, i32, ctypes.POINTER(ShardInfo)]
    L.fw_multi_last_solve_ms.restype = ctypes.c_int
    L.fw_multi_last_solve_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    L.fw_multi_set_profiling.restype = ctypes.c_int
    L.fw_multi_set_profiling.argtypes = [vp, i32]
    L.fw_multi_phase_ms.restype = ctypes.c_int
    L.fw_multi_phase_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i64)]
    L.fw_multi_record_row_snapshots.restype = ctypes.c_int
    L.fw_multi_record_row_snapshots.argtypes = [vp, i32]
    L.fw_multi_download_sink.restype = ctypes.c_int
    L.fw_multi_download_sink.argtypes = [vp, i32, i32, vp]
    L.fw_multi_resolve.restype = ctypes.c_int
    L.fw_multi_resolve.argtypes = [vp]
    L.fw_multi_download_local.restype = ctypes.c_int
    L.fw_multi_download_local.argtypes = [vp, i32, vp, vp]
    L.fw_multi_download_locals.restype = ctypes.c_int
    L.fw_multi_download_locals.argtypes = [vp, ctypes.POINTER(vp), ctypes.POINTER(vp)]
    L.fw_multi_transport.restype = ctypes.c_char_p
    L.fw_multi_transport.argtypes = [vp]
    L.fw_multi_plan.restype = i64
    L.fw_multi_plan.argtypes = [i32, i32, i32, i32, i32, ctypes.POINTER(PlanOp), i64]
    L.fw_paths.restype = ctypes.c_int
    L.fw_paths.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, i64]
    L.fw_paths_device.restype = ctypes.c_int
    L.fw_paths_device.argtypes = [vp, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, i64]
    L.fw_build_matrix_device.restype = ctypes.c_int
    L.fw_build_matrix_device.argtypes = [vp, i32, i64, vp, i32, vp, vp, vp, vp, vp]
    L.fw_state_create.restype = ctypes.c_int
    L.fw_state_create.argtypes = [vp, ctypes.POINTER(vp)]
    L.fw_state_destroy.restype = None
    L.fw_state_destroy.argtypes = [vp]
    L.fw_state_sync.restype = ctypes.c_int
    L.fw_state_sync.argtypes = [vp, i32, vp, i32, vp, vp, vp]
    L.fw_state_optimum.restype = ctypes.c_int
    L.fw_state_optimum.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double), vp, i32,
                                   ctypes.POINTER(i32)]
    L.fw_state_download.restype = ctypes.c_int
    L.fw_state_download.argtypes = [vp, vp, vp]
    L.fw_solve_edges.restype = ctypes.c_int
    L.fw_solve_edges.argtypes = [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fw_ctx_last_error.restype = ctypes.c_char_p
    L.fw_ctx_last_error.argtypes = [vp]
    L.fw_solve_device_range.restype = ctypes.c_int
    L.fw_solve_device_range.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp, i32, i32]
    L.fw_ctx_set_row_snapshot_sink.restype = ctypes.c_int
    L.fw_ctx_set_row_snapshot_sink.argtypes = [vp, vp, i64]
    L.fw_tables_create.restype = ctypes.c_int
    L.fw_tables_create.argtypes = [vp, i32, vp, vp, vp, vp, ctypes.POINTER(vp)]
    L.fw_tables_destroy.restype = None
    L.fw_tables_destroy.argtypes = [vp]
    L.fw_tables_paths.restype = ctypes.c_int
    L.fw_tables_paths.argtypes = [vp, i32, vp, vp, vp, i64]
    L.fw_solve.restype = ctypes.c_int
    L.fw_solve.argtypes = [vp, i32, vp, vp, vp, vp, vp]
    L.fw_solve_device.restype = ctypes.c_int
    L.fw_solve_device.argtypes = [vp, i32, i64, vp, vp, vp, vp, vp]
    L.fw_solve_batched.restype = ctypes.c_int
    L.fw_solve_batched.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    L.fw_solve_batched_device.restype = ctypes.c_int
    L.fw_solve_batched_device.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp]
    _lib = L
    return L


def check(rc: int, ctx_handle=None):
    """Raise FwError for a non-zero status; the text comes from the context the call was made on
    when its handle is given, else from this OS thread's last error."""
    if rc != FW_OK:
        L = load()
        msg = L.fw_ctx_last_error(ctx_handle) if ctx_handle else L.fw_last_error()   # Python stays on one OS thread
        raise FwError(rc, (msg or b"").decode("utf-8", "replace"))


class Context:
    """Owns one fw_ctx (device, stream, workspace)."""

    def __init__(self, device: int = 0):
        L = load()
        h = ctypes.c_void_p()
        check(L.fw_ctx_create(device, ctypes.byref(h)))
        self._h = h
        self.device = device

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream: int | None):
        """cuda_stream: a cudaStream_t handle as int (0 = legacy default stream), or None for
        the context's own stream."""
        if cuda_stream is None:
            check(load().fw_ctx_set_stream(self._h, None, 0))
        else:
            check(load().fw_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream), 1))

    def synchronize(self):
        check(load().fw_ctx_synchronize(self._h))

    def set_profiling(self, on: bool):
        check(load().fw_ctx_set_profiling(self._h, 1 if on else 0))

    def phase_ms(self):
        """(ms[4], count[4]) of the last solve: tile, column panel, row panel, bulk."""
        ms = (ctypes.c_double * 4)()
        cnt = (ctypes.c_int64 * 4)()
        check(load().fw_ctx_phase_ms(self._h, ms, cnt))
        return list(ms), list(cnt)

    def phase_spans(self, phase: int):
        """Per-launch ms of one phase (0 tile, 1 col panel, 2 row panel, 3 bulk) of the last solve."""
        L = load()
        n = int(L.fw_ctx_phase_spans(self._h, phase, None, 0))
        if n < 0:
            check(n)
        buf = (ctypes.c_double * max(n, 1))()
        L.fw_ctx_phase_spans(self._h, phase, buf, n)
        return list(buf)[:n]

    @property
    def last_launches(self) -> int:
        return int(load().fw_ctx_last_launches(self._h))

    def close(self):
        if self._h:
            load().fw_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
