"""CPU oracle for the matrix-optimisation path of jinilover/floydWarshall.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; nothing under
floydwarshall_b200/ does.  Parity status: PINNED against the reference's own
golden vectors (tests/golden/reference_vectors.json; see fw_oracle.c header).

Two layers:

* A *literal* pure-Python twin of the reference functions with materialised
  `_path` lists -- `build_matrix`, `run_algo`, `floyd_warshall`, `optimum` --
  for tiny graphs (O(N^3) Python loops).  It follows, line by line,
  /root/reference/src/lib/Algorithms.hs:19-78 and Utils.hs:13-14.
* ctypes bindings to `libfworacle.so` (fw_oracle.c: the same loop on the dense
  rate/next encoding, optionally OpenMP) for sizes up to a few thousand, plus
  the exact-path reconstruction from (mid, csT, rs) of SURVEY.md 7.4.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfworacle.so")


# --------------------------------------------------------------------------
# Literal twin (reference types: src/lib/Types.hs:13-39)
# --------------------------------------------------------------------------
@dataclass(frozen=True, order=True)
class Vertex:
    """Types.hs:13-17 -- derives (Ord, Eq): exchange-major, then currency."""
    exch: str
    ccy: str

    def show(self) -> str:                       # Types.hs:19-20
        return f"({self.exch}, {self.ccy})"


@dataclass
class RateEntry:
    """Types.hs:24-29"""
    best_rate: float
    start: Vertex
    path: List[Vertex] = field(default_factory=list)


def isolated_entry(start: Vertex) -> RateEntry:  # Utils.hs:13-14
    return RateEntry(0.0, start, [])


def sorted_vertices(ex_rates: Dict[Tuple[Vertex, Vertex], float]) -> List[Vertex]:
    """Algorithms.hs:29  V.fromList . sort . nub $ keys >>= \\(k1,k2) -> [k1,k2]"""
    seen = set()
    for (k1, k2) in ex_rates.keys():
        seen.add(k1)
        seen.add(k2)
    return sorted(seen)


def build_matrix(ex_rates: Dict[Tuple[Vertex, Vertex], float]) -> List[List[RateEntry]]:
    """Algorithms.hs:26-40"""
    vertices = sorted_vertices(ex_rates)
    n = len(vertices)
    matrix = []
    for i in range(n):
        vtx_i = vertices[i]
        row = []
        for j in range(n):
            vtx_j = vertices[j]
            if i == j:                                            # :33
                row.append(isolated_entry(vtx_i))
            elif vtx_i.ccy == vtx_j.ccy:                          # :34
                row.append(RateEntry(1.0, vtx_i, [vtx_j]))
            elif (vtx_i, vtx_j) in ex_rates:                      # :35-36
                row.append(RateEntry(ex_rates[(vtx_i, vtx_j)], vtx_i, [vtx_j]))
            else:                                                 # :37
                row.append(isolated_entry(vtx_i))
        matrix.append(row)
    return matrix


def run_algo(matrix: List[List[RateEntry]]) -> List[List[RateEntry]]:
    """Algorithms.hs:42-61 -- one fresh generation per k, reads from the old one."""
    n = len(matrix)
    for k in range(n):
        new_matrix = []
        for i in range(n):
            if i == k:                                            # :50
                new_matrix.append(matrix[k])
                continue
            new_row = []
            for j in range(n):
                orig = matrix[i][j]
                if j == i or j == k:                              # :54
                    new_row.append(orig)
                    continue
                ik = matrix[i][k]
                kj = matrix[k][j]
                new_rate = ik.best_rate * kj.best_rate            # :61
                if orig.best_rate < new_rate:                     # :55
                    new_row.append(RateEntry(new_rate, orig.start, ik.path + kj.path))
                else:
                    new_row.append(orig)
            new_matrix.append(new_row)
        matrix = new_matrix
    return matrix


def floyd_warshall(ex_rates: Dict[Tuple[Vertex, Vertex], float]) -> List[List[RateEntry]]:
    """Algorithms.hs:19-20"""
    return run_algo(build_matrix(ex_rates))


class AlgoOptimumError(Exception):
    """Types.hs:62-63 (AlgoOptimumError Text)"""


def optimum(src: Vertex, dest: Vertex, matrix: Sequence[Sequence[RateEntry]]) -> RateEntry:
    """Algorithms.hs:65-78"""
    starts = []
    for row in matrix:                                            # :70 traverse (!? 0)
        if len(row) == 0:
            raise AlgoOptimumError("The matrix is empty")
        starts.append(row[0].start)

    def vertice_idx(v: Vertex) -> int:                            # :77
        try:
            return starts.index(v)
        except ValueError:
            raise AlgoOptimumError(f"{v.show()} is not entered before") from None

    src_idx = vertice_idx(src)
    dest_idx = vertice_idx(dest)
    not_reachable = f"There is no exchange between {src.show()} and {dest.show()}"
    if dest_idx >= len(matrix[src_idx]):
        raise AlgoOptimumError(not_reachable)
    entry = matrix[src_idx][dest_idx]
    if not entry.path:                                            # :75
        raise AlgoOptimumError(not_reachable)
    return entry


def dense_from_entries(matrix: Sequence[Sequence[RateEntry]]):
    """RateEntry matrix -> (vertices, rate f64[n,n], next i32[n,n], paths as index lists)."""
    n = len(matrix)
    vertices = [row[0].start for row in matrix]
    index = {v: i for i, v in enumerate(vertices)}
    rate = np.zeros((n, n), dtype=np.float64)
    nxt = np.full((n, n), -1, dtype=np.int32)
    paths = [[[] for _ in range(n)] for _ in range(n)]
    for i in range(n):
        for j in range(n):
            e = matrix[i][j]
            rate[i, j] = e.best_rate
            paths[i][j] = [index[v] for v in e.path]
            if e.path:
                nxt[i, j] = index[e.path[0]]
    return vertices, rate, nxt, paths


# --------------------------------------------------------------------------
# Dense C oracle (fw_oracle.c)
# --------------------------------------------------------------------------
def build_lib(force: bool = False) -> str:
    """Compile fw_oracle.c -> libfworacle.so (gcc -O2 -ffp-contract=off -fopenmp)."""
    src = os.path.join(_HERE, "fw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(
            ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC",
             "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_lib())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int32)
        L.fw_oracle_run_generations.restype = ctypes.c_int64
        L.fw_oracle_run_generations.argtypes = [ctypes.c_int32, dp, ip, ip, ip, ip]
        L.fw_oracle_run_inplace.restype = ctypes.c_int64
        L.fw_oracle_run_inplace.argtypes = [ctypes.c_int32, dp, ip, ip, ip, ip, ctypes.c_int32]
        L.fw_oracle_run_ksteps.restype = ctypes.c_int64
        L.fw_oracle_run_ksteps.argtypes = [ctypes.c_int32, dp, ip, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_int32]
        L.fw_oracle_max_threads.restype = ctypes.c_int32
        L.fw_oracle_run_batched.restype = ctypes.c_int64
        L.fw_oracle_run_batched.argtypes = [ctypes.c_int32, ctypes.c_int32, dp, ip, ip, ip, ip, ctypes.c_int32]
        L.fw_oracle_replay_rows.restype = ctypes.c_int64
        L.fw_oracle_replay_rows.argtypes = [ctypes.c_int32, ctypes.c_int32, ip, dp, ctypes.c_int64, dp, ip, ip, ip,
                                            dp, ip, ctypes.c_int32]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if a is not None else None


@dataclass
class DenseResult:
    rate: np.ndarray
    next: np.ndarray
    mid: Optional[np.ndarray] = None
    csT: Optional[np.ndarray] = None
    rs: Optional[np.ndarray] = None
    updates: int = 0


def solve_dense(rate: np.ndarray, nxt: np.ndarray, *, paths: bool = False,
                literal: bool = False, threads: int = 1) -> DenseResult:
    """Run the oracle on a dense (rate, next) pair; inputs are not modified."""
    n = rate.shape[0]
    assert rate.shape == (n, n) and nxt.shape == (n, n)
    r = np.ascontiguousarray(rate, dtype=np.float64).copy()
    x = np.ascontiguousarray(nxt, dtype=np.int32).copy()
    mid = np.empty((n, n), dtype=np.int32) if paths else None
    csT = np.empty((n, n), dtype=np.int32) if paths else None
    rs = np.empty((n, n), dtype=np.int32) if paths else None
    if n == 0:
        return DenseResult(r, x, mid, csT, rs, 0)
    if literal:
        u = lib().fw_oracle_run_generations(n, _dp(r), _ip(x), _ip(mid), _ip(csT), _ip(rs))
    else:
        u = lib().fw_oracle_run_inplace(n, _dp(r), _ip(x), _ip(mid), _ip(csT), _ip(rs), threads)
    if u < 0:
        raise MemoryError("fw_oracle: allocation failed")
    return DenseResult(r, x, mid, csT, rs, int(u))


def run_ksteps(rate: np.ndarray, nxt: np.ndarray, k0: int, k1: int, threads: int = 0) -> int:
    """In-place: apply steps k0..k1-1 only (bounded CPU-baseline sample)."""
    n = rate.shape[0]
    assert rate.flags.c_contiguous and nxt.flags.c_contiguous
    return int(lib().fw_oracle_run_ksteps(n, _dp(rate), _ip(nxt), k0, k1, threads))


def solve_batched(rate: np.ndarray, nxt: np.ndarray, *, paths: bool = False, threads: int = 0) -> DenseResult:
    """`batch` independent graphs [batch, n, n], each through the reference loop (OpenMP over graphs)."""
    b, n = rate.shape[0], rate.shape[1]
    assert rate.shape == (b, n, n) and nxt.shape == (b, n, n)
    r = np.ascontiguousarray(rate, dtype=np.float64).copy()
    x = np.ascontiguousarray(nxt, dtype=np.int32).copy()
    mid = np.empty((b, n, n), dtype=np.int32) if paths else None
    csT = np.empty((b, n, n), dtype=np.int32) if paths else None
    rs = np.empty((b, n, n), dtype=np.int32) if paths else None
    u = lib().fw_oracle_run_batched(b, n, _dp(r), _ip(x), _ip(mid), _ip(csT), _ip(rs), threads)
    return DenseResult(r, x, mid, csT, rs, int(u))


@dataclass
class RowReplay:
    rows: np.ndarray       # sampled row indices
    rate: np.ndarray       # [nrows, n] final rows
    next: np.ndarray
    mid: np.ndarray
    csT: np.ndarray        # [nrows, n] mid[i][k] as step k begins
    at_i: np.ndarray       # [nrows, n] row i as step i begins (must equal the recorded pivot row S[i])
    mid_at_i: np.ndarray   # [nrows, n] = rs[i][:]
    updates: int = 0


def replay_rows(rows, S: np.ndarray, init_rate_rows: np.ndarray, init_next_rows: np.ndarray,
                threads: int = 0) -> RowReplay:
    """The reference loop on a SAMPLE of rows, given the recorded pivot rows S[k] = row k as step k begins
    (fw_oracle_replay_rows).  init_*_rows: [nrows, n] initial contents of the sampled rows."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    nr, n = len(rows), S.shape[1]
    assert S.dtype == np.float64 and S.shape[0] >= n and S.strides[1] == 8
    r = np.ascontiguousarray(init_rate_rows, dtype=np.float64).copy()
    x = np.ascontiguousarray(init_next_rows, dtype=np.int32).copy()
    assert r.shape == (nr, n) and x.shape == (nr, n)
    mid = np.empty((nr, n), dtype=np.int32)
    csT = np.empty((nr, n), dtype=np.int32)
    at_i = np.empty((nr, n), dtype=np.float64)
    mid_at_i = np.empty((nr, n), dtype=np.int32)
    u = lib().fw_oracle_replay_rows(n, nr, _ip(rows), _dp(S), S.strides[0] // 8, _dp(r), _ip(x), _ip(mid),
                                    _ip(csT), _dp(at_i), _ip(mid_at_i), threads)
    if u < 0:
        raise ValueError(f"replay_rows: row {rows[-1 - u]} replaced an entry through an empty path (out of domain)")
    return RowReplay(rows, r, x, mid, csT, at_i, mid_at_i, int(u))


def max_threads() -> int:
    return int(lib().fw_oracle_max_threads())


# --------------------------------------------------------------------------
# Exact path reconstruction from (mid, csT, rs)  -- SURVEY.md 7.4
# --------------------------------------------------------------------------
def reconstruct_path(i: int, j: int, init_next: np.ndarray, mid: np.ndarray,
                     csT: np.ndarray, rs: np.ndarray, cap: int = 1 << 20) -> List[int]:
    """Index path of entry (i,j) exactly as the reference's `_path` (start excluded).

    init_next is the next matrix *before* the solve (edge(a,b) = [b] iff
    init_next[a,b] >= 0).  Iterative (explicit stack) so deep arbitrage paths
    do not hit the Python recursion limit; `cap` bounds the output length.
    """
    out: List[int] = []
    # work items: ("final", a, b) | ("col", a, k) = entry (a,k) as of step k | ("row", k, b)
    stack = [("final", i, j)]
    while stack:
        kind, a, b = stack.pop()
        if kind == "final":
            m = int(mid[a, b])
        elif kind == "col":
            m = int(csT[a, b])
        else:
            m = int(rs[a, b])
        if m < 0:
            if init_next[a, b] >= 0:
                out.append(b)
                if len(out) > cap:
                    raise OverflowError("path longer than cap")
            continue
        # path = col(a, m) ++ row(m, b): push right part first
        stack.append(("row", m, b))
        stack.append(("col", a, m))
    return out
