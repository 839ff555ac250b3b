"""Host-side mirror of the reference's Algorithms module
(/root/reference/src/lib/Algorithms.hs:2-5 exports buildMatrix, floydWarshall,
optimum) over libfwgpu's C ABI.  Same names, argument meaning and error strings.

    floyd_warshall = run_algo . build_matrix           (Algorithms.hs:19-20)

build_matrix packs the `Map (Vertex, Vertex) Double` into the dense fp64 rate
matrix + int32 next-hop matrix; run_algo is ONE call into the CUDA library
(fw_solve); the result is wrapped as a lazy `Matrix RateEntry`.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Sequence, Tuple

import numpy as np

from . import _lib, dense
from .types import AlgoOptimumError, RateEntry, RateMatrix, Vertex, isolated_entry  # noqa: F401

ExRates = Mapping[Tuple[Vertex, Vertex], float]


def sorted_vertices(ex_rates: ExRates) -> List[Vertex]:
    """Algorithms.hs:29  vertices = V.fromList . sort . nub $ M.keys exRates >>= \\(k1,k2) -> [k1,k2]."""
    seen = set()
    for (k1, k2) in ex_rates.keys():
        seen.add(k1)
        seen.add(k2)
    return sorted(seen)


def pack(ex_rates: ExRates):
    """Dense form of buildMatrix (Algorithms.hs:26-40): (vertices, rate f64[n,n], next i32[n,n]).

    Rule order as in the reference: i == j -> (0.0, []); same currency -> (1.0, [j]) BEFORE the
    map lookup; map hit -> (rate, [j]); otherwise (0.0, []).
    """
    vertices = sorted_vertices(ex_rates)
    n = len(vertices)
    index = {v: i for i, v in enumerate(vertices)}
    rate = np.zeros((n, n), dtype=np.float64)
    nxt = np.full((n, n), -1, dtype=np.int32)
    if n == 0:
        return vertices, rate, nxt
    for (s, d), r in ex_rates.items():                      # :36 M.lookup (vtxI, vtxJ)
        i, j = index[s], index[d]
        rate[i, j] = r
        nxt[i, j] = j
    ccy_ids: Dict[str, int] = {}
    cid = np.array([ccy_ids.setdefault(v.ccy, len(ccy_ids)) for v in vertices])
    same = cid[:, None] == cid[None, :]                     # :35 same currency wins over the map
    cols = np.broadcast_to(np.arange(n, dtype=np.int32)[None, :], (n, n))
    rate[same] = 1.0
    nxt[same] = cols[same]
    idx = np.arange(n)
    rate[idx, idx] = 0.0                                    # :34 i == j
    nxt[idx, idx] = -1
    return vertices, rate, nxt


def build_matrix(ex_rates: ExRates) -> RateMatrix:
    """Algorithms.hs:26-40."""
    vertices, rate, nxt = pack(ex_rates)
    return RateMatrix(vertices, rate, nxt)


def coo(ex_rates: ExRates):
    """The cache as the C ABI takes it: (vertices, ccy ids i32[n], src i32[m], dst i32[m], val f64[m])."""
    vertices = sorted_vertices(ex_rates)
    index = {v: i for i, v in enumerate(vertices)}
    ccy_ids: Dict[str, int] = {}
    ccy = np.array([ccy_ids.setdefault(v.ccy, len(ccy_ids)) for v in vertices], dtype=np.int32)
    m = len(ex_rates)
    src = np.fromiter((index[s] for (s, _d) in ex_rates), dtype=np.int32, count=m)
    dst = np.fromiter((index[d] for (_s, d) in ex_rates), dtype=np.int32, count=m)
    val = np.fromiter(ex_rates.values(), dtype=np.float64, count=m)
    return vertices, ccy, src, dst, val


def floyd_warshall(ex_rates: ExRates, ctx: _lib.Context | None = None) -> RateMatrix:
    """Algorithms.hs:19-20  floydWarshall = runAlgo 0 . buildMatrix -- both on the GPU (fw_solve_edges):
    the map goes up in COO form, the dense matrix and the exact-path tables come back."""
    import ctypes
    vertices, ccy, src, dst, val = coo(ex_rates)
    n = len(vertices)
    if n == 0:
        return RateMatrix(vertices, np.zeros((0, 0)), np.zeros((0, 0), dtype=np.int32))   # M.empty -> V.empty
    rate = np.empty((n, n), dtype=np.float64)
    out = [np.empty((n, n), dtype=np.int32) for _ in range(5)]      # next, init_next, mid, csT, rs
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    L = _lib.load()                                                  # raises FwError if not built: no fallback
    _lib.check(L.fw_solve_edges(ctx.handle if ctx else None, n, vp(ccy), len(src), vp(src), vp(dst), vp(val),
                                vp(rate), *[vp(a) for a in out]))
    nxt, init_next, mid, csT, rs = out
    return RateMatrix(vertices, rate, init_next, nxt, mid, csT, rs, ctx=ctx)


def optimum(src: Vertex, dest: Vertex, matrix: Sequence[Sequence[RateEntry]]) -> RateEntry:
    """Algorithms.hs:65-78 -- same checks in the same order, same error texts."""
    starts = []
    for row in matrix:                                       # :70 traverse ((fmap _start) . (!? 0))
        if len(row) == 0:
            raise AlgoOptimumError("The matrix is empty")
        starts.append(row[0].start if not isinstance(matrix, RateMatrix) else None)
    if isinstance(matrix, RateMatrix):
        starts = matrix.vertices

    def vertice_idx(v: Vertex) -> int:                       # :77
        try:
            return starts.index(v)
        except ValueError:
            raise AlgoOptimumError(f"{v.show()} is not entered before") from None

    src_idx = vertice_idx(src)
    dest_idx = vertice_idx(dest)
    not_reachable = f"There is no exchange between {src.show()} and {dest.show()}"
    if dest_idx >= len(matrix[src_idx]):                     # :74 matrix !? srcIdx >>= (!? destIdx)
        raise AlgoOptimumError(not_reachable)
    entry = matrix[src_idx][dest_idx]
    if not entry.path:                                       # :75
        raise AlgoOptimumError(not_reachable)
    return entry
