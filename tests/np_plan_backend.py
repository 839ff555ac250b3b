"""numpy model of the multi-GPU plan operations (TEST INFRASTRUCTURE).

libfwgpu's fw_multi_plan() returns the schedule of the row-sharded solve as a list of semantic operations
(include/fwgpu.h: FW_OP_*).  The CUDA executor maps them to kernels; this file gives every operation its
literal meaning in terms of the reference loop (Algorithms.hs:42-61) so that the SAME schedule can be run on
CPU -- as virtual ranks in one process, or with one process per rank under gloo -- and compared with the
oracle."""
import ctypes

import numpy as np

from floydwarshall_b200 import _lib


def get_plan(n, world, B, G, cbr):
    L = _lib.load()
    cnt = int(L.fw_multi_plan(n, world, B, G, cbr, None, 0))
    assert cnt > 0, "bad layout"
    ops = (_lib.PlanOp * cnt)()
    assert int(L.fw_multi_plan(n, world, B, G, cbr, ops, cnt)) == cnt
    return list(ops)


class Layout:
    def __init__(self, n, world, B, G, cbr):
        self.n, self.world, self.B, self.G, self.cbr = n, world, B, G, cbr
        self.rows = n // world

    def glob(self, rank, l):
        return ((l // self.cbr) * self.world + rank) * self.cbr + l % self.cbr

    def local_rows_of(self, rank):
        return np.array([self.glob(rank, l) for l in range(self.rows)])


class NpRank:
    """One rank's shard (local rows of rate / next) and its Rw panel buffers."""

    def __init__(self, lay: Layout, rank: int, rate_full: np.ndarray, next_full: np.ndarray):
        self.lay, self.rank = lay, rank
        self.gidx = lay.local_rows_of(rank)
        self.rate = rate_full[self.gidx].copy()
        self.next = next_full[self.gidx].copy()
        self.Rw = [np.zeros((lay.B, lay.n)) for _ in range(2 * lay.G)]

    def _relax(self, rows, k, brow):
        """rows (local indices) take step k against the pivot row `brow` (row k as step k began)."""
        if len(rows) == 0:
            return
        R, X = self.rate[rows], self.next[rows]
        a, an = R[:, k].copy(), X[:, k].copy()
        with np.errstate(invalid="ignore", over="ignore"):
            nv = np.outer(a, brow)                     # one rounded multiply (Algorithms.hs:61)
            upd = R < nv                               # strict (Algorithms.hs:55)
        upd[:, k] = False                              # j == k  (Algorithms.hs:54)
        upd[np.arange(len(rows)), self.gidx[rows]] = False   # j == i
        R[upd] = nv[upd]
        X[upd] = np.broadcast_to(an[:, None], R.shape)[upd]
        self.rate[rows], self.next[rows] = R, X

    def pivot(self, op):
        B = self.lay.B
        blk = np.arange(op.row_lo, op.row_lo + B)
        for kk in range(B):
            k = op.b0 + kk
            assert self.gidx[op.row_lo + kk] == k
            self.Rw[op.buf][kk] = self.rate[op.row_lo + kk]
            self._relax(blk[blk != op.row_lo + kk], k, self.Rw[op.buf][kk])      # i == k is skipped (Algorithms.hs:50)

    def apply(self, op):
        B = self.lay.B
        rows = np.arange(op.row_lo, op.row_lo + op.row_n)
        if op.ex_n > 0:
            rows = rows[(rows < op.ex_lo) | (rows >= op.ex_lo + op.ex_n)]
        for blk in range(op.nb):
            sel = rows
            if op.grp_lo >= 0:       # a row in the blocks' own rows, block i, takes only blocks i+1 ..
                own = (rows >= op.grp_lo) & (rows < op.grp_lo + op.nb * B)
                sel = rows[~own | ((rows - op.grp_lo) // B < blk)]
            for kk in range(B):
                self._relax(sel, op.b0 + blk * B + kk, self.Rw[op.buf + blk][kk])


def run_virtual(lay: Layout, rate, nxt):
    """All ranks in one process, operations in the plan's issue order; returns the gathered (rate, next)."""
    ranks = [NpRank(lay, r, rate, nxt) for r in range(lay.world)]
    for op in get_plan(lay.n, lay.world, lay.B, lay.G, lay.cbr):
        if op.kind == _lib.OP_PIVOT:
            ranks[op.rank].pivot(op)
        elif op.kind == _lib.OP_APPLY:
            ranks[op.rank].apply(op)
        elif op.kind == _lib.OP_BCAST:
            for t in ranks:
                if t.rank != op.rank:
                    t.Rw[op.buf][:] = ranks[op.rank].Rw[op.buf]
    out_r, out_x = np.empty_like(rate), np.empty_like(nxt)
    for t in ranks:
        out_r[t.gidx], out_x[t.gidx] = t.rate, t.next
    return out_r, out_x
