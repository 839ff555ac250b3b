"""Host-side mirror of the reference's core types (/root/reference/src/lib/Types.hs).

The reference is Haskell; there is no GHC in this image, so the host side above
the C ABI is written in Python with the same names, field meaning and error
behaviour.  `RateMatrix` is the `Matrix RateEntry` (Types.hs:39): a sequence of
rows of `RateEntry`, backed by the dense buffers the CUDA library fills; entries
and their `_path` lists are materialised lazily on access.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True, order=True)
class Vertex:
    """Types.hs:13-17  data Vertex = Vertex {_exch, _ccy :: Text} deriving (Ord, Eq)."""
    exch: str
    ccy: str

    def show(self) -> str:
        """Types.hs:19-20  show = "(" <> _exch <> ", " <> _ccy <> ")"."""
        return f"({self.exch}, {self.ccy})"

    def __str__(self) -> str:
        return self.show()


@dataclass
class RateEntry:
    """Types.hs:24-29  RateEntry {_bestRate :: Double, _start :: Vertex, _path :: [Vertex]}."""
    best_rate: float
    start: Vertex
    path: List[Vertex] = field(default_factory=list)


def isolated_entry(start: Vertex) -> RateEntry:
    """Utils.hs:13-14  isolatedEntry start = RateEntry 0.0 start []."""
    return RateEntry(0.0, start, [])


class AlgoOptimumError(Exception):
    """Types.hs:62-63  newtype AlgoError = AlgoOptimumError Text."""

    def __init__(self, msg: str):
        super().__init__(msg)
        self.msg = msg


class _Row(Sequence):
    def __init__(self, m: "RateMatrix", i: int):
        self._m, self._i = m, i

    def __len__(self) -> int:
        return self._m.n

    def __getitem__(self, j):
        if isinstance(j, slice):
            return [self[k] for k in range(*j.indices(len(self)))]
        if j < 0:
            j += len(self)
        if not 0 <= j < len(self):
            raise IndexError(j)
        return self._m.entry(self._i, j)

    def materialize(self) -> List["RateEntry"]:
        n = len(self)
        ps = self._m.index_paths([(self._i, j) for j in range(n)])        # one device call for the whole row
        return [RateEntry(float(self._m.rate[self._i, j]), self._m.vertices[self._i],
                          [self._m.vertices[k] for k in ps[j]]) for j in range(n)]

    def __eq__(self, other):
        if len(self) != len(other):
            return False
        theirs = other.materialize() if isinstance(other, _Row) else other
        return all(a == b for a, b in zip(self.materialize(), theirs))


class RateMatrix(Sequence):
    """`Matrix RateEntry` over dense buffers.

    rate[i,j] = _bestRate, vertices[i] = _start of row i; `_path` of entry (i,j):
      * for a buildMatrix result (no side tables): [vertices[j]] iff init_next[i,j] >= 0
      * for a floydWarshall result: expanded by libfwgpu's fw_paths from (mid, csT, rs).
    """

    def __init__(self, vertices: List[Vertex], rate: np.ndarray, init_next: np.ndarray,
                 nxt: Optional[np.ndarray] = None, mid: Optional[np.ndarray] = None,
                 csT: Optional[np.ndarray] = None, rs: Optional[np.ndarray] = None, ctx=None):
        self.vertices = vertices
        self.n = len(vertices)
        self.rate = rate
        self.init_next = init_next
        self.next = nxt if nxt is not None else init_next
        self.mid, self.csT, self.rs = mid, csT, rs
        self._ctx = ctx
        self._tables = None       # device copy of (init_next, mid, csT, rs): uploaded on first use, once

    # -- Sequence protocol: rows of RateEntry
    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(self.n))]
        if i < 0:
            i += self.n
        if not 0 <= i < self.n:
            raise IndexError(i)
        return _Row(self, i)

    def __iter__(self) -> Iterator[_Row]:
        return (_Row(self, i) for i in range(self.n))

    def __eq__(self, other):
        """Types.hs:24-29 derives Eq entry by entry; here all n*n paths come from ONE device call per side."""
        try:
            if len(self) != len(other):
                return False
            mine = self.to_lists()
            theirs = other.to_lists() if isinstance(other, RateMatrix) else other
            return all(len(a) == len(b) and all(x == y for x, y in zip(a, b)) for a, b in zip(mine, theirs))
        except TypeError:
            return NotImplemented

    # -- entries
    def index_paths(self, pairs: Sequence[Tuple[int, int]]) -> List[List[int]]:
        """Exact reference paths (index lists) for many (i,j) pairs in one device call."""
        if self.mid is None:
            return [[j] if self.init_next[i, j] >= 0 else [] for i, j in pairs]
        if self._tables is None:
            from . import paths
            self._tables = paths.DeviceTables(self.init_next, self.mid, self.csT, self.rs, ctx=self._ctx)
        return self._tables.expand(pairs)

    def entry(self, i: int, j: int) -> RateEntry:
        p = self.index_paths([(i, j)])[0]
        return RateEntry(float(self.rate[i, j]), self.vertices[i], [self.vertices[k] for k in p])

    def to_lists(self) -> List[List[RateEntry]]:
        """Fully materialised nested lists (small graphs / tests)."""
        pairs = [(i, j) for i in range(self.n) for j in range(self.n)]
        ps = self.index_paths(pairs)
        out, t = [], 0
        for i in range(self.n):
            row = []
            for j in range(self.n):
                row.append(RateEntry(float(self.rate[i, j]), self.vertices[i],
                                     [self.vertices[k] for k in ps[t]]))
                t += 1
            out.append(row)
        return out
