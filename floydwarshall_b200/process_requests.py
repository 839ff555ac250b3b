"""Mirror of the reference's ProcessRequests module
(/root/reference/src/lib/ProcessRequests.hs): the InSync/OutSync state machine that is the only
production caller of the hot path.  `find_best_rate` triggers `floyd_warshall` (the CUDA solve)
exactly where the reference does (ProcessRequests.hs:82-84) and nowhere else.

Two engines for the InSync matrix:
  * default: `algorithms.floyd_warshall` -> a `RateMatrix` with host-side tables;
  * `ResidentEngine`: libfwgpu's fw_state_* -- buildMatrix + runAlgo on the device, the optimised
    matrix stays in HBM and `optimum` reads one (rate, path) answer back per query.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from datetime import datetime
from typing import Dict, List, Optional, Tuple, Union

import numpy as np

from . import _lib, algorithms
from .parsers import ParseInputError, parse_exch_pair, parse_rates, show_double, show_utc
from .types import AlgoOptimumError, RateEntry, Vertex

ExchRateTimes = Dict[Tuple[Vertex, Vertex], Tuple[float, datetime]]      # Types.hs:31


@dataclass
class OutSync:                      # Types.hs:37
    ex_rates: ExchRateTimes


@dataclass
class InSync:                       # Types.hs:36
    ex_rates: ExchRateTimes
    matrix: object                  # Matrix RateEntry (RateMatrix or ResidentMatrix)


AppState = Union[InSync, OutSync]


def blank_state() -> AppState:      # Utils.hs:16-17
    return OutSync({})


@dataclass
class DisplayMessage:               # Types.hs:43-57 (a monoid)
    err: List[str] = field(default_factory=list)
    res: List[str] = field(default_factory=list)


class ResidentMatrix:
    """`Matrix RateEntry` that lives in HBM (fw_state_*); answers `optimum` lookups only."""

    def __init__(self, ex_rates: Dict[Tuple[Vertex, Vertex], float], ctx: Optional[_lib.Context] = None):
        L = _lib.load()
        self.vertices, ccy, src, dst, val = algorithms.coo(ex_rates)
        self.index = {v: i for i, v in enumerate(self.vertices)}
        n, m = len(self.vertices), len(src)
        h = ctypes.c_void_p()
        _lib.check(L.fw_state_create(ctx.handle if ctx else None, ctypes.byref(h)))
        self._h, self.n = h, n
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        _lib.check(L.fw_state_sync(h, n, vp(ccy), m, vp(src), vp(dst), vp(val)))

    def __len__(self):
        return self.n

    def lookup(self, i: int, j: int, cap: int = 4096) -> RateEntry:
        L = _lib.load()
        rate = ctypes.c_double()
        plen = ctypes.c_int32()
        while True:
            path = np.empty(cap, dtype=np.int32)
            rc = L.fw_state_optimum(self._h, i, j, ctypes.byref(rate), ctypes.c_void_p(path.ctypes.data), cap,
                                    ctypes.byref(plen))
            if rc == _lib.FW_ERR_CAP and plen.value > cap:
                cap = plen.value
                continue
            _lib.check(rc)
            break
        return RateEntry(rate.value, self.vertices[i], [self.vertices[k] for k in path[:plen.value]])

    def download(self):
        L = _lib.load()
        r = np.empty((self.n, self.n), dtype=np.float64)
        x = np.empty((self.n, self.n), dtype=np.int32)
        _lib.check(L.fw_state_download(self._h, ctypes.c_void_p(r.ctypes.data), ctypes.c_void_p(x.ctypes.data)))
        return r, x

    def close(self):
        if self._h:
            _lib.load().fw_state_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def optimum_resident(src: Vertex, dest: Vertex, m: ResidentMatrix) -> RateEntry:
    """Algorithms.hs:65-78 against the device-resident matrix (same checks, same texts)."""
    if src not in m.index:
        raise AlgoOptimumError(f"{src.show()} is not entered before")
    if dest not in m.index:
        raise AlgoOptimumError(f"{dest.show()} is not entered before")
    e = m.lookup(m.index[src], m.index[dest])
    if not e.path:
        raise AlgoOptimumError(f"There is no exchange between {src.show()} and {dest.show()}")
    return e


def _strip_times(ex: ExchRateTimes) -> Dict[Tuple[Vertex, Vertex], float]:
    return {k: v[0] for k, v in ex.items()}                # ProcessRequests.hs:83  M.map fst


def update_rates(line: str, state: AppState) -> AppState:
    """ProcessRequests.hs:89-102.  Raises ParseInputError; returns the (possibly unchanged) state."""
    time, src, dest, fwd, bkd = parse_rates(line)
    ex = state.ex_rates
    old = ex.get((src, dest))
    if old is None or old[1] < time:                        # :97-98 newer timestamp only
        new = dict(ex)
        new[(dest, src)] = (bkd, time)                      # :101-102 both directions
        new[(src, dest)] = (fwd, time)
        return OutSync(new)                                 # any accepted update drops the matrix
    return state


def find_best_rate(line: str, state: AppState, resident: bool = False, ctx=None) -> Tuple[RateEntry, AppState]:
    """ProcessRequests.hs:70-85.  Raises ParseInputError / AlgoOptimumError."""
    src, dest = parse_exch_pair(line)
    if isinstance(state, OutSync):                          # :82-84 syncMatrix: the ONE call into the hot path
        rates = _strip_times(state.ex_rates)
        matrix = ResidentMatrix(rates, ctx) if (resident and rates) else algorithms.floyd_warshall(rates, ctx)
        state = InSync(state.ex_rates, matrix)              # :79 put (InSync exRates matrix)
    matrix = state.matrix
    entry = optimum_resident(src, dest, matrix) if isinstance(matrix, ResidentMatrix) \
        else algorithms.optimum(src, dest, matrix)          # :80
    return entry, state


def present_rate_entry(e: RateEntry) -> List[str]:
    """ProcessRequests.hs:53-63."""
    d = e.path[-1]
    header = f"BEST_RATES_BEGIN {e.start.exch} {e.start.ccy} {d.exch} {d.ccy} {show_double(e.best_rate)}"
    return [header] + [v.show() for v in [e.start] + e.path] + ["BEST_RATES_END"]


def serve_req(line: str, state: AppState, resident: bool = False, ctx=None) -> Tuple[AppState, DisplayMessage]:
    """ProcessRequests.hs:31-52: try updateRates; on a parse error try findBestRate; collect messages."""
    msg = DisplayMessage()
    try:
        state = update_rates(line, state)
        for (s, d), (rate, time) in sorted(state.ex_rates.items()):          # :47-50 M.toAscList
            msg.res.append(f"{s.show()} -- {show_double(rate)} {show_utc(time)} --> {d.show()}")
        return state, msg
    except ParseInputError as e1:
        msg.err += [e1.msg, "Invalid request to update rates, probably a request for best rate"]
    try:
        entry, state = find_best_rate(line, state, resident, ctx)
        msg.res += present_rate_entry(entry)
    except (ParseInputError, AlgoOptimumError) as e2:
        msg.err.append(e2.msg)
    return state, msg


def user_prompt_lines(line: str, state: AppState, resident: bool = False, ctx=None) -> Tuple[AppState, List[str]]:
    """What one iteration of Main.userPrompt prints (src/app/Main.hs:18-37)."""
    new_state, m = serve_req(line, state, resident, ctx)
    if not m.res:
        return new_state, m.err + ["You neither enter exchange rates or request best rate, please enter a valid input\n"]
    return new_state, m.res + [""]
