"""Synthetic exchange/currency rate graphs.

Vertex id = e*C + c, which is the reference's index order (`sort . nub` over
`Vertex`, exchange-major -- /root/reference/src/lib/Algorithms.hs:29,
Types.hs:13-17) when exchange and currency names are fixed-width letters.

Dense encoding: rate f64[N,N] and next i32[N,N] exactly as `buildMatrix`
(/root/reference/src/lib/Algorithms.hs:26-40) would produce them:
  i == j              -> (0.0, -1)
  same currency       -> (1.0, j)      cross-exchange transfer, exactly 1.0
  same exchange, pair -> (rate, j)
  otherwise           -> (0.0, -1)
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

MODES = ("consistent", "arbitrage", "ones", "pow2")


def _alpha(i: int, width: int) -> str:
    s = []
    for _ in range(width):
        s.append(chr(ord("A") + i % 26))
        i //= 26
    return "".join(reversed(s))


def vertex_names(E: int, C: int):
    """Fixed-width alphabetic names so that string order == (e, c) order."""
    we = max(2, int(np.ceil(np.log(max(E, 2)) / np.log(26))) + 1)
    wc = max(2, int(np.ceil(np.log(max(C, 2)) / np.log(26))) + 1)
    return [("X" + _alpha(e, we), "C" + _alpha(c, wc)) for e in range(E) for c in range(C)]


def exchange_blocks(E: int, C: int, seed: int, density: float = 1.0, mode: str = "consistent"):
    """Per-exchange C x C rate blocks, f64[E,C,C] (0.0 = pair not quoted).

    consistent: shared mid-prices, spread s~U(5e-4,1e-2) per (exchange, pair):
                rate(a->b) = p_a/p_b*(1-s), rate(b->a) = p_b/p_a*(1-s); arbitrage-free.
    arbitrage:  per-exchange price noise larger than the spread (cross-exchange cycles > 1).
    ones:       every quoted pair is exactly 1.0 (all ties).
    pow2:       rates are exact powers of two (products exact; ties structural).
    """
    assert mode in MODES
    rng = np.random.Generator(np.random.PCG64(seed))
    blocks = np.zeros((E, C, C), dtype=np.float64)
    iu, ju = np.triu_indices(C, k=1)
    npair = iu.size
    present = rng.random((E, npair)) < density
    if mode == "ones":
        fwd = np.ones((E, npair))
        bkd = np.ones((E, npair))
    elif mode == "pow2":
        ex = rng.integers(-6, 7, size=C)
        d = (ex[iu] - ex[ju])[None, :] - rng.integers(0, 2, size=(E, npair))
        d2 = (ex[ju] - ex[iu])[None, :] - rng.integers(0, 2, size=(E, npair))
        fwd = np.ldexp(1.0, d)
        bkd = np.ldexp(1.0, np.minimum(d2, -d))  # fwd*bkd <= 1
    else:
        p = rng.uniform(0.01, 5.0e4, size=C)
        s = rng.uniform(5.0e-4, 1.0e-2, size=(E, npair))
        if mode == "arbitrage":
            noise = np.exp(rng.normal(0.0, 0.03, size=(E, C)))
            pe = p[None, :] * noise
        else:
            pe = np.broadcast_to(p[None, :], (E, C))
        fwd = pe[:, iu] / pe[:, ju] * (1.0 - s)
        bkd = pe[:, ju] / pe[:, iu] * (1.0 - s)
    blocks[:, iu, ju] = np.where(present, fwd, 0.0)
    blocks[:, ju, iu] = np.where(present, bkd, 0.0)
    return blocks


def dense_from_blocks(blocks: np.ndarray, out_rate=None, out_next=None):
    """(rate f64[N,N], next i32[N,N]) of `buildMatrix` for E exchanges x C currencies."""
    E, C, _ = blocks.shape
    N = E * C
    rate = out_rate if out_rate is not None else np.zeros((N, N), dtype=np.float64)
    nxt = out_next if out_next is not None else np.empty((N, N), dtype=np.int32)
    if out_rate is not None:
        rate[...] = 0.0
    nxt[...] = -1
    r4 = rate.reshape(E, C, E, C)
    n4 = nxt.reshape(E, C, E, C)
    cols = np.arange(N, dtype=np.int32).reshape(E, C)
    for c in range(C):                       # same currency, any two exchanges: exactly 1.0
        r4[:, c, :, c] = 1.0
        n4[:, c, :, c] = cols[None, :, c]
    for e in range(E):                       # same exchange: quoted pairs
        blk = blocks[e]
        r4[e, :, e, :] = blk
        n4[e, :, e, :] = np.where(blk != 0.0, cols[e][None, :], -1)
    idx = np.arange(N)
    rate[idx, idx] = 0.0
    nxt[idx, idx] = -1
    return rate, nxt


def exchange_graph(E: int, C: int, seed: int, density: float = 1.0, mode: str = "consistent"):
    return dense_from_blocks(exchange_blocks(E, C, seed, density, mode))


def rates_map_from_blocks(blocks: np.ndarray) -> Dict[Tuple[Tuple[str, str], Tuple[str, str]], float]:
    """The same graph as the reference's input type: Map (Vertex, Vertex) Double.

    Only same-exchange quoted pairs are map entries (ProcessRequests.hs:101-102);
    the 1.0 cross-exchange edges are synthesised by buildMatrix itself.
    """
    E, C, _ = blocks.shape
    names = vertex_names(E, C)
    m = {}
    for e in range(E):
        for a in range(C):
            for b in range(C):
                r = blocks[e, a, b]
                if a != b and r != 0.0:
                    m[(names[e * C + a], names[e * C + b])] = float(r)
    return m


def fsm_replay_batch(E: int, C: int, T: int, seed: int):
    """T successive snapshots of an E x C graph under single-pair rate updates.

    Mirrors the FSM: every accepted `updateRates` request flips the state to
    OutSync (/root/reference/src/lib/ProcessRequests.hs:97-102), so the next
    best-rate query recomputes from scratch on the mutated cache
    (ProcessRequests.hs:82-84).  Snapshot t is the dense buildMatrix of the
    cache after update t.  Returns rate f64[T,N,N], next i32[T,N,N].
    """
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    blocks = exchange_blocks(E, C, seed, 1.0, "consistent")
    rng0 = np.random.Generator(np.random.PCG64(seed))
    p = rng0.uniform(0.01, 5.0e4, size=C)     # same mid-prices exchange_blocks drew first
    N = E * C
    rate = np.empty((T, N, N), dtype=np.float64)
    nxt = np.empty((T, N, N), dtype=np.int32)
    base_r, base_n = dense_from_blocks(blocks)
    for t in range(T):
        e = int(rng.integers(0, E))
        a = int(rng.integers(0, C))
        b = int(rng.integers(0, C - 1))
        if b >= a:
            b += 1
        s = rng.uniform(5.0e-4, 1.0e-2)
        fwd = p[a] / p[b] * (1.0 - s)
        bkd = p[b] / p[a] * (1.0 - s)
        base_r[e * C + a, e * C + b] = fwd      # both directions (ProcessRequests.hs:101-102)
        base_r[e * C + b, e * C + a] = bkd
        rate[t] = base_r
        nxt[t] = base_n
    return rate, nxt


def readme_graph():
    """Config C1: the four MockData rows (/root/reference/src/test/MockData.hs:47-57).

    Vertex order GDAX-BTC, GDAX-USD, KRAKEN-BTC, KRAKEN-USD."""
    blocks = np.zeros((2, 2, 2))
    blocks[0, 0, 1] = 1001.0   # GDAX BTC->USD
    blocks[0, 1, 0] = 0.0008   # GDAX USD->BTC
    blocks[1, 0, 1] = 1000.0   # KRAKEN BTC->USD
    blocks[1, 1, 0] = 0.0009   # KRAKEN USD->BTC
    return dense_from_blocks(blocks)
