"""Bit-exact parity of the SHIPPED large-N schedules against the CPU oracle.

The small-N tests (test_gpu_parity.py) cannot reach the code paths the headline numbers run on:
k-blocks in groups of 8 WITH a next group (from 128 k-blocks), groups of 4 (from 48), two panel jobs
per half-warp (from 9600 rows).  Here:

  * N=4096 with groups of 8 forced (4 groups), N=8192 under the default policy (groups of 4, 16 groups),
    N=9728 (default policy, two panel jobs per half-warp): the full oracle loop, rates / next / mid / csT / rs;
  * config C3 at its full batch of 4096 graphs;
  * N=32768 (config C4, default policy = groups of 8):
      - row replay: the solve records every pivot row as its step begins (fw_ctx_set_row_snapshot_sink);
        the oracle replays the whole history of sampled rows from them (fw_oracle_replay_rows) and the
        rows' final rates / next / mid, their csT rows, the recorded pivot rows and their rs rows must match
        bit for bit;
      - an independent schedule (one k-block per launch) must give the identical full result;
      - 128-step windows: GPU state at k0 -> 128 oracle steps on the host -> GPU state at k0+128.

Semantics under test: /root/reference/src/lib/Algorithms.hs:42-61 (strict `<` :55, one rounded multiply :61,
ascending k, ties keep the older entry).
"""
import ctypes
import os

import numpy as np
import pytest

from floydwarshall_b200 import _lib, dense, graphs
from oracle import fw_oracle as O

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_same(res, ref):
    assert np.array_equal(bits(res.rate), bits(ref.rate)), \
        f"rate mismatch at {np.argwhere(bits(res.rate) != bits(ref.rate))[:5]}"
    assert np.array_equal(res.next, ref.next), f"next mismatch at {np.argwhere(res.next != ref.next)[:5]}"
    for f in ("mid", "csT", "rs"):
        a, b = getattr(res, f), getattr(ref, f)
        if a is not None and b is not None:
            assert np.array_equal(a, b), f"{f} mismatch at {np.argwhere(a != b)[:5]}"


def fresh_ctx(monkeypatch, **knobs):
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    return _lib.Context(0)      # knobs are read when a context first launches


def test_n4096_groups_of_8(monkeypatch):
    """4 groups of 8 k-blocks: the next-group strips, panel sets 8..15 and the set rotation all run."""
    c = fresh_ctx(monkeypatch, FW_FUSE_GROUP="8")
    try:
        rate, nxt = graphs.exchange_graph(256, 16, seed=1301)
        ref = O.solve_dense(rate, nxt, paths=True, threads=0)
        assert_same(dense.solve(rate, nxt, paths=True, ctx=c), ref)
        res = dense.solve(rate, nxt, paths=False, ctx=c)
        assert np.array_equal(bits(res.rate), bits(ref.rate)) and np.array_equal(res.next, ref.next)
    finally:
        c.close()


def test_n4096_groups_of_8_arbitrage(monkeypatch):
    c = fresh_ctx(monkeypatch, FW_FUSE_GROUP="8")
    try:
        rate, nxt = graphs.exchange_graph(256, 16, seed=1302, density=0.5, mode="arbitrage")
        ref = O.solve_dense(rate, nxt, paths=True, threads=0)
        assert_same(dense.solve(rate, nxt, paths=True, ctx=c), ref)
    finally:
        c.close()


def test_n8192_default_policy():
    """North-star size N=8192 on one GPU: the default policy (groups of 4, 16 groups)."""
    c = _lib.Context(0)
    try:
        rate, nxt = graphs.exchange_graph(512, 16, seed=1303)
        ref = O.solve_dense(rate, nxt, paths=True, threads=0)
        assert_same(dense.solve(rate, nxt, paths=True, ctx=c), ref)
    finally:
        c.close()


def test_n9728_two_panel_jobs_by_default():
    """76 k-blocks: groups of 4 and, from 9600 panel rows, two jobs per half-warp in the panel kernels."""
    c = _lib.Context(0)
    try:
        rate, nxt = graphs.exchange_graph(608, 16, seed=1304)
        ref = O.solve_dense(rate, nxt, threads=0)
        res = dense.solve(rate, nxt, ctx=c)
        assert np.array_equal(bits(res.rate), bits(ref.rate)) and np.array_equal(res.next, ref.next)
    finally:
        c.close()


def test_c3_full_batch():
    """Config C3 as BASELINE.json states it: 4096 snapshot graphs of the N=128 FSM replay."""
    c = _lib.Context(0)
    try:
        rate, nxt = graphs.fsm_replay_batch(8, 16, 4096, seed=1236)
        res = dense.solve_batched(rate, nxt, paths=True, ctx=c)
        ref = O.solve_batched(rate, nxt, paths=True, threads=0)
        assert_same(res, ref)
    finally:
        c.close()


# ---------------------------------------------------------------------------- N = 32768
N = 32768
SEED = 1237


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _need_hbm(gib):
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < gib * 2 ** 30:
        pytest.skip(f"needs ~{gib} GiB of free HBM")


def test_n32768_row_replay_and_independent_schedule(monkeypatch):
    import torch
    from bench import device_graph
    _need_hbm(90)
    dev = torch.device("cuda", 0)
    L = _lib.load()
    r0, x0 = device_graph(N, SEED, dev)
    ctx = _lib.Context(0)                                   # default policy: groups of 8
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    r, x = r0.clone(), x0.clone()
    mid, csT, rs = (torch.empty_like(x0) for _ in range(3))
    sink = torch.empty((N, N), dtype=torch.float64, device=dev)
    _lib.check(L.fw_ctx_set_row_snapshot_sink(ctx.handle, _p(sink), N))
    dense.solve_device(ctx, r, x, mid, csT, rs)
    torch.cuda.synchronize()
    _lib.check(L.fw_ctx_set_row_snapshot_sink(ctx.handle, None, 0))

    # ---- sampled rows: first / last rows, group and block borders, random ones
    rng = np.random.default_rng(7)
    rows = sorted(set([0, 1, 127, 128, 1023, 1024, 1025, 16383, 16384, N - 1025, N - 129, N - 128, N - 1]
                      + rng.integers(0, N, size=51).tolist()))
    rows = np.array(rows, dtype=np.int32)
    S = sink.cpu().numpy()
    del sink
    idx = torch.from_numpy(rows.astype(np.int64)).to(dev)
    rp = O.replay_rows(rows, S, r0[idx].cpu().numpy(), x0[idx].cpu().numpy(), threads=0)
    assert rp.updates > 0
    got_r = r[idx].cpu().numpy()
    assert np.array_equal(bits(got_r), bits(rp.rate)), np.argwhere(bits(got_r) != bits(rp.rate))[:5]
    assert np.array_equal(x[idx].cpu().numpy(), rp.next)
    assert np.array_equal(mid[idx].cpu().numpy(), rp.mid)
    assert np.array_equal(csT[idx].cpu().numpy(), rp.csT)
    assert np.array_equal(rs[idx].cpu().numpy(), rp.mid_at_i)
    # the recorded pivot rows are what the loop itself has in those rows as their steps begin
    at = rp.at_i.copy()
    at[np.arange(len(rows)), rows] = 0.0                     # the pivot entry is recorded as 0.0
    assert np.array_equal(bits(S[rows]), bits(at))
    del S

    # ---- an independent schedule: one k-block per bulk launch, no look-ahead stream
    c1 = fresh_ctx(monkeypatch, FW_FUSE_GROUP="1", FW_OVERLAP="0")
    c1.set_stream(torch.cuda.current_stream().cuda_stream)
    r1, x1 = r0.clone(), x0.clone()
    mid1, csT1, rs1 = (torch.empty_like(x0) for _ in range(3))
    dense.solve_device(c1, r1, x1, mid1, csT1, rs1)
    torch.cuda.synchronize()
    assert torch.equal(r1.view(torch.int64), r.view(torch.int64))
    assert torch.equal(x1, x) and torch.equal(mid1, mid) and torch.equal(csT1, csT) and torch.equal(rs1, rs)
    c1.close()
    ctx.close()


@pytest.mark.parametrize("kb0", [0, 127])
def test_n32768_window_against_128_oracle_steps(kb0):
    """GPU state as step kb0*128 begins (default schedule up to there) -> 128 steps of the oracle loop on the
    host -> must equal the GPU state 128 steps later, over the full matrix."""
    import torch
    from bench import device_graph
    _need_hbm(40)
    dev = torch.device("cuda", 0)
    L = _lib.load()
    r, x = device_graph(N, SEED, dev)
    ctx = _lib.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    _lib.check(L.fw_solve_device_range(ctx.handle, N, N, _p(r), _p(x), None, None, None, 0, kb0), ctx.handle)
    torch.cuda.synchronize()
    rh, xh = r.cpu().numpy(), x.cpu().numpy()
    O.run_ksteps(rh, xh, kb0 * 128, kb0 * 128 + 128, 0)
    _lib.check(L.fw_solve_device_range(ctx.handle, N, N, _p(r), _p(x), None, None, None, kb0, kb0 + 1), ctx.handle)
    torch.cuda.synchronize()
    # compare on the device (12 GB up is cheaper than 12 GB down + a host compare)
    rt = torch.from_numpy(rh).to(dev)
    assert torch.equal(rt.view(torch.int64), r.view(torch.int64))
    del rt
    assert torch.equal(torch.from_numpy(xh).to(dev), x)
    ctx.close()
