"""GPU parity: libfwgpu (through the C ABI) vs the CPU oracle, bit-exact.

rates: bit-identical binary64; next / mid / csT / rs: identical int32.
"""
import numpy as np
import pytest

from floydwarshall_b200 import _lib, dense, graphs
from oracle import fw_oracle as O

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_same(res, ref, paths=False):
    assert np.array_equal(bits(res.rate), bits(ref.rate)), \
        f"rate mismatch at {np.argwhere(bits(res.rate) != bits(ref.rate))[:5]}"
    assert np.array_equal(res.next, ref.next), f"next mismatch at {np.argwhere(res.next != ref.next)[:5]}"
    if paths:
        for f in ("mid", "csT", "rs"):
            a, b = getattr(res, f), getattr(ref, f)
            assert np.array_equal(a, b), f"{f} mismatch at {np.argwhere(a != b)[:5]}"


@pytest.fixture(scope="module")
def ctx():
    c = _lib.Context(0)
    yield c
    c.close()


def test_c1_readme_graph(ctx):
    """Config C1: KRAKEN BTC -> GDAX USD = 1001.0 via GDAX BTC (README.md:234-239)."""
    rate, nxt = graphs.readme_graph()
    res = dense.solve(rate, nxt, paths=True, ctx=ctx)
    ref = O.solve_dense(rate, nxt, paths=True, literal=True)
    assert_same(res, ref, paths=True)
    # vertex order GDAX-BTC, GDAX-USD, KRAKEN-BTC, KRAKEN-USD
    assert res.rate[2, 1] == 1001.0 and res.next[2, 1] == 0
    assert O.reconstruct_path(2, 1, nxt, res.mid, res.csT, res.rs) == [0, 1]
    assert res.rate[1, 0] == 0.0009
    assert O.reconstruct_path(1, 0, nxt, res.mid, res.csT, res.rs) == [3, 2, 0]


def test_empty_graph(ctx):
    r = dense.solve(np.zeros((0, 0)), np.zeros((0, 0), dtype=np.int32), ctx=ctx)
    assert r.rate.shape == (0, 0)


@pytest.mark.parametrize("mode", graphs.MODES)
@pytest.mark.parametrize("E,C", [(1, 1), (1, 2), (1, 5), (5, 7), (9, 14), (16, 8)])
def test_single_tile_sizes(ctx, E, C, mode):
    rate, nxt = graphs.exchange_graph(E, C, seed=E * 100 + C, density=0.8, mode=mode)
    ref = O.solve_dense(rate, nxt, paths=True)
    assert_same(dense.solve(rate, nxt, paths=True, ctx=ctx), ref, paths=True)
    assert_same(dense.solve(rate, nxt, paths=False, ctx=ctx), ref)


@pytest.mark.parametrize("mode", graphs.MODES)
@pytest.mark.parametrize("E,C", [(43, 3), (20, 10), (16, 16), (30, 10), (24, 16), (40, 16)])
def test_blocked_sizes(ctx, E, C, mode):
    """n = 129, 200, 256, 300, 384, 640: padded and exact multiples of the k-block."""
    rate, nxt = graphs.exchange_graph(E, C, seed=E + C, density=0.7, mode=mode)
    ref = O.solve_dense(rate, nxt, paths=True, threads=0)
    assert_same(dense.solve(rate, nxt, paths=False, ctx=ctx), ref)
    assert_same(dense.solve(rate, nxt, paths=True, ctx=ctx), ref, paths=True)


def test_c2_n1024(ctx):
    """Config C2: 64 exchanges x 16 currencies dense."""
    rate, nxt = graphs.exchange_graph(64, 16, seed=1235)
    ref = O.solve_dense(rate, nxt, paths=True, threads=0)
    assert_same(dense.solve(rate, nxt, paths=True, ctx=ctx), ref, paths=True)


def test_n2048_sparse_pairs(ctx):
    rate, nxt = graphs.exchange_graph(128, 16, seed=77, density=0.3)
    ref = O.solve_dense(rate, nxt, threads=0)
    assert_same(dense.solve(rate, nxt, ctx=ctx), ref)


@pytest.mark.parametrize("n_e,n_c,batch", [(8, 16, 48), (5, 10, 33), (1, 3, 7)])
def test_batched_fsm_replay(ctx, n_e, n_c, batch):
    """Config C3 shape (reduced batch): every snapshot solved independently."""
    rate, nxt = graphs.fsm_replay_batch(n_e, n_c, batch, seed=1236)
    res = dense.solve_batched(rate, nxt, paths=True, ctx=ctx)
    for g in range(batch):
        ref = O.solve_dense(rate[g], nxt[g], paths=True)
        got = dense.DenseResult(res.rate[g], res.next[g], res.mid[g], res.csT[g], res.rs[g])
        assert_same(got, ref, paths=True)


def test_batched_large_n_uses_blocked_path(ctx):
    rate, nxt = graphs.fsm_replay_batch(10, 16, 3, seed=5)   # n = 160 > tile
    res = dense.solve_batched(rate, nxt, ctx=ctx)
    for g in range(3):
        ref = O.solve_dense(rate[g], nxt[g])
        assert_same(dense.DenseResult(res.rate[g], res.next[g]), ref)


def test_special_values(ctx):
    """+inf, NaN, -0.0, subnormals and overflow behave as in the reference loop."""
    rng = np.random.default_rng(3)
    n = 150
    rate = rng.uniform(0.5, 1.6, size=(n, n))
    nxt = np.tile(np.arange(n, dtype=np.int32), (n, 1))
    rate[rng.random((n, n)) < 0.3] = 0.0
    nxt[rate == 0.0] = -1
    rate[3, 7] = np.inf
    rate[9, 4] = 1e308
    rate[4, 11] = 1e10
    rate[20, 21] = 5e-324
    rate[21, 22] = 0.5
    rate[30, 31] = np.nan
    rate[40, 41] = -0.0
    np.fill_diagonal(rate, 0.0)
    np.fill_diagonal(nxt, -1)
    nxt[(rate != 0) & ~np.isnan(rate) & (nxt < 0)] = 0
    with np.errstate(all="ignore"):
        ref = O.solve_dense(rate, nxt, paths=True)
    res = dense.solve(rate, nxt, paths=True, ctx=ctx)
    assert np.array_equal(bits(res.rate), bits(ref.rate))
    assert np.array_equal(res.next, ref.next) and np.array_equal(res.mid, ref.mid)


def test_diagonal_is_preserved(ctx):
    rate, nxt = graphs.exchange_graph(20, 10, seed=9)
    d = np.arange(200, dtype=np.float64) + 0.25
    np.fill_diagonal(rate, d)
    res = dense.solve(rate, nxt, ctx=ctx)
    assert np.array_equal(res.rate.diagonal(), d)
    ref = O.solve_dense(rate, nxt)
    assert_same(res, ref)


def test_domain_errors(ctx):
    rate, nxt = graphs.exchange_graph(3, 4, seed=1)
    bad = rate.copy(); bad[1, 2] = -1.0
    with pytest.raises(_lib.FwError) as ei:
        dense.solve(bad, nxt, ctx=ctx)
    assert ei.value.code == _lib.FW_ERR_DOMAIN
    badn = nxt.copy(); badn[np.argwhere(rate > 0)[0][0], np.argwhere(rate > 0)[0][1]] = -1
    with pytest.raises(_lib.FwError) as ei:
        dense.solve(rate, badn, ctx=ctx)
    assert ei.value.code == _lib.FW_ERR_DOMAIN


def test_device_resident_api(ctx):
    import torch
    rate, nxt = graphs.exchange_graph(24, 16, seed=4)     # n = 384
    ref = O.solve_dense(rate, nxt)
    rt = torch.from_numpy(rate).cuda()
    xt = torch.from_numpy(nxt).cuda()
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    dense.solve_device(ctx, rt, xt)
    torch.cuda.synchronize()
    ctx.set_stream(None)
    assert ctx.last_launches > 0
    assert_same(dense.DenseResult(rt.cpu().numpy(), xt.cpu().numpy()), ref)


@pytest.mark.parametrize("knobs", [{"FW_FUSE_PAIRS": "2"}, {"FW_FUSE_PAIRS": "0"},
                                   {"FW_FUSE_PAIRS": "2", "FW_OVERLAP": "0"},
                                   {"FW_FUSE_PAIRS": "2", "FW_BULK_BAND": "3"},
                                   {"FW_PANEL_NJ": "2"}, {"FW_PANEL_NJ": "2", "FW_FUSE_PAIRS": "2"},
                                   {"FW_FUSE_GROUP": "4"}, {"FW_FUSE_GROUP": "4", "FW_OVERLAP": "0"},
                                   {"FW_FUSE_GROUP": "8"}, {"FW_FUSE_GROUP": "1"}])
@pytest.mark.parametrize("E,C", [(24, 16), (45, 16), (70, 10), (64, 16)])
def test_schedule_variants(monkeypatch, E, C, knobs):
    """n = 384, 720, 700, 1024: every schedule the large solves use (k-blocks in groups of 2 or 4, which by default
    starts at 48 k-blocks; side-stream look-ahead; raster bands; two panel jobs per half-warp, which by
    default starts at 9472 rows) must give the reference's bits.
    The knobs are read when a context first launches, so each case takes a fresh context."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    c = _lib.Context(0)
    try:
        rate, nxt = graphs.exchange_graph(E, C, seed=3 * E + C, density=0.6)
        ref = O.solve_dense(rate, nxt, paths=True, threads=0)
        assert_same(dense.solve(rate, nxt, paths=True, ctx=c), ref, paths=True)
        assert_same(dense.solve(rate, nxt, paths=False, ctx=c), ref)
    finally:
        c.close()
