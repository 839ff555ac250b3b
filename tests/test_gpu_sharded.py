"""The multi-GPU solve behind the C ABI (fw_multi_*) against the CPU oracle, bit for bit.

On a one-GPU box the ranks are VIRTUAL: fw_multi_create is given the same device several times, so the real
C++ schedule (csrc/fw_plan.hpp), the real executor (two stream lanes per shard, events, panel copies) and the
real kernels run -- only the transport is a device-to-device copy instead of NVLink.  With two or more GPUs
the same tests also run across devices with both transports, and once through torchrun (one process per GPU,
NCCL) via bench.py --check."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from floydwarshall_b200 import _lib, algorithms, graphs, sharded
from oracle import fw_oracle as O

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_same(res, ref, paths):
    assert np.array_equal(bits(res.rate), bits(ref.rate)), np.argwhere(bits(res.rate) != bits(ref.rate))[:5]
    assert np.array_equal(res.next, ref.next), np.argwhere(res.next != ref.next)[:5]
    if paths:
        for f in ("mid", "csT", "rs"):
            a, b = getattr(res, f), getattr(ref, f)
            assert np.array_equal(a, b), f"{f} mismatch at {np.argwhere(a != b)[:5]}"


@pytest.mark.parametrize("cyclic", ["1", "0"])
@pytest.mark.parametrize("world,E,C,mode,G", [
    (1, 24, 16, "consistent", 1), (2, 32, 16, "consistent", 1), (4, 32, 16, "arbitrage", 1),
    (1, 32, 16, "consistent", 2), (2, 32, 16, "arbitrage", 2), (2, 64, 16, "consistent", 2), (4, 64, 16, "pow2", 2),
    (1, 64, 16, "arbitrage", 4), (2, 64, 16, "consistent", 4), (2, 128, 16, "consistent", 4),
    (1, 128, 16, "consistent", 8), (2, 128, 16, "arbitrage", 8), (3, 96, 16, "consistent", 2), (8, 128, 16, "consistent", 2)])
def test_virtual_ranks_match_oracle(monkeypatch, world, E, C, mode, G, cyclic):
    """n = 384 .. 2048; k-blocks in groups of 1 / 2 / 4 / 8; cyclic and contiguous row blocks; with path tables."""
    monkeypatch.setenv("FW_MULTI_GROUP", str(G))
    monkeypatch.setenv("FW_MULTI_CYCLIC", cyclic)
    rate, nxt = graphs.exchange_graph(E, C, seed=33, density=0.7, mode=mode)
    ref = O.solve_dense(rate, nxt, paths=True, threads=0)
    ms = sharded.MultiSolver(devices=[0] * world)
    try:
        assert_same(ms.solve_dense(rate, nxt, paths=True), ref, True)
        assert_same(ms.solve_dense(rate, nxt, paths=False), ref, False)
        t, launches = ms.last_solve()
        assert t > 0 and launches > 0
    finally:
        ms.close()


@pytest.mark.parametrize("world,n_e,n_c", [(2, 30, 10), (3, 43, 3), (4, 17, 16)])
def test_padded_orders(world, n_e, n_c):
    """n = 300, 129, 272: not a multiple of 128 * world -- padded inside the library."""
    rate, nxt = graphs.exchange_graph(n_e, n_c, seed=5, density=0.8)
    ref = O.solve_dense(rate, nxt, paths=True, threads=0)
    ms = sharded.MultiSolver(devices=[0] * world)
    try:
        assert_same(ms.solve_dense(rate, nxt, paths=True), ref, True)
    finally:
        ms.close()


@pytest.mark.parametrize("mode", ["consistent", "arbitrage"])
def test_sync_and_optimum_across_shards(mode):
    """COO in (buildMatrix per shard), matrix resident in 4 shards, `optimum` read across them: rates and the exact
    `_path` of the literal oracle twin; dense read-back equals the oracle."""
    E, C = 24, 8                                   # n = 192 -> padded to 512 over 4 ranks
    blocks = graphs.exchange_blocks(E, C, 11, density=0.7, mode=mode)
    ex_rates_named = graphs.rates_map_from_blocks(blocks)
    from floydwarshall_b200.types import Vertex
    ex_rates = {(Vertex(*a), Vertex(*b)): v for (a, b), v in ex_rates_named.items()}
    vertices, ccy, src, dst, val = algorithms.coo(ex_rates)
    n = len(vertices)
    _v, rate0, next0 = algorithms.pack(ex_rates)            # host-side buildMatrix (Algorithms.hs:26-40)
    ref = O.solve_dense(rate0, next0, paths=True, threads=0)
    ms = sharded.MultiSolver(devices=[0, 0, 0, 0])
    try:
        ms.sync(n, ccy, src, dst, val, paths=True)
        got = ms.download(0, n, want=("rate", "next", "init_next", "mid", "csT", "rs"))
        assert np.array_equal(bits(got["rate"]), bits(ref.rate)) and np.array_equal(got["next"], ref.next)
        assert np.array_equal(got["init_next"], next0)
        assert np.array_equal(got["mid"], ref.mid) and np.array_equal(got["csT"], ref.csT) and np.array_equal(got["rs"], ref.rs)
        rng = np.random.default_rng(3)
        for i, j in rng.integers(0, n, size=(40, 2)):
            i, j = int(i), int(j)
            r, path = ms.optimum(i, j)
            assert r == ref.rate[i, j] or (np.isnan(r) and np.isnan(ref.rate[i, j]))
            want = [] if i == j else O.reconstruct_path(i, j, next0, ref.mid, ref.csT, ref.rs)
            assert path == want
        # a partial row range
        part = ms.download(100, 60)
        assert np.array_equal(bits(part["rate"]), bits(ref.rate[100:160]))
    finally:
        ms.close()


def test_domain_error_and_error_text():
    rate, nxt = graphs.exchange_graph(20, 16, seed=1)
    bad = rate.copy(); bad[200, 7] = -1.0
    ms = sharded.MultiSolver(devices=[0, 0])
    try:
        with pytest.raises(_lib.FwError) as ei:
            ms.solve_dense(bad, nxt)
        assert ei.value.code == _lib.FW_ERR_DOMAIN and "negative" in ei.value.msg
        ref = O.solve_dense(rate, nxt, threads=0)
        assert_same(ms.solve_dense(rate, nxt), ref, False)      # the object is still usable
    finally:
        ms.close()


def test_row_replay_n8192_four_virtual_ranks():
    """Default policy at N=8192 over 4 ranks (groups of 4, cyclic): sampled rows replayed by the oracle from the
    recorded pivot rows -- final rates / next / mid, csT rows, pivot rows and rs rows."""
    n = 8192
    rate, nxt = graphs.exchange_graph(n // 16, 16, seed=1305)
    ms = sharded.MultiSolver(devices=[0, 0, 0, 0])
    try:
        ms.record_row_snapshots(True)
        res = ms.solve_dense(rate, nxt, paths=True)
        S = ms.download_sink(0, n)
        rng = np.random.default_rng(9)
        rows = np.array(sorted(set([0, 127, 128, 511, 512, 513, 2047, 2048, n - 513, n - 1] + rng.integers(0, n, 54).tolist())),
                        dtype=np.int32)
        rp = O.replay_rows(rows, S, rate[rows], nxt[rows], threads=0)
        assert np.array_equal(bits(res.rate[rows]), bits(rp.rate))
        assert np.array_equal(res.next[rows], rp.next) and np.array_equal(res.mid[rows], rp.mid)
        assert np.array_equal(res.csT[rows], rp.csT) and np.array_equal(res.rs[rows], rp.mid_at_i)
        at = rp.at_i.copy()
        at[np.arange(len(rows)), rows] = 0.0
        assert np.array_equal(bits(S[rows]), bits(at))
    finally:
        ms.close()


# ---------------------------------------------------------------- real multi-GPU (skipped on a one-GPU box)
def _ngpu():
    return _lib.load().fw_device_count()


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_real_devices_one_process(monkeypatch, transport):
    ng = _ngpu()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("FW_MULTI_TRANSPORT", transport)
    rate, nxt = graphs.exchange_graph(128, 16, seed=41, density=0.7)      # n = 2048
    ref = O.solve_dense(rate, nxt, paths=True, threads=0)
    ms = sharded.MultiSolver(devices=list(range(min(ng, 8))))
    try:
        assert_same(ms.solve_dense(rate, nxt, paths=True), ref, True)
        ms.solve_dense(rate, nxt, paths=True)
        r, path = ms.optimum(5, 1777)
        assert r == ref.rate[5, 1777] and path == O.reconstruct_path(5, 1777, nxt, ref.mid, ref.csT, ref.rs)
    finally:
        ms.close()


def test_torchrun_one_process_per_gpu():
    """bench.py --gpus 2 through torchrun: fw_multi_create_rank + NCCL, checked against the oracle in-run."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(root, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
                          "--order", "4096", "--skip-e2e"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and "bit-exact" in d["config"]["check"]
