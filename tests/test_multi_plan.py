"""CPU tests of the multi-GPU schedule itself (fw_multi_plan: the operation list the CUDA executor issues),
replayed with the numpy model of the operations: as virtual ranks in one process over many layouts, and with
world_size 2 under gloo (one process per rank, BCAST = dist.broadcast).  Results must equal the oracle's bits."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

from floydwarshall_b200 import _lib, graphs      # noqa: E402
from oracle import fw_oracle as O                # noqa: E402
import np_plan_backend as NP                     # noqa: E402


def _graph(n, mode, seed=21):
    return graphs.exchange_graph(n // 8, 8, seed=seed, density=0.8, mode=mode)


# (n, world, B, G, cbr): contiguous row blocks (cbr = n / world) and cyclic blocks of one or two groups
LAYOUTS = [(64, 1, 8, 1, 64), (64, 1, 8, 4, 64), (64, 2, 8, 1, 32), (64, 2, 8, 2, 32), (64, 2, 8, 2, 16),
           (64, 2, 8, 4, 32), (128, 4, 8, 2, 16), (128, 4, 8, 4, 32), (128, 2, 8, 8, 64), (96, 3, 8, 2, 16),
           (128, 2, 8, 2, 32), (128, 8, 8, 1, 8), (128, 8, 8, 2, 16)]


@pytest.mark.parametrize("mode", ["consistent", "arbitrage"])
@pytest.mark.parametrize("n,world,B,G,cbr", LAYOUTS)
def test_plan_virtual_ranks_match_oracle(n, world, B, G, cbr, mode):
    rate, nxt = _graph(n, mode)
    ref = O.solve_dense(rate, nxt)
    got_r, got_x = NP.run_virtual(NP.Layout(n, world, B, G, cbr), rate, nxt)
    assert np.array_equal(got_r.view(np.uint64), ref.rate.view(np.uint64))
    assert np.array_equal(got_x, ref.next)


def test_plan_rejects_bad_layouts():
    L = _lib.load()
    assert L.fw_multi_plan(64, 2, 8, 2, 24, None, 0) < 0        # cyclic block not a multiple of the group
    assert L.fw_multi_plan(72, 2, 8, 2, 16, None, 0) < 0        # n not a whole number of cyclic rounds
    assert L.fw_multi_plan(64, 2, 8, 2, 16, None, 0) > 0


def test_plan_structure():
    """Every k-block is pivoted exactly once by the rank that holds its rows, broadcast once, and every rank's
    main lane applies every group exactly once."""
    n, world, B, G, cbr = 256, 4, 8, 2, 16
    lay = NP.Layout(n, world, B, G, cbr)
    ops = NP.get_plan(n, world, B, G, cbr)
    piv = [(o.b0, o.rank) for o in ops if o.kind == _lib.OP_PIVOT]
    assert sorted(b for b, _ in piv) == list(range(0, n, B))
    for b0, r in piv:
        assert (b0 // cbr) % world == r
    assert sorted(o.b0 for o in ops if o.kind == _lib.OP_BCAST) == list(range(0, n, B))
    for r in range(world):
        mains = [o.b0 for o in ops if o.kind == _lib.OP_APPLY and o.lane == 0 and o.rank == r]
        assert mains == list(range(0, n, G * B))
    # ownership rotates: consecutive groups belong to consecutive ranks
    owners = [r for b0, r in piv if b0 % (G * B) == 0]
    assert owners[:8] == [0, 1, 2, 3, 0, 1, 2, 3]
    assert lay.glob(1, 0) == cbr and lay.glob(1, cbr) == cbr * (world + 1)


def _worker(rank, world, port, n, B, G, cbr, mode, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import np_plan_backend as NPW
    from floydwarshall_b200 import _lib as LW
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rate, nxt = _graph(n, mode)
    lay = NPW.Layout(n, world, B, G, cbr)
    me = NPW.NpRank(lay, rank, rate, nxt)
    for op in NPW.get_plan(n, world, B, G, cbr):
        if op.kind == LW.OP_BCAST:                       # every rank takes part, `rank` is the root
            t = torch.from_numpy(me.Rw[op.buf])
            dist.broadcast(t, src=op.rank)
        elif op.rank == rank:
            if op.kind == LW.OP_PIVOT:
                me.pivot(op)
            elif op.kind == LW.OP_APPLY:
                me.apply(op)
    np.save(os.path.join(out_dir, f"rate{rank}.npy"), me.rate)
    np.save(os.path.join(out_dir, f"next{rank}.npy"), me.next)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("G,cbr", [(1, 32), (2, 16), (4, 32)])
@pytest.mark.parametrize("mode", ["consistent", "arbitrage"])
def test_two_rank_gloo_plan_matches_oracle(tmp_path, mode, G, cbr):
    import torch.multiprocessing as mp
    n, B, world = 64, 8, 2
    port = 29500 + (os.getpid() % 2000) + 10 * G + (3 if mode == "arbitrage" else 0)
    mp.spawn(_worker, args=(world, port, n, B, G, cbr, mode, str(tmp_path)), nprocs=world, join=True)
    rate, nxt = _graph(n, mode)
    ref = O.solve_dense(rate, nxt)
    lay = NP.Layout(n, world, B, G, cbr)
    got_r, got_x = np.empty_like(rate), np.empty_like(nxt)
    for r in range(world):
        idx = lay.local_rows_of(r)
        got_r[idx] = np.load(tmp_path / f"rate{r}.npy")
        got_x[idx] = np.load(tmp_path / f"next{r}.npy")
    assert np.array_equal(got_r.view(np.uint64), ref.rate.view(np.uint64))
    assert np.array_equal(got_x, ref.next)


def test_global_rows_matches_the_plan_layout():
    """sharded.global_rows (what bench.py uses to compare a shard with the oracle) is the layout of fw_plan.hpp."""
    from floydwarshall_b200 import sharded
    for world, rows, cbr in ((2, 32, 16), (4, 64, 16), (8, 1024, 1024), (3, 48, 8)):
        lay = NP.Layout(rows * world, world, 8, 1, cbr)
        for rank in range(world):
            info = _lib.ShardInfo(rank=rank, world=world, rows=rows, cyclic_rows=cbr)
            assert np.array_equal(sharded.global_rows(info), lay.local_rows_of(rank))
    # every global row belongs to exactly one (rank, local row)
    lay = NP.Layout(96, 3, 8, 2, 16)
    allrows = np.concatenate([lay.local_rows_of(r) for r in range(3)])
    assert sorted(allrows.tolist()) == list(range(96))


def test_plan_random_layouts_match_oracle():
    """Random layouts (ranks, k-blocks per group, cyclic block size in groups, number of rounds) with a tiny k-block."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=20, deadline=None)
    @given(world=st.integers(1, 5), G=st.sampled_from([1, 2, 4]), groups_per_block=st.integers(1, 2),
           rounds=st.integers(1, 2), seed=st.integers(0, 1000), mode=st.sampled_from(["consistent", "arbitrage", "ones"]))
    def run(world, G, groups_per_block, rounds, seed, mode):
        B = 4
        cbr = G * B * groups_per_block
        n = cbr * world * rounds
        if n > 192:
            return
        C = 4
        rate, nxt = graphs.exchange_graph(n // C, C, seed=seed, density=0.8, mode=mode)
        ref = O.solve_dense(rate, nxt)
        got_r, got_x = NP.run_virtual(NP.Layout(n, world, B, G, cbr), rate, nxt)
        assert np.array_equal(got_r.view(np.uint64), ref.rate.view(np.uint64))
        assert np.array_equal(got_x, ref.next)

    run()
