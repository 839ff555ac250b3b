"""world_size-2 gloo test (CPU) of the row-sharded schedule: ownership, broadcast roots, local
row offsets.  The kernels are replaced by a numpy backend with the same data flow; the result
must equal the single-process oracle bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, block, mode, out_dir, lookahead=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from floydwarshall_b200 import graphs, sharded
    from np_shard_backend import NumpyShardBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rate, nxt = graphs.exchange_graph(n // 8, 8, seed=21, density=0.8, mode=mode)
    rows = n // world
    r = rate[rank * rows:(rank + 1) * rows].copy()
    x = nxt[rank * rows:(rank + 1) * rows].copy()
    be = NumpyShardBackend(n, rank * rows, r, x, block=block)

    def bcast(owner):
        t = torch.from_numpy(be.Rw)
        dist.broadcast(t, src=owner)

    def bcast2(buf, owner):
        t = torch.from_numpy(be.Rw2[buf])
        dist.broadcast(t, src=owner)

    sharded.B = block      # the schedule is block-size agnostic; shrink it so the test is fast
    try:
        if isinstance(lookahead, str):
            sharded.run_schedule_lookahead_groups(be, n, rank, world, sharded.SerialRuntime(bcast2),
                                                  {"pairs": 2, "quads": 4}[lookahead])
        elif lookahead:
            sharded.run_schedule_lookahead(be, n, rank, world, sharded.SerialRuntime(bcast2))
        else:
            sharded.run_schedule(be, n, rank, world, bcast)
    finally:
        sharded.B = 128
    be.finish()
    np.save(os.path.join(out_dir, f"rate{rank}.npy"), r)
    np.save(os.path.join(out_dir, f"next{rank}.npy"), x)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("lookahead", [False, True, "pairs", "quads"])
@pytest.mark.parametrize("mode", ["consistent", "arbitrage"])
def test_two_rank_schedule_matches_oracle(tmp_path, mode, lookahead):
    from floydwarshall_b200 import graphs
    from oracle import fw_oracle as O
    n, block, world = 64, 8, 2
    port = 29500 + (os.getpid() % 2000) + {False: 0, True: 7, "pairs": 13, "quads": 19}[lookahead]
    mp.spawn(_worker, args=(world, port, n, block, mode, str(tmp_path), lookahead), nprocs=world, join=True)
    rate, nxt = graphs.exchange_graph(n // 8, 8, seed=21, density=0.8, mode=mode)
    ref = O.solve_dense(rate, nxt)
    got_r = np.concatenate([np.load(tmp_path / f"rate{r}.npy") for r in range(world)])
    got_x = np.concatenate([np.load(tmp_path / f"next{r}.npy") for r in range(world)])
    assert np.array_equal(got_r.view(np.uint64), ref.rate.view(np.uint64))
    assert np.array_equal(got_x, ref.next)


def test_shard_rows_validation():
    from floydwarshall_b200 import sharded
    assert sharded.shard_rows(65536, 8) == 8192
    with pytest.raises(ValueError):
        sharded.shard_rows(1000, 2)


def test_shard_group_policy():
    """Pairs by default wherever a pair never straddles two ranks; the plain per-block schedule otherwise."""
    from floydwarshall_b200 import sharded
    assert sharded.shard_group(65536, 8) == 2 and sharded.shard_group(65536, 2) == 2
    assert sharded.shard_group(1024, 8) == 1          # 128 rows per rank: no room for a pair
    assert sharded.shard_group(256, 1) == 1           # a single pair: nothing to look ahead to
