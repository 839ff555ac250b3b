"""Row-sharded entry points (fw_shard_*) on ONE GPU: the P shards of a matrix are driven in
sequence through the same schedule the multi-GPU path uses (the Rw panel is copied between
the shard backends instead of NCCL-broadcast), and the gathered result must equal the CPU
oracle bit for bit.  (No inter-waiting kernels: everything is stream-ordered on one device.)"""
import numpy as np
import pytest

from floydwarshall_b200 import _lib, graphs, sharded
from oracle import fw_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lookahead", [False, True])
@pytest.mark.parametrize("world,E,C,mode", [(1, 24, 16, "consistent"), (2, 32, 16, "consistent"),
                                            (4, 32, 16, "arbitrage"), (2, 16, 16, "pow2")])
def test_virtual_ranks_on_one_gpu(world, E, C, mode, lookahead):
    import torch
    n = E * C
    rate, nxt = graphs.exchange_graph(E, C, seed=31, density=0.7, mode=mode)
    ref = O.solve_dense(rate, nxt, threads=0)
    rows = sharded.shard_rows(n, world)
    ctxs = [_lib.Context(0) for _ in range(world)]
    stream = torch.cuda.current_stream().cuda_stream
    bes = []
    for r in range(world):
        ctxs[r].set_stream(stream)
        rt = torch.from_numpy(rate[r * rows:(r + 1) * rows].copy()).cuda()
        xt = torch.from_numpy(nxt[r * rows:(r + 1) * rows].copy()).cuda()
        be = sharded.GpuShardBackend(ctxs[r], n, r * rows, rt, xt)
        be.validate()
        bes.append(be)
    if not lookahead:
        for b0 in range(0, n, sharded.B):
            owner = b0 // rows
            bes[owner].pivot(b0)
            for r in range(world):
                if r != owner:
                    bes[r].Rw.copy_(bes[owner].Rw)      # stands in for dist.broadcast(Rw, src=owner)
            for r in range(world):
                bes[r].update(b0)
    else:
        # the look-ahead call sequence (fw_shard_update_ex modes 1/2), serialised on one stream
        nblk = n // sharded.B
        bes[0].pivot(0, 0)
        for r in range(1, world):
            bes[r].Rw2[0].copy_(bes[0].Rw2[0])
        for b in range(nblk):
            b0, buf = b * sharded.B, b & 1
            nxt = b + 1 < nblk
            on = (b0 + sharded.B) // rows if nxt else -1
            if nxt:
                lr = (b0 + sharded.B) - on * rows
                bes[on].update(b0, buf, 1, lr)
                bes[on].pivot(b0 + sharded.B, buf ^ 1)
                for r in range(world):
                    if r != on:
                        bes[r].Rw2[buf ^ 1].copy_(bes[on].Rw2[buf ^ 1])
            for r in range(world):
                if r == on:
                    bes[r].update(b0, buf, 2, (b0 + sharded.B) - r * rows)
                else:
                    bes[r].update(b0, buf, 0, 0)
    torch.cuda.synchronize()
    got_r = np.concatenate([be.rate.cpu().numpy() for be in bes])
    got_x = np.concatenate([be.next.cpu().numpy() for be in bes])
    assert np.array_equal(got_r.view(np.uint64), ref.rate.view(np.uint64))
    assert np.array_equal(got_x, ref.next)
    for c in ctxs:
        c.close()


@pytest.mark.parametrize("G,world,E,C,mode", [(2, 1, 32, 16, "consistent"), (2, 2, 32, 16, "arbitrage"),
                                              (2, 2, 64, 16, "consistent"), (2, 4, 64, 16, "pow2"),
                                              (4, 1, 64, 16, "arbitrage"), (4, 2, 64, 16, "consistent"),
                                              (4, 2, 128, 16, "consistent"), (8, 1, 128, 16, "consistent"),
                                              (8, 2, 128, 16, "arbitrage")])
def test_virtual_ranks_groups_on_one_gpu(G, world, E, C, mode):
    """The call sequence of sharded.run_schedule_lookahead_groups (fw_shard_update_group: G k-blocks per
    bulk launch), serialised on one stream; n = 512 / 1024 / 2048 so that a rank owns one or two groups."""
    import torch
    n = E * C
    B = sharded.B
    rate, nxt = graphs.exchange_graph(E, C, seed=33, density=0.7, mode=mode)
    ref = O.solve_dense(rate, nxt, threads=0)
    rows = sharded.shard_rows(n, world)
    assert rows % (G * B) == 0
    ctxs = [_lib.Context(0) for _ in range(world)]
    stream = torch.cuda.current_stream().cuda_stream
    bes = []
    for r in range(world):
        ctxs[r].set_stream(stream)
        rt = torch.from_numpy(rate[r * rows:(r + 1) * rows].copy()).cuda()
        xt = torch.from_numpy(nxt[r * rows:(r + 1) * rows].copy()).cuda()
        be = sharded.GpuShardBackend(ctxs[r], n, r * rows, rt, xt)
        be.validate()
        bes.append(be)

    def share(buf, owner):                      # stands in for dist.broadcast(Rw[buf], src=owner)
        for r in range(world):
            if r != owner:
                bes[r].Rw2[buf].copy_(bes[owner].Rw2[buf])

    def factor(p):
        b0, s = G * p * B, G * (p & 1)
        ow = b0 // rows
        for j in range(G):
            if j > 0:
                bes[ow].update_group(b0, j, s, 1, (b0 + j * B) - ow * rows, B)
            bes[ow].pivot(b0 + j * B, s + j)
            share(s + j, ow)

    ngrp = n // (G * B)
    factor(0)
    for p in range(ngrp):
        b0, s = G * p * B, G * (p & 1)
        on = (b0 + G * B) // rows if p + 1 < ngrp else -1
        if on >= 0:
            bes[on].update_group(b0, G, s, 1, (b0 + G * B) - on * rows, G * B)
            factor(p + 1)
        for r in range(world):
            if r == on:
                bes[r].update_group(b0, G, s, 2, (b0 + G * B) - r * rows, G * B)
            else:
                bes[r].update_group(b0, G, s, 0, 0, 0)
        ow = b0 // rows
        for i in range(G - 1):
            bes[ow].update_group(b0 + (i + 1) * B, G - 1 - i, s + i + 1, 1, (b0 + i * B) - ow * rows, B)
    torch.cuda.synchronize()
    got_r = np.concatenate([be.rate.cpu().numpy() for be in bes])
    got_x = np.concatenate([be.next.cpu().numpy() for be in bes])
    assert np.array_equal(got_r.view(np.uint64), ref.rate.view(np.uint64))
    assert np.array_equal(got_x, ref.next)
    for c in ctxs:
        c.close()


def test_device_graph_shard_matches_host_generator():
    import torch
    n, ccy = 512, 16
    rate, nxt = graphs.exchange_graph(n // ccy, ccy, seed=77)
    for row0, rows in ((0, 512), (128, 256), (384, 128)):
        r, x = sharded.device_graph_shard(n, ccy, 77, row0, rows, torch.device("cuda", 0))
        assert np.array_equal(r.cpu().numpy(), rate[row0:row0 + rows])
        assert np.array_equal(x.cpu().numpy(), nxt[row0:row0 + rows])


def test_two_process_nccl_if_two_gpus(tmp_path):
    """Real 2-rank NCCL run when the box has >= 2 GPUs (skipped otherwise)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import subprocess, sys, os, json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29611",
                          os.path.join(root, "bench.py"), "--gpus", "2", "--steps", "1", "--warmup", "1",
                          "--order", "2048", "--skip-e2e", "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["config"].get("check") == "bit-exact vs single-GPU fw_solve_device"
