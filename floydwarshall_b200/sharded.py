"""Row-block-sharded solve across the GPUs of one box (config C5).

One process per GPU (torch.distributed, NCCL).  Rank r keeps rows
[r*n/P, (r+1)*n/P) of `rate` and `next`.  Per k-block of 128 pivots:

    owner of the pivot rows : diagonal tile + row panel  -> Rw (128 x n, fp64)
    all ranks               : broadcast Rw from the owner          (the one collective)
    all ranks               : column panel + bulk update on the local rows

Only Rw travels: next-hops are row-local (NX[i][j] <- NX[i][k]), and the column
snapshots every rank needs for its own rows come out of its own column panel.
By default the k-blocks are taken in PAIRS with a look-ahead lane
(run_schedule_lookahead_pairs / fw_shard_update_pair): two broadcasts per pair,
one fused bulk launch per pair and rank.
The schedule is backend-agnostic: the GPU backend drives libfwgpu's fw_shard_*
entry points; tests drive the same schedule with a numpy backend under gloo.
"""
from __future__ import annotations

import ctypes
import json
import os
import time
from typing import Callable

import numpy as np

from . import _lib

B = _lib.FW_TILE
LOOKAHEAD = os.environ.get("FW_LOOKAHEAD", "1") != "0"   # pivot-panel look-ahead on a second stream
MAX_GROUP = 8                                            # most k-blocks per fused bulk launch (fw::BULK_MAXNB)


def shard_rows(n: int, world: int) -> int:
    if n % (B * world) != 0:
        raise ValueError(f"n={n} must be a multiple of {B}*world_size={B * world}")
    return n // world


def run_schedule(backend, n: int, rank: int, world: int, bcast: Callable[[int], None],
                 k_blocks: range | None = None) -> None:
    """The k-block loop.  `bcast(owner)` broadcasts backend's Rw panel from rank `owner`."""
    rows = shard_rows(n, world)
    for b0 in (k_blocks if k_blocks is not None else range(0, n, B)):
        owner = b0 // rows
        if rank == owner:
            backend.pivot(b0)
        if world > 1:
            bcast(owner)
        backend.update(b0)


def run_schedule_lookahead(backend, n: int, rank: int, world: int, rt) -> None:
    """Same result as run_schedule, but the owner of k-block b+1 brings that block's 128 pivot rows
    up to date FIRST (on the look-ahead lane), factors them and starts their broadcast while every
    rank is still busy with the bulk of k-block b.  `rt` provides the two lanes:

        rt.lane_b()            context manager: work issued inside goes to the look-ahead lane
        rt.a_done() / rt.wait_a_done()   main lane finished update(b)  /  look-ahead lane waits for it
        rt.b_done() / rt.wait_b_done()   Rw(b+1) has arrived           /  main lane waits for it
        rt.bcast(buf, owner)   broadcast backend.Rw[buf] from `owner` (collective, look-ahead lane)

    Program order is a valid serial order, so a synchronous runtime (tests) gives the same answer.
    """
    rows = shard_rows(n, world)
    nblk = n // B
    if rank == 0:
        with rt.lane_b():
            backend.pivot(0, 0)
    if world > 1:
        with rt.lane_b():
            rt.bcast(0, 0)
    rt.b_done()
    for b in range(nblk):
        b0, buf = b * B, b & 1
        nxt = b + 1 < nblk
        own_next = nxt and ((b0 + B) // rows == rank)
        lr_next = (b0 + B) - rank * rows
        rt.wait_b_done()                      # Rw(b) is here (and the early rows are current)
        if nxt:
            with rt.lane_b():
                rt.wait_a_done()              # update(b-1) finished: rows of b+1 and buffer buf^1 are free
                if own_next:
                    backend.update(b0, buf, 1, lr_next)       # only the next pivot rows
                    backend.pivot(b0 + B, buf ^ 1)
                if world > 1:
                    rt.bcast(buf ^ 1, (b0 + B) // rows)
                rt.b_done()
        backend.update(b0, buf, 2 if own_next else 0, lr_next if own_next else 0)
        rt.a_done()


def run_schedule_lookahead_groups(backend, n: int, rank: int, world: int, rt, G: int = 2) -> None:
    """run_schedule_lookahead with the k-blocks taken in GROUPS of G consecutive blocks, so that the bulk
    kernel loads every tile of the shard once per G*128 steps (backend.update_group).  A group never
    straddles two ranks (rows per rank is a multiple of G*B).  Per group its owner factors the G blocks on
    the look-ahead lane -- for block j: its 128 rows take the group's blocks 0..j-1, pivot, broadcast --
    after bringing the group's rows up to date with the previous group; the main lane gives every other
    local row all G blocks in one update_group, and the owner's rows of block i take the blocks after i on
    their own.  Panel buffers: group p uses backend.Rw buffers G*(p&1) .. G*(p&1)+G-1."""
    rows = shard_rows(n, world)
    if rows % (G * B) != 0:
        raise ValueError(f"groups of {G} need rows per rank ({rows}) to be a multiple of {G * B}")
    ngrp = n // (G * B)

    def factor(p):
        b0, s = G * p * B, G * (p & 1)
        owner = b0 // rows
        for j in range(G):
            if rank == owner:
                if j > 0:                                      # rows of block j take blocks 0..j-1 of the group
                    backend.update_group(b0, j, s, 1, (b0 + j * B) - rank * rows, B)
                backend.pivot(b0 + j * B, s + j)
            if world > 1:
                rt.bcast(s + j, owner)

    with rt.lane_b():
        factor(0)
    rt.b_done()
    for p in range(ngrp):
        b0, s = G * p * B, G * (p & 1)
        nxt = p + 1 < ngrp
        own = (b0 // rows == rank)
        own_next = nxt and ((b0 + G * B) // rows == rank)
        lr_next = (b0 + G * B) - rank * rows
        rt.wait_b_done()                      # all panels of group p are here
        if nxt:
            with rt.lane_b():
                rt.wait_a_done()              # group p-1 is finished: the rows of group p+1 and its buffers are free
                if own_next:
                    backend.update_group(b0, G, s, 1, lr_next, G * B)     # only the next group's rows
                factor(p + 1)
                rt.b_done()
        backend.update_group(b0, G, s, 2 if own_next else 0, lr_next if own_next else 0, G * B if own_next else 0)
        if own:
            for i in range(G - 1):            # the rows of block i take the blocks after it
                backend.update_group(b0 + (i + 1) * B, G - 1 - i, s + i + 1, 1, (b0 + i * B) - rank * rows, B)
        rt.a_done()


def run_schedule_lookahead_pairs(backend, n: int, rank: int, world: int, rt) -> None:
    run_schedule_lookahead_groups(backend, n, rank, world, rt, 2)


class SerialRuntime:
    """Both lanes are the caller's thread (CPU tests): events are no-ops."""

    def __init__(self, bcast):
        self._bcast = bcast

    class _Null:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    def lane_b(self):
        return self._Null()

    def a_done(self): pass
    def wait_a_done(self): pass
    def b_done(self): pass
    def wait_b_done(self): pass

    def bcast(self, buf, owner):
        self._bcast(buf, owner)


class TorchRuntime:
    """Main lane = the current stream, look-ahead lane = a second CUDA stream; NCCL broadcast."""

    def __init__(self, backend, world):
        import torch
        self.torch = torch
        self.be = backend
        self.world = world
        self.sa = torch.cuda.current_stream()
        self.sb = torch.cuda.Stream(priority=-1)   # high priority: its CTAs cut in front of the bulk kernel
        self.sb.wait_stream(self.sa)
        self.ev_a = None
        self.ev_b = None

    def lane_b(self):
        rt = self

        class _Lane:
            def __enter__(self_inner):
                rt.be.ctx.set_stream(rt.sb.cuda_stream)
                self_inner.cm = rt.torch.cuda.stream(rt.sb)
                self_inner.cm.__enter__()

            def __exit__(self_inner, *a):
                self_inner.cm.__exit__(*a)
                rt.be.ctx.set_stream(rt.sa.cuda_stream)
                return False

        return _Lane()

    def a_done(self):
        self.ev_a = self.torch.cuda.Event()
        self.ev_a.record(self.sa)

    def wait_a_done(self):
        if self.ev_a is not None:
            self.sb.wait_event(self.ev_a)

    def b_done(self):
        self.ev_b = self.torch.cuda.Event()
        self.ev_b.record(self.sb)

    def wait_b_done(self):
        if self.ev_b is not None:
            self.sa.wait_event(self.ev_b)

    def bcast(self, buf, owner):
        import torch.distributed as dist
        dist.broadcast(self.be.Rw2[buf], src=owner)

    def finish(self):
        self.sa.wait_stream(self.sb)
        self.be.ctx.set_stream(self.sa.cuda_stream)


class GpuShardBackend:
    """Local rows of the matrix as torch CUDA tensors + the fw_shard_* entry points."""

    def __init__(self, ctx: _lib.Context, n: int, row0: int, rate_t, next_t):
        import torch
        assert rate_t.is_cuda and rate_t.is_contiguous() and next_t.is_contiguous()
        assert rate_t.shape[1] == n and rate_t.dtype == torch.float64 and next_t.dtype == torch.int32
        self.ctx, self.n, self.row0, self.rows = ctx, n, row0, rate_t.shape[0]
        self.rate, self.next = rate_t, next_t
        self.Rw2 = [torch.empty((B, n), dtype=torch.float64, device=rate_t.device) for _ in range(2 * MAX_GROUP)]
        self.Rw = self.Rw2[0]
        self.L = _lib.load()
        self.launches = 0

    def _p(self, t):
        return ctypes.c_void_p(t.data_ptr())

    def validate(self):
        _lib.check(self.L.fw_shard_validate(self.ctx.handle, self.n, self.row0, self.rows, self.n,
                                            self._p(self.rate), self._p(self.next)))

    def pivot(self, b0: int, buf: int = 0):
        _lib.check(self.L.fw_shard_pivot(self.ctx.handle, self.n, self.row0, self.rows, self.n,
                                         self._p(self.rate), self._p(self.next), b0, self._p(self.Rw2[buf])))
        self.launches += self.ctx.last_launches

    def update(self, b0: int, buf: int = 0, mode: int = 0, lr0: int = 0):
        _lib.check(self.L.fw_shard_update_ex(self.ctx.handle, self.n, self.row0, self.rows, self.n,
                                             self._p(self.rate), self._p(self.next), b0, self._p(self.Rw2[buf]),
                                             mode, lr0))
        self.launches += self.ctx.last_launches

    def update_group(self, b0: int, nb: int, buf: int = 0, mode: int = 0, lr0: int = 0, lrn: int = 0):
        """nb consecutive k-blocks from b0 with the panels Rw2[buf .. buf+nb-1] (fw_shard_update_group)."""
        arr = (ctypes.c_void_p * nb)(*[self.Rw2[buf + i].data_ptr() for i in range(nb)])
        _lib.check(self.L.fw_shard_update_group(self.ctx.handle, self.n, self.row0, self.rows, self.n,
                                                self._p(self.rate), self._p(self.next), b0, nb, arr, mode, lr0, lrn))
        self.launches += self.ctx.last_launches

    def update_pair(self, b0: int, buf: int = 0, mode: int = 0, lr0: int = 0, lrn: int = 0):
        self.update_group(b0, 2, buf, mode, lr0, lrn)


PAIRS = os.environ.get("FW_SHARD_PAIRS", "1") != "0"     # k-blocks in groups when the shard geometry allows it
# k-blocks per fused bulk launch of the sharded solve.  Measured at 8 GPUs, N=65536: pairs 4513 ms, groups of 8
# 4597 ms (unpaired 4906 ms): the owner brings its group's own rows through the group in 128-row slices on the
# look-ahead lane, and with 8 blocks per group those small launches outweigh the saved tile loads.
GROUP = int(os.environ.get("FW_SHARD_GROUP", "2"))


def shard_group(n: int, world: int) -> int:
    """k-blocks per fused bulk launch of the sharded solve (1 = the plain per-block schedule)."""
    rows = shard_rows(n, world)
    if not PAIRS:
        return 1
    for g in ((GROUP, 2) if GROUP in (1, 2, 4, 8) else (8, 4, 2)):
        if g == 1 or (rows % (g * B) == 0 and n // (g * B) >= 2):
            return g
    return 1


def solve_shard(backend, n: int, rank: int, world: int, lookahead: bool = True):
    """Run the k-block schedule on a GpuShardBackend (collective: call on every rank)."""
    import torch.distributed as dist
    if lookahead:
        rt = TorchRuntime(backend, world)
        G = shard_group(n, world)
        if G > 1:
            run_schedule_lookahead_groups(backend, n, rank, world, rt, G)
        else:
            run_schedule_lookahead(backend, n, rank, world, rt)
        rt.finish()
    else:
        run_schedule(backend, n, rank, world, lambda owner: dist.broadcast(backend.Rw, src=owner))


def solve_sharded_device(ctx: _lib.Context, n: int, rate_t, next_t, validate: bool = True):
    """Solve in place on this rank's row shard (torch CUDA tensors).  Collective: call on every rank."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    rows = shard_rows(n, world)
    be = GpuShardBackend(ctx, n, rank * rows, rate_t, next_t)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if validate:
        be.validate()
    solve_shard(be, n, rank, world, lookahead=LOOKAHEAD)
    return be


def solve_sharded_host(ctx: _lib.Context, n: int, rate_h, next_h, work_rate=None, work_next=None):
    """Public multi-GPU entry on HOST shards (pinned torch CPU tensors [rows, n]): H2D, solve, D2H."""
    import torch
    dev = torch.device("cuda", ctx.device)
    r = work_rate if work_rate is not None else torch.empty(rate_h.shape, dtype=torch.float64, device=dev)
    x = work_next if work_next is not None else torch.empty(next_h.shape, dtype=torch.int32, device=dev)
    r.copy_(rate_h, non_blocking=True)
    x.copy_(next_h, non_blocking=True)
    be = solve_sharded_device(ctx, n, r, x)
    rate_h.copy_(r, non_blocking=True)
    next_h.copy_(x, non_blocking=True)
    torch.cuda.synchronize()
    return be


# --------------------------------------------------------------------------
def device_graph_shard(n: int, ccy: int, seed: int, row0: int, rows: int, device):
    """Rows [row0, row0+rows) of buildMatrix of the synthetic E x C graph, built in HBM."""
    import torch
    from . import graphs
    E, C = n // ccy, ccy
    e0, El = row0 // C, rows // C
    blocks = torch.from_numpy(graphs.exchange_blocks(E, C, seed)[e0:e0 + El]).to(device)   # [El,C,C]
    rate = torch.zeros((rows, n), dtype=torch.float64, device=device)
    nxt = torch.full((rows, n), -1, dtype=torch.int32, device=device)
    r4 = rate.view(El, C, E, C)
    n4 = nxt.view(El, C, E, C)
    cols = torch.arange(n, dtype=torch.int32, device=device).view(E, C)
    for c in range(C):
        r4[:, c, :, c] = 1.0
        n4[:, c, :, c] = cols[None, :, c]
    el = torch.arange(El, device=device)
    r4[el, :, el + e0, :] = blocks
    n4[el, :, el + e0, :] = torch.where(blocks != 0, cols[e0:e0 + El][:, None, :].expand(El, C, C),
                                        torch.full((), -1, dtype=torch.int32, device=device))
    li = torch.arange(rows, device=device)
    rate[li, li + row0] = 0.0
    nxt[li, li + row0] = -1
    return rate, nxt


def bench_main(args, METRIC, UNIT, SEED, workload_n, workload_name, ClockSampler, measured_fp64_peak):
    """bench.py --gpus N (N > 1): config C5, strong scaling, one rank per GPU."""
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    n = workload_n(world, args.order)
    rows = shard_rows(n, world)
    row0 = rank * rows
    peak_tflops, peak_src, peak_raw = measured_fp64_peak() if rank == 0 else (None, None, None)

    ctx = _lib.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    r0, x0 = device_graph_shard(n, 16, SEED + 1, row0, rows, dev)
    r = torch.empty_like(r0)
    x = torch.empty_like(x0)
    be = GpuShardBackend(ctx, n, row0, r, x)

    def step():
        r.copy_(r0)
        x.copy_(x0)
        be.validate()
        solve_shard(be, n, rank, world, lookahead=LOOKAHEAD)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    be.launches = 0
    sampler = ClockSampler(local)
    dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = float(n) ** 3 / (ms_per_step * 1e-3)
    launches = torch.tensor([be.launches], dtype=torch.int64, device=dev)
    dist.all_reduce(launches)

    check = None
    if getattr(args, "check", False):
        from . import dense
        full_r = [torch.empty_like(r) for _ in range(world)]
        full_x = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(full_r, r)
        dist.all_gather(full_x, x)
        p_r = [torch.empty_like(r0) for _ in range(world)]
        p_x = [torch.empty_like(x0) for _ in range(world)]
        dist.all_gather(p_r, r0)
        dist.all_gather(p_x, x0)
        if rank == 0:
            fr, fx = torch.cat(p_r).contiguous(), torch.cat(p_x).contiguous()
            c2 = _lib.Context(local)
            c2.set_stream(torch.cuda.current_stream().cuda_stream)
            dense.solve_device(c2, fr, fx)
            torch.cuda.synchronize()
            same = bool(torch.equal(fr.view(torch.int64), torch.cat(full_r).view(torch.int64))) and \
                bool(torch.equal(fx, torch.cat(full_x)))
            c2.close()
            if not same:
                raise SystemExit("sharded result differs from the single-GPU solve")
            check = "bit-exact vs single-GPU fw_solve_device"

    # per-phase profile of one step on this rank (bulk kernel roofline)
    ctx.set_profiling(True)
    r.copy_(r0); x.copy_(x0)
    bulk_ms, bulk_cnt = 0.0, 0
    for b0 in range(0, n, B):
        owner = b0 // rows
        if rank == owner:
            be.pivot(b0)
        dist.broadcast(be.Rw, src=owner)
        be.update(b0)
        ms, cnt = ctx.phase_ms()
        bulk_ms += ms[3]; bulk_cnt += cnt[3]
    ctx.set_profiling(False)

    # e2e: host shards through solve_sharded_host (H2D + solve + D2H), wall clock, max over ranks
    e2e = None
    if not args.skip_e2e:
        rh = torch.empty((rows, n), dtype=torch.float64, pin_memory=True)
        xh = torch.empty((rows, n), dtype=torch.int32, pin_memory=True)
        ts = []
        for it in range(2):
            rh.copy_(r0); xh.copy_(x0)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            solve_sharded_host(ctx, n, rh, xh, work_rate=r, work_next=x)
            dist.barrier()
            t1 = time.perf_counter()
            if it > 0:
                ts.append(t1 - t0)
        tt = torch.tensor([float(np.mean(ts))], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": float(n) ** 3 / float(tt.item()), "unit": UNIT,
               "h2d_bytes_per_step": n * n * 12, "d2h_bytes_per_step": n * n * 12,
               "ms_per_step": float(tt.item()) * 1e3,
               "api": "sharded.solve_sharded_host (pinned host row shards, one rank per GPU)"}

    if rank == 0:
        relax_per_launch = float(rows) * (n - B) * B      # non-owner launch; owner launches cover 128 rows fewer
        ach = (2.0 * relax_per_launch / (bulk_ms / max(bulk_cnt, 1) * 1e-3) / 1e12) if bulk_cnt else None
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "n": n, "seed": SEED + 1, "k_block": B,
                       "sharding": f"row blocks of {rows} rows per rank; per-k-block NCCL broadcast of the "
                                   f"128 x {n} fp64 pivot-row snapshot panel ({B * n * 8 / 2**20:.0f} MiB)",
                       "lookahead": LOOKAHEAD, "k_blocks_per_bulk_launch": (shard_group(n, world) if LOOKAHEAD else 1),
                       "l2": "per-rank inputs are far larger than the 126 MB L2; no flush needed",
                       "check": check},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches.item()),
            "roofline": {"bound": "fp64", "kernel": "fw_bulk_kernel", "achieved": ach, "peak": peak_tflops,
                         "unit": "TFLOP/s", "frac": (ach / peak_tflops) if ach else None, "traffic": None,
                         "peak_source": peak_src, "avg_launch_ms": bulk_ms / max(bulk_cnt, 1),
                         "note": "rank 0's launches; per-GPU figure"},
            "cpu_baseline": None, "fp64_peak_probe": peak_raw,
        }))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
