#!/usr/bin/env python
"""bench.py -- FW relaxations/s (N^3/t) of the matrix-optimisation hot path.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W`
prints ONE JSON line on rank 0.  A "step" is one full solve of the synthetic
exchange/currency graph BASELINE.json's metric is quoted on:

  N=1  : config C4, N=32768 (2048 exchanges x 16 currencies, dense), fp64
  N>1  : config C5, N=65536 row-block-sharded over the ranks with a per-k-block
         pivot-row-panel broadcast (strong scaling: the problem is fixed)

`value`   whole-job relaxations/s with the matrices resident in HBM (CUDA events)
`e2e`     same metric through the reference-facing C ABI `fw_solve` on HOST buffers
          (pinned), H2D + validation + solve + D2H inside the timed region
`roofline` the dominant kernel (fw_bulk_kernel) against the measured FP64 peak
`cpu_baseline` the CPU oracle (C restatement of the reference loop, OpenMP) on a
          bounded sample of the same workload
`configs` (N=1) / `config.check` (N>1): untimed comparisons of the CUDA results with the
          oracle, computed in the same run.  The oracle is only ever the checker or the CPU
          baseline here; nothing that is timed as "ours" calls it.

`--impl reference` times the reference algorithm's CPU restatement (oracle/; the
Haskell reference itself cannot be built: no ghc/cabal in the image) on the
host cores with the same config/metric/unit.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "fw_relaxations_per_s"
UNIT = "relaxations/s"
SEED = 1234 + 3          # PCG64 seed = 1234 + config index
CCY = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--order", type=int, default=0, help="override the matrix order (debug only)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-paths", action="store_true", help="N=1: leave out the path-table timings")
    ap.add_argument("--skip-configs", action="store_true", help="N=1: leave out the C2 / C3 / N=8192 block")
    ap.add_argument("--skip-check", action="store_true", help="N>1: leave out the pre-flight oracle comparison")
    ap.add_argument("--single-process", action="store_true",
                    help="N>1: rank 0 alone drives all GPUs through fw_multi_create (the in-process call)")
    return ap.parse_args()


def workload_n(gpus: int, override: int) -> int:
    if override:
        return override
    return 32768 if gpus == 1 else 65536


def workload_name(n: int) -> str:
    return f"synthetic {n // CCY} exchanges x {CCY} currencies dense rate graph (N={n}), fp64"


# --------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_fp64_peak():
    """DFMA peak measured on this box by tools/fp64_peak (MEASURED_PEAKS.json has no FP64 figure)."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120).stdout.strip().splitlines()[-1]
        d = json.loads(out)
        return float(d["dfma_tflops"]), "tools/fp64_peak DFMA chain measured in this run", d
    except Exception as ex:  # noqa: BLE001
        return 37.2, f"nominal 148 SM x 64 lanes x 2 x 1.965 GHz (fp64_peak failed: {ex})", None


def device_graph(n: int, seed: int, device):
    """buildMatrix of the synthetic E x C graph, built directly in HBM (torch = plumbing)."""
    import torch
    from floydwarshall_b200 import graphs
    E, C = n // CCY, CCY
    blocks = torch.from_numpy(graphs.exchange_blocks(E, C, seed)).to(device)       # [E,C,C]
    rate = torch.zeros((n, n), dtype=torch.float64, device=device)
    nxt = torch.full((n, n), -1, dtype=torch.int32, device=device)
    r4 = rate.view(E, C, E, C)
    n4 = nxt.view(E, C, E, C)
    cols = torch.arange(n, dtype=torch.int32, device=device).view(E, C)
    for c in range(C):                         # same currency on another exchange: exactly 1.0
        r4[:, c, :, c] = 1.0
        n4[:, c, :, c] = cols[None, :, c]
    rd = r4.diagonal(dim1=0, dim2=2)           # [C, C, E]: the same-exchange blocks
    nd = n4.diagonal(dim1=0, dim2=2)
    rd.copy_(blocks.permute(1, 2, 0))
    nd.copy_(torch.where(blocks.permute(1, 2, 0) != 0, cols.t()[None, :, :].expand(C, C, E).to(torch.int32),
                         torch.full((), -1, dtype=torch.int32, device=device)))
    idx = torch.arange(n, device=device)
    rate[idx, idx] = 0.0
    nxt[idx, idx] = -1
    return rate, nxt


def host_threads() -> int:
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers, which
    would make the CPU arm single-threaded under `--gpus N`; the oracle takes an explicit thread count."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def host_graph(n: int, seed: int):
    from floydwarshall_b200 import graphs
    return graphs.exchange_graph(n // CCY, CCY, seed)


# --------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle's OpenMP loop on all host cores, bounded k-step samples."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fw_oracle as O
    n = workload_n(args.gpus, args.order)
    n_config = n
    try:                                   # the dense host matrix is 12 B per entry (+ generator temporaries)
        import psutil
        avail = psutil.virtual_memory().available
        while n > 4096 and 14.0 * n * n > 0.7 * avail:
            n //= 2
    except Exception:  # noqa: BLE001
        pass
    threads = host_threads()
    rate, nxt = host_graph(n, SEED)
    ksteps = max(1, int(8 * (32768 / n) ** 2))          # ~8.6e9 relaxations per step
    ksteps = min(ksteps, max(1, n // (args.warmup + args.steps)))   # stay inside the n pivots
    t_steps = []
    k = 0
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        O.run_ksteps(rate, nxt, k, k + ksteps, threads)
        t1 = time.perf_counter()
        k += ksteps
        if s >= args.warmup:
            t_steps.append(t1 - t0)
    per = float(np.mean(t_steps))
    value = ksteps * float(n) * n / per
    sample = f"{ksteps} consecutive k-steps of the N={n} matrix per step ({ksteps * n * n:.3e} relaxations)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(n_config), "n": n_config, "n_sampled": n,
                                        "note": "C restatement of the reference loop (oracle/fw_oracle.c, OpenMP "
                                                "over i); the Haskell reference cannot be built here (no ghc)"
                                                + ("" if n == n_config else
                                                   f"; host RAM too small for N={n_config}, k-steps sampled on the "
                                                   f"N={n} graph of the same family (relaxations/s is per-relaxation)")},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def _p(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def coo_graph(n: int, seed: int):
    """The synthetic E x C graph as the reference's input: the rate map in COO form (src, dst, val) + currency ids."""
    from floydwarshall_b200 import graphs
    E, C = n // CCY, CCY
    blocks = graphs.exchange_blocks(E, C, seed)
    ei, ai, bi = np.nonzero(blocks)
    src = (ei * C + ai).astype(np.int32)
    dst = (ei * C + bi).astype(np.int32)
    val = np.ascontiguousarray(blocks[ei, ai, bi], dtype=np.float64)
    ccy = (np.arange(n) % C).astype(np.int32)
    return ccy, src, dst, val


def time_device(fn, restore, reps=5, warm=2):
    """Best CUDA-event time (ms) of fn() on torch's current stream; restore() runs untimed before each call."""
    import torch
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    best = 1e30
    for it in range(warm + reps):
        restore()
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            best = min(best, e0.elapsed_time(e1))
    return best


def side_configs(ctx, dev, peak_tflops):
    """BASELINE.json's other single-GPU configurations (C2, C3 at its full batch, north-star N=8192): device time
    (CUDA events, best of 5) and bit-exact parity against the CPU oracle, computed in this run."""
    import torch
    from floydwarshall_b200 import dense, graphs
    from oracle import fw_oracle as O        # CHECKER only: the timed calls above it never touch the oracle
    out = {}

    def bits_equal(t, ref):
        return bool(np.array_equal(t.cpu().numpy().view(np.uint64), np.ascontiguousarray(ref).view(np.uint64)))

    # C2: N = 1024 (64 exchanges x 16 currencies)
    rate, nxt = graphs.exchange_graph(64, 16, seed=1235)
    r0, x0 = torch.from_numpy(rate).to(dev), torch.from_numpy(nxt).to(dev)
    r, x = torch.empty_like(r0), torch.empty_like(x0)
    ms = time_device(lambda: dense.solve_device(ctx, r, x), lambda: (r.copy_(r0), x.copy_(x0)))
    ref = O.solve_dense(rate, nxt, threads=0)
    out["C2_n1024"] = {"device_ms": ms, "relax_per_s": 1024.0 ** 3 / (ms * 1e-3), "launches": ctx.last_launches,
                       "frac_of_fp64_peak": 2 * 1024.0 ** 3 / (ms * 1e-3) / 1e12 / peak_tflops,
                       "parity": "bit-exact vs oracle (rates + next, full matrix)"
                       if bits_equal(r, ref.rate) and np.array_equal(x.cpu().numpy(), ref.next) else "MISMATCH"}
    # C3: 4096 snapshot graphs of N = 128 (FSM replay), one CTA per graph
    T = 4096
    rate, nxt = graphs.fsm_replay_batch(8, 16, T, seed=1236)
    r0, x0 = torch.from_numpy(rate).to(dev), torch.from_numpy(nxt).to(dev)
    r, x = torch.empty_like(r0), torch.empty_like(x0)
    ms = time_device(lambda: dense.solve_batched_device(ctx, r, x), lambda: (r.copy_(r0), x.copy_(x0)))
    ref = O.solve_batched(rate, nxt, threads=0)
    out["C3_4096x128"] = {"device_ms": ms, "relax_per_s": T * 128.0 ** 3 / (ms * 1e-3), "launches": ctx.last_launches,
                          "frac_of_fp64_peak": 2 * T * 128.0 ** 3 / (ms * 1e-3) / 1e12 / peak_tflops,
                          "parity": "bit-exact vs oracle (rates + next, all 4096 graphs)"
                          if bits_equal(r, ref.rate) and np.array_equal(x.cpu().numpy(), ref.next) else "MISMATCH"}
    # north-star N = 8192: parity by oracle row replay (the full loop would take ~1 min of host time)
    n = 8192
    rate, nxt = graphs.exchange_graph(n // 16, 16, seed=1303)
    r0, x0 = torch.from_numpy(rate).to(dev), torch.from_numpy(nxt).to(dev)
    r, x = torch.empty_like(r0), torch.empty_like(x0)
    ms = time_device(lambda: dense.solve_device(ctx, r, x), lambda: (r.copy_(r0), x.copy_(x0)), reps=3, warm=1)
    launches = ctx.last_launches
    sink = torch.empty((n, n), dtype=torch.float64, device=dev)
    L = _lib_load()
    L.fw_ctx_set_row_snapshot_sink(ctx.handle, _p(sink), n)
    r.copy_(r0); x.copy_(x0)
    dense.solve_device(ctx, r, x)
    torch.cuda.synchronize()
    L.fw_ctx_set_row_snapshot_sink(ctx.handle, None, 0)
    rows = np.unique(np.random.default_rng(5).integers(0, n, 64)).astype(np.int32)
    rp = O.replay_rows(rows, sink.cpu().numpy(), rate[rows], nxt[rows], threads=0)
    idx = torch.from_numpy(rows.astype(np.int64)).to(dev)
    ok = bits_equal(r[idx], rp.rate) and np.array_equal(x[idx].cpu().numpy(), rp.next)
    out["N8192"] = {"device_ms": ms, "relax_per_s": float(n) ** 3 / (ms * 1e-3), "launches": launches,
                    "frac_of_fp64_peak": 2 * float(n) ** 3 / (ms * 1e-3) / 1e12 / peak_tflops,
                    "parity": f"bit-exact vs oracle row replay ({len(rows)} sampled rows: rates + next; full-matrix "
                              "comparison in tests/test_gpu_parity_large.py)" if ok else "MISMATCH"}
    return out


def _lib_load():
    from floydwarshall_b200 import _lib
    return _lib.load()


def run_ours(args):
    import torch
    from floydwarshall_b200 import _lib, dense

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 or args.gpus > 1:
        return run_multi(args)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = workload_n(1, args.order)
    peak_tflops, peak_src, peak_raw = measured_fp64_peak()

    ctx = _lib.Context(local)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    r0, x0 = device_graph(n, SEED, dev)
    r = torch.empty_like(r0)
    x = torch.empty_like(x0)

    def step():
        r.copy_(r0)            # in-place solve: restore the inputs (device copy, inside the timed region)
        x.copy_(x0)
        dense.solve_device(ctx, r, x)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = ctx.last_launches
    sampler = ClockSampler(local)
    sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    ms_per_step = ms_total / args.steps
    value = float(n) ** 3 / (ms_per_step * 1e-3)
    sum_r = int(r.view(torch.int64).sum().item())
    sum_x = int(x.to(torch.int64).sum().item())

    # ---- roofline of the dominant kernel.  Two passes with per-launch CUDA events on the launching stream:
    # (1) the shipped schedule (pivot phases of the next group on a side stream: spans overlap, their sum can
    #     exceed the step), (2) a SERIALISED pass (FW_OVERLAP=0: one stream, every span is a kernel duration and
    #     their sum is below that pass's step time).  `achieved` comes from (2).
    ctx.set_profiling(True)
    step()
    phase_ms, phase_cnt = ctx.phase_ms()
    ctx.set_profiling(False)
    os.environ["FW_OVERLAP"] = "0"
    cs = _lib.Context(local)
    del os.environ["FW_OVERLAP"]
    cs.set_stream(torch.cuda.current_stream().cuda_stream)
    cs.set_profiling(True)
    r.copy_(r0); x.copy_(x0)
    torch.cuda.synchronize()
    e0.record(); dense.solve_device(cs, r, x); e1.record()
    torch.cuda.synchronize()
    ser_step_ms = e0.elapsed_time(e1)
    ser_ms, ser_cnt = cs.phase_ms()
    cs.close()
    npad = (n + 127) // 128 * 128
    # every entry outside a k-block's own strips takes that block's 128 steps in fw_bulk_kernel, however
    # the launches are arranged (groups, strips first, ...): relaxations per SOLVE done by that kernel
    bulk_relax_total = (npad // 128) * float(npad - 128) ** 2 * 128
    bulk_relax = bulk_relax_total / max(ser_cnt[3], 1)              # average per launch
    achieved = 2.0 * bulk_relax_total / (ser_ms[3] * 1e-3) / 1e12 if ser_cnt[3] else None
    nblk = npad // 128
    group = int(os.environ.get("FW_FUSE_GROUP", "0")) or (8 if nblk >= 128 else (4 if nblk >= 48 else 1))
    traffic = traffic_note = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get("fw_bulk_kernel_dram_bytes_per_launch")
            traffic_note = tj.get("_schedule_note")
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {
        "bound": "fp64", "kernel": "fw_bulk_kernel", "achieved": achieved, "peak": peak_tflops,
        "unit": "TFLOP/s", "frac": (achieved / peak_tflops) if achieved else None, "traffic": traffic,
        "traffic_note": traffic_note, "peak_source": peak_src,
        "avg_launch_ms": ser_ms[3] / max(ser_cnt[3], 1), "launches_per_step": ser_cnt[3],
        "algorithmic_flops_per_launch": 2.0 * bulk_relax,
        "share_of_step": ser_ms[3] / ser_step_ms,
        "serialized_pass": {"step_ms": ser_step_ms, "sum_kernel_ms": sum(ser_ms),
                            "phase_ms": {"tile": ser_ms[0], "col_panel": ser_ms[1], "row_panel": ser_ms[2], "bulk": ser_ms[3]}},
        "overlapped_pass": {"step_ms": ms_per_step, "sum_span_ms": sum(phase_ms),
                            "phase_ms": {"tile": phase_ms[0], "col_panel": phase_ms[1], "row_panel": phase_ms[2], "bulk": phase_ms[3]},
                            "bulk_tflops_from_spans": 2.0 * bulk_relax_total / (phase_ms[3] * 1e-3) / 1e12 if phase_cnt[3] else None},
        "whole_solve_tflops": 2.0 * value / 1e12, "whole_solve_frac": 2.0 * value / 1e12 / peak_tflops,
        "note": "achieved = algorithmic flops of fw_bulk_kernel (2 per relaxation) / its summed launch durations in the "
                "serialised pass; share_of_step = those durations / that pass's step time",
        "hbm_side": {"k_blocks_per_fused_launch": group,
                     "algorithmic_bytes_per_solve": (npad // (128 * group)) * float(npad - 128) ** 2 * 8,
                     "note": "bulk reads every entry once per GROUP of k-blocks (8 B) and writes only replaced entries"},
    }

    # ---- e2e: floydWarshall as the reference's caller sees it -- the rate map goes in (COO, host
    # arrays), the dense matrix comes back (pinned host buffers): H2D + buildMatrix + validation +
    # solve + D2H all inside the timed call fw_solve_edges ----
    e2e = e2e_paths = tables_on = None
    rh = xh = None
    if not args.skip_e2e:
        import ctypes
        ccy, src, dst, val = coo_graph(n, SEED)
        rh = torch.empty((n, n), dtype=torch.float64, pin_memory=True)
        xh = torch.empty((n, n), dtype=torch.int32, pin_memory=True)
        L = _lib.load()
        vp = lambda arr: ctypes.c_void_p(arr.ctypes.data)
        ctx.set_stream(None)
        ts = []
        for it in range(1 + min(args.steps, 2)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _lib.check(L.fw_solve_edges(ctx.handle, n, vp(ccy), len(src), vp(src), vp(dst), vp(val),
                                        _p(rh), _p(xh), None, None, None, None))
            t1 = time.perf_counter()
            if it > 0:
                ts.append(t1 - t0)
        e2e_s = float(np.mean(ts))
        # same answer as the resident solve timed above (bit patterns summed mod 2^64)
        same = sum_r == int(rh.view(torch.int64).sum().item()) and sum_x == int(xh.to(torch.int64).sum().item())
        e2e = {"value": float(n) ** 3 / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(ccy.nbytes + src.nbytes + dst.nbytes + val.nbytes),
               "d2h_bytes_per_step": n * n * 12, "ms_per_step": e2e_s * 1e3,
               "api": "fw_solve_edges (C ABI: rate map in COO form in, dense rate/next out, pinned host buffers)",
               "matches_resident_solve": bool(same)}
        if not args.skip_paths:
            # the same call with the exact-path tables on: init_next / mid / csT / rs come back too (28 B per entry)
            tabs = [torch.empty((n, n), dtype=torch.int32, pin_memory=True) for _ in range(4)]
            ts = []
            for it in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                _lib.check(L.fw_solve_edges(ctx.handle, n, vp(ccy), len(src), vp(src), vp(dst), vp(val),
                                            _p(rh), _p(xh), _p(tabs[0]), _p(tabs[1]), _p(tabs[2]), _p(tabs[3])))
                ts.append(time.perf_counter() - t0)
            same_p = sum_r == int(rh.view(torch.int64).sum().item()) and sum_x == int(xh.to(torch.int64).sum().item())
            e2e_paths = {"value": float(n) ** 3 / ts[-1], "unit": UNIT, "ms_per_step": ts[-1] * 1e3,
                         "h2d_bytes_per_step": e2e["h2d_bytes_per_step"], "d2h_bytes_per_step": n * n * 28,
                         "api": "fw_solve_edges with init_next + mid/csT/rs outputs (what an exact `_path` needs)",
                         "matches_resident_solve": bool(same_p)}
            del tabs
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    if not args.skip_paths:
        # device-only cost of recording mid / csT / rs (exact `_path`, Algorithms.hs:55) at the same size
        tm = [torch.empty_like(x0) for _ in range(3)]
        r.copy_(r0); x.copy_(x0)
        torch.cuda.synchronize()
        e0.record(); dense.solve_device(ctx, r, x, tm[0], tm[1], tm[2]); e1.record()
        torch.cuda.synchronize()
        t_on = e0.elapsed_time(e1)
        same_t = sum_r == int(r.view(torch.int64).sum().item()) and sum_x == int(x.to(torch.int64).sum().item())
        tables_on = {"ms_per_solve": t_on, "relax_per_s": float(n) ** 3 / (t_on * 1e-3),
                     "vs_tables_off": t_on / ms_per_step, "same_rates_and_next": bool(same_t)}
        del tm

    # ---- cpu baseline: bounded sample on the host cores ----
    cpu = None
    if not args.skip_cpu:
        from oracle import fw_oracle as O
        if rh is None:
            rh = torch.empty((n, n), dtype=torch.float64)
            xh = torch.empty((n, n), dtype=torch.int32)
        rh.copy_(r0); xh.copy_(x0)
        torch.cuda.synchronize()
        threads = host_threads()
        ks = max(2, int(16 * (32768 / n) ** 2))
        O.run_ksteps(rh.numpy(), xh.numpy(), 0, 1, threads)      # warm the pages / threads
        t0 = time.perf_counter()
        O.run_ksteps(rh.numpy(), xh.numpy(), 1, 1 + ks, threads)
        t1 = time.perf_counter()
        cpu = {"value": ks * float(n) * n / (t1 - t0), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"k-steps 1..{ks} of the same N={n} matrix ({ks * n * n:.3e} relaxations, "
                         f"{t1 - t0:.1f} s), oracle/fw_oracle.c OpenMP over i"}
    del rh, xh, r, x, r0, x0
    torch.cuda.empty_cache()

    configs = None
    if not args.skip_configs:
        configs = side_configs(ctx, dev, peak_tflops)

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "n": n, "seed": SEED, "k_block": 128,
                   "l2": "inputs (12 GiB at N=32768) are far larger than the 126 MB L2; no flush needed",
                   "timed_region": "device copy of the pristine inputs + fw_solve_device (validation included)",
                   "check": "tests/test_gpu_parity_large.py: oracle row replay, independent schedule and 128-step oracle "
                            "windows at this N; " + ("side configs checked in this run" if configs else "side configs skipped")},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline, "cpu_baseline": cpu, "e2e_paths": e2e_paths, "tables_on": tables_on,
        "configs": configs, "fp64_peak_probe": peak_raw,
    }))
    ctx.close()


def preflight_check(ms, rank, world, dev):
    """Sharded solve against the CPU oracle with the headline run's group policy (groups of 8, cyclic rows), small
    enough for the oracle: N = max(4096, 1024 * world).  Every rank compares the rows it holds, bit for bit."""
    import torch
    import torch.distributed as dist
    from floydwarshall_b200 import graphs
    from oracle import fw_oracle as O        # CHECKER only: untimed pre-flight comparison
    n = max(4096, 1024 * world)
    os.environ["FW_MULTI_GROUP"] = "8"
    try:
        ccy, src, dst, val = coo_graph(n, SEED + 7)
        ms.sync(n, ccy, src, dst, val, paths=False)
    finally:
        del os.environ["FW_MULTI_GROUP"]
    info = ms.shards()[0]
    ref_r = torch.empty((n, n), dtype=torch.float64, device=dev)
    ref_x = torch.empty((n, n), dtype=torch.int32, device=dev)
    if rank == 0:
        rate, nxt = host_graph(n, SEED + 7)
        ref = O.solve_dense(rate, nxt, threads=host_threads())
        ref_r.copy_(torch.from_numpy(ref.rate)); ref_x.copy_(torch.from_numpy(ref.next))
    if world > 1:
        dist.broadcast(ref_r, src=0)
        dist.broadcast(ref_x, src=0)
    rh = torch.empty((info.rows, info.n_padded), dtype=torch.float64)
    xh = torch.empty((info.rows, info.n_padded), dtype=torch.int32)
    ms.download_local(0, rh.data_ptr(), xh.data_ptr())
    from floydwarshall_b200.sharded import global_rows
    g = torch.from_numpy(global_rows(info))
    keep = g < n
    gr = g[keep].to(dev)
    ok = torch.equal(rh[keep][:, :n].to(dev).view(torch.int64), ref_r[gr].view(torch.int64)) and \
        torch.equal(xh[keep][:, :n].to(dev), ref_x[gr])
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) != 1:
        raise SystemExit("sharded result differs from the CPU oracle")
    return (f"bit-exact vs CPU oracle at N={n} on {world} ranks (rates + next, every row), k-blocks in groups of "
            f"{info.group}, cyclic row blocks of {info.cyclic_rows}")


def run_multi(args):
    """bench.py --gpus N (N > 1): config C5, N=65536 row-sharded, strong scaling.  One rank per GPU under torchrun
    (fw_multi_create_rank: the C++ schedule, NCCL broadcast of the pivot-row panels); with --single-process rank 0
    alone drives all N GPUs through fw_multi_create (copy-engine panel transport), the reference's in-process call."""
    import torch
    import torch.distributed as dist
    from floydwarshall_b200 import sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    single = args.single_process or world == 1
    if single and rank != 0:            # one process drives every GPU: the other torchrun ranks exit without work
        return
    if single:
        world, local = 1, 0             # no process group at all in this mode
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    ngpu = args.gpus
    n = workload_n(ngpu, args.order)

    def barrier():
        if world > 1:
            dist.barrier()

    peak_tflops, peak_src, peak_raw = measured_fp64_peak() if rank == 0 else (None, None, None)
    ms = sharded.MultiSolver(devices=list(range(ngpu))) if single else sharded.MultiSolver.for_torchrun(local)
    eff_world = ngpu if single else world

    # the full-size buffers first: the small pre-flight problem then reuses them (nothing that a peer has mapped is
    # reallocated between the two problems)
    ms.alloc(n, paths=False)
    check = None
    if not args.skip_check:
        if single:
            from floydwarshall_b200 import graphs
            from oracle import fw_oracle as O
            nc = max(4096, 1024 * eff_world)
            os.environ["FW_MULTI_GROUP"] = "8"
            ccy, src, dst, val = coo_graph(nc, SEED + 7)
            res = ms.solve_edges(nc, ccy, src, dst, val)
            del os.environ["FW_MULTI_GROUP"]
            rate, nxt = host_graph(nc, SEED + 7)
            ref = O.solve_dense(rate, nxt, threads=host_threads())
            if not (np.array_equal(res.rate.view(np.uint64), ref.rate.view(np.uint64)) and np.array_equal(res.next, ref.next)):
                raise SystemExit("sharded result differs from the CPU oracle")
            check = f"bit-exact vs CPU oracle at N={nc} on {eff_world} GPUs of one process (rates + next), groups of 8"
        else:
            check = preflight_check(ms, rank, world, dev)

    ccy, src, dst, val = coo_graph(n, SEED + 1)
    ms.sync(n, ccy, src, dst, val, paths=False)             # uploads the COO; counts as the first warm-up step
    for _ in range(max(args.warmup - 1, 0)):
        ms.resolve()
    sampler = ClockSampler(local)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    dev_ms, launches = [], 0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ms.resolve()                                        # buildMatrix from the resident COO + validation + solve
        t, ln = ms.last_solve()
        dev_ms.append(t); launches += ln
    torch.cuda.synchronize()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()
    tt = torch.tensor([float(np.mean(dev_ms)), wall_ms, float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tt.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_per_step, wall_ms, launches = float(mx[0]), float(mx[1]), int(sm[2])
    else:
        ms_per_step = float(tt[0])
    value = float(n) ** 3 / (ms_per_step * 1e-3)
    info = ms.shards()[0]

    # per-phase profile of one step on this rank's (first) shard: bulk kernel roofline, per GPU
    ms.set_profiling(True)
    ms.resolve()
    pms, pcnt = ms.phase_ms()
    ms.set_profiling(False)

    # e2e: the rate map goes in from the host (COO), every shard's rows come back into pinned host buffers
    e2e = e2e_resident = None
    if not args.skip_e2e:
        nloc = len(ms.shards())
        bufs = [(torch.empty((info.rows, info.n_padded), dtype=torch.float64, pin_memory=True),
                 torch.empty((info.rows, info.n_padded), dtype=torch.int32, pin_memory=True)) for _ in range(nloc)]
        ts = []
        for it in range(2):
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            ms.sync(n, ccy, src, dst, val, paths=False)
            ms.download_locals([b[0].data_ptr() for b in bufs], [b[1].data_ptr() for b in bufs])
            barrier()
            ts.append(time.perf_counter() - t0)
        te = torch.tensor([ts[-1]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        coo_bytes = int(ccy.nbytes + src.nbytes + dst.nbytes + val.nbytes)
        e2e = {"value": float(n) ** 3 / float(te.item()), "unit": UNIT,
               "h2d_bytes_per_step": coo_bytes * eff_world, "d2h_bytes_per_step": info.n_padded * info.n_padded * 12,
               "ms_per_step": float(te.item()) * 1e3,
               "api": "fw_multi_sync (rate map in COO form from host arrays) + fw_multi_download_locals (every shard's "
                      "rows into pinned host buffers)"}
        del bufs
        if single:
            # the REPL flow (ProcessRequests.hs:78-85): syncMatrix, then `optimum` reads ONE entry + its path; the
            # optimised matrix stays sharded in HBM
            t0 = time.perf_counter()
            ms.sync(n, ccy, src, dst, val, paths=True)
            rate_q, path_q = ms.optimum(3, n - 5)
            t1 = time.perf_counter()
            tq = []
            for q in range(20):
                a = time.perf_counter(); ms.optimum((q * 7919) % n, (q * 104729 + 1) % n); tq.append(time.perf_counter() - a)
            e2e_resident = {"value": float(n) ** 3 / (t1 - t0), "unit": UNIT, "ms_per_step": (t1 - t0) * 1e3,
                            "h2d_bytes_per_step": coo_bytes * eff_world, "d2h_bytes_per_step": 16 + 4 * len(path_q),
                            "optimum_query_us": float(np.median(tq)) * 1e6, "path_tables": True,
                            "api": "fw_multi_sync(want_paths=1) + fw_multi_optimum: syncMatrix + optimum, matrix resident"}

    cpu = None
    if rank == 0 and not args.skip_cpu:
        from oracle import fw_oracle as O
        nh = 16384
        rate, nxt = host_graph(nh, SEED + 1)
        threads = host_threads()
        O.run_ksteps(rate, nxt, 0, 1, threads)
        ks = 32
        t0 = time.perf_counter()
        O.run_ksteps(rate, nxt, 1, 1 + ks, threads)
        t1 = time.perf_counter()
        cpu = {"value": ks * float(nh) * nh / (t1 - t0), "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"k-steps 1..{ks} of the N={nh} graph of the same family ({ks * nh * nh:.3e} relaxations, "
                         f"{t1 - t0:.1f} s; relaxations/s is per-relaxation), oracle/fw_oracle.c OpenMP over i"}
    barrier()

    if rank == 0:
        rows = info.rows
        G = info.group
        ngrp = info.n_padded // (G * 128)
        bulk_relax_total = float(rows) * (info.n_padded - 128) * 128 * (info.n_padded // 128)   # this GPU's share
        ach = (2.0 * bulk_relax_total / (pms[3] * 1e-3) / 1e12) if pcnt[3] else None
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": eff_world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(n), "n": n, "seed": SEED + 1, "k_block": 128,
                       "processes": "one process drives all GPUs (fw_multi_create)" if single
                       else "one process per GPU (fw_multi_create_rank; NCCL bootstraps the ranks)",
                       "panel_transport": ms.transport(),
                       "sharding": f"rows in cyclic blocks of {info.cyclic_rows} over {eff_world} ranks ({rows} rows per rank); "
                                   f"per k-block the 128 x {info.n_padded} fp64 pivot-row snapshot panel "
                                   f"({128 * info.n_padded * 8 / 2**20:.0f} MiB) goes from its owner to every rank",
                       "k_blocks_per_bulk_launch": G, "groups": ngrp,
                       "timed_region": "buildMatrix from the device-resident rate map (COO) + validation + solve; CUDA "
                                       "events inside the library around every step, max over ranks",
                       "wall_ms_per_step": wall_ms,
                       "l2": "per-rank inputs are far larger than the 126 MB L2; no flush needed",
                       "check": check},
            "clocks": clocks, "e2e": e2e, "e2e_resident": e2e_resident, "gpu_launches": int(launches),
            "roofline": {"bound": "fp64", "kernel": "fw_bulk_kernel", "achieved": ach, "peak": peak_tflops,
                         "unit": "TFLOP/s", "frac": (ach / peak_tflops) if ach else None, "traffic": None,
                         "peak_source": peak_src, "avg_launch_ms": pms[3] / max(pcnt[3], 1), "launches_per_step": pcnt[3],
                         "phase_ms": {"tile": pms[0], "col_panel": pms[1], "row_panel": pms[2], "bulk": pms[3]},
                         "whole_solve_tflops_per_gpu": 2.0 * value / 1e12 / eff_world,
                         "note": "rank 0's launches (per-GPU figure); spans of the two lanes overlap"},
            "cpu_baseline": cpu, "fp64_peak_probe": peak_raw,
        }))
    ms.close()
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: native libraries write banners to fd 1 directly (NCCL prints
    # "NCCL version ..." there), so fd 1 is pointed at stderr and Python's stdout keeps the real one.
    sys.stdout.flush()
    real_out = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_out, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
