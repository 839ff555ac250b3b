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
    ap.add_argument("--check", action="store_true",
                    help="multi-GPU: verify the sharded result against a single-GPU solve (small --n only)")
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    from floydwarshall_b200 import _lib, dense

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 or args.gpus > 1:
        from floydwarshall_b200 import sharded
        return sharded.bench_main(args, METRIC, UNIT, SEED, workload_n, workload_name, ClockSampler,
                                  measured_fp64_peak)
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

    # ---- roofline of the dominant kernel: per-launch CUDA events on the launching stream ----
    ctx.set_profiling(True)
    step()
    phase_ms, phase_cnt = ctx.phase_ms()
    ctx.set_profiling(False)
    npad = (n + 127) // 128 * 128
    # every entry outside a k-block's own strips takes that block's 128 steps in fw_bulk_kernel, however
    # the launches are arranged (pairs, strips first, ...): relaxations per SOLVE done by that kernel
    bulk_relax_total = (npad // 128) * float(npad - 128) ** 2 * 128
    bulk_relax = bulk_relax_total / max(phase_cnt[3], 1)            # average per launch
    if phase_cnt[3] > 0:
        bulk_ms = phase_ms[3] / phase_cnt[3]
        achieved = 2.0 * bulk_relax_total / (phase_ms[3] * 1e-3) / 1e12   # algorithmic FLOPs: 1 mul + 1 compare
    else:
        bulk_ms, achieved = None, None
    # k-blocks per fused bulk launch: the library's size-based policy (fwgpu.cu, solve_blocked) unless forced
    nblk = npad // 128
    group = int(os.environ.get("FW_FUSE_GROUP", "0")) or (8 if nblk >= 128 else (4 if nblk >= 48 else 1))
    traffic = None
    traffic_note = None
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
        "traffic_note": traffic_note,
        "peak_source": peak_src, "avg_launch_ms": bulk_ms, "launches_per_step": phase_cnt[3],
        "algorithmic_flops_per_launch": 2.0 * bulk_relax,
        "share_of_step": (phase_ms[3] / sum(phase_ms)) if sum(phase_ms) > 0 else None,
        "whole_solve_tflops": 2.0 * value / 1e12,
        "note": "per-launch CUDA-event spans on the launching streams; the strip launches of the NEXT group run on "
                "the side stream while the main stream's bulk launch is busy, so the spans overlap and their sum "
                "can exceed the step time (achieved is the conservative figure, whole_solve_tflops = 2 N^3 / t)",
        "phase_ms": {"tile": phase_ms[0], "col_panel": phase_ms[1], "row_panel": phase_ms[2], "bulk": phase_ms[3]},
        "hbm_side": {"k_blocks_per_fused_launch": group,
                     "algorithmic_bytes_per_solve": (npad // (128 * group)) * float(npad - 128) ** 2 * 8,
                     "note": "bulk reads every entry once per GROUP of k-blocks (8 B) and writes only replaced "
                             "entries; launches of different sizes (strips / rest) are averaged"},
    }

    # ---- e2e: floydWarshall as the reference's caller sees it -- the rate map goes in (COO, host
    # arrays), the dense matrix comes back (pinned host buffers): H2D + buildMatrix + validation +
    # solve + D2H all inside the timed call fw_solve_edges ----
    e2e = None
    rh = xh = None
    if not args.skip_e2e:
        import ctypes
        from floydwarshall_b200 import graphs
        E, C = n // CCY, CCY
        blocks = graphs.exchange_blocks(E, C, SEED)
        ei, ai, bi = np.nonzero(blocks)
        src = (ei * C + ai).astype(np.int32)
        dst = (ei * C + bi).astype(np.int32)
        val = np.ascontiguousarray(blocks[ei, ai, bi], dtype=np.float64)
        ccy = (np.arange(n) % C).astype(np.int32)
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
                                        ctypes.c_void_p(rh.data_ptr()), ctypes.c_void_p(xh.data_ptr()),
                                        None, None, None, None))
            t1 = time.perf_counter()
            if it > 0:
                ts.append(t1 - t0)
        e2e_s = float(np.mean(ts))
        # same answer as the resident solve timed above (bit patterns summed mod 2^64)
        same = int(r.view(torch.int64).sum().item()) == int(rh.view(torch.int64).sum().item()) and \
            int(x.to(torch.int64).sum().item()) == int(xh.to(torch.int64).sum().item())
        e2e = {"value": float(n) ** 3 / e2e_s, "unit": UNIT,
               "h2d_bytes_per_step": int(ccy.nbytes + src.nbytes + dst.nbytes + val.nbytes),
               "d2h_bytes_per_step": n * n * 12, "ms_per_step": e2e_s * 1e3,
               "api": "fw_solve_edges (C ABI: rate map in COO form in, dense rate/next out, pinned host buffers)",
               "matches_resident_solve": bool(same)}
        ctx.set_stream(torch.cuda.current_stream().cuda_stream)

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

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n), "n": n, "seed": SEED, "k_block": 128,
                   "l2": "inputs (12 GiB at N=32768) are far larger than the 126 MB L2; no flush needed",
                   "timed_region": "device copy of the pristine inputs + fw_solve_device (validation included)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
        "roofline": roofline, "cpu_baseline": cpu, "fp64_peak_probe": peak_raw,
    }))
    ctx.close()


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
