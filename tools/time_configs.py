"""Timings of the BASELINE.json parity configs C2 (N=1024) and C3 (4096 x N=128 batched FSM replay):
device-resident (CUDA events) and host-buffer (wall clock through fw_solve / fw_solve_batched)."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from floydwarshall_b200 import _lib, dense, graphs

ctx = _lib.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
out = {}

# C2
rate, nxt = graphs.exchange_graph(64, 16, seed=1235)
r0 = torch.from_numpy(rate).cuda(); x0 = torch.from_numpy(nxt).cuda()
r = torch.empty_like(r0); x = torch.empty_like(x0)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
best = 1e30
for it in range(6):
    r.copy_(r0); x.copy_(x0)
    e0.record(); dense.solve_device(ctx, r, x); e1.record(); torch.cuda.synchronize()
    if it >= 2:
        best = min(best, e0.elapsed_time(e1))
ctx.set_profiling(True); r.copy_(r0); x.copy_(x0); dense.solve_device(ctx, r, x); ms, cnt = ctx.phase_ms(); ctx.set_profiling(False)
t0 = time.perf_counter(); res = dense.solve(rate, nxt, ctx=None); t1 = time.perf_counter()
t0 = time.perf_counter(); res = dense.solve(rate, nxt, ctx=None); t1 = time.perf_counter()
out["C2_n1024"] = {"device_ms": best, "relax_per_s": 1024 ** 3 / (best * 1e-3), "phase_ms": ms, "launches": cnt,
                   "host_api_ms": (t1 - t0) * 1e3, "host_relax_per_s": 1024 ** 3 / (t1 - t0)}

# C3
T = 4096
rate, nxt = graphs.fsm_replay_batch(8, 16, T, seed=1236)
r0 = torch.from_numpy(rate).cuda(); x0 = torch.from_numpy(nxt).cuda()
r = torch.empty_like(r0); x = torch.empty_like(x0)
best = 1e30
for it in range(6):
    r.copy_(r0); x.copy_(x0)
    e0.record(); dense.solve_batched_device(ctx, r, x); e1.record(); torch.cuda.synchronize()
    if it >= 2:
        best = min(best, e0.elapsed_time(e1))
rp = torch.from_numpy(rate).pin_memory(); xp = torch.from_numpy(nxt).pin_memory()
L = _lib.load()
import ctypes
ts = []
for it in range(3):
    rw = rp.clone().pin_memory(); xw = xp.clone().pin_memory()
    t0 = time.perf_counter()
    _lib.check(L.fw_solve_batched(None, T, 128, ctypes.c_void_p(rw.data_ptr()), ctypes.c_void_p(xw.data_ptr()), None, None, None))
    ts.append(time.perf_counter() - t0)
out["C3_4096x128"] = {"device_ms": best, "relax_per_s": T * 128 ** 3 / (best * 1e-3),
                      "host_api_ms": min(ts) * 1e3, "host_relax_per_s": T * 128 ** 3 / min(ts)}
print(json.dumps(out, indent=1))
