"""Device timings (CUDA events, best of 5) of the side configurations: C2 (N=1024), C3 (4096 x N=128 batched FSM
replay), N=4096 and N=8192.  Select an experiment build with FWGPU_LIB=..., knobs with FW_* variables."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from floydwarshall_b200 import _lib, dense, graphs

ctx = _lib.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
out = {"lib": os.path.basename(_lib.LIB_PATH), "knobs": {k: v for k, v in os.environ.items() if k.startswith("FW_")}}


def best(fn, restore, reps=5, warm=2):
    b = 1e30
    for it in range(warm + reps):
        restore(); torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if it >= warm:
            b = min(b, e0.elapsed_time(e1))
    return b


for name, E in (("C2_n1024", 64), ("n4096", 256), ("n8192", 512)):
    rate, nxt = graphs.exchange_graph(E, 16, seed=1235)
    n = E * 16
    r0 = torch.from_numpy(rate).cuda(); x0 = torch.from_numpy(nxt).cuda()
    r = torch.empty_like(r0); x = torch.empty_like(x0)
    ms = best(lambda: dense.solve_device(ctx, r, x), lambda: (r.copy_(r0), x.copy_(x0)), reps=5 if n < 8192 else 3)
    ctx.set_profiling(True); r.copy_(r0); x.copy_(x0); dense.solve_device(ctx, r, x); pm, cnt = ctx.phase_ms(); ctx.set_profiling(False)
    out[name] = {"device_ms": ms, "relax_per_s": float(n) ** 3 / (ms * 1e-3), "phase_ms": pm, "launches": cnt}
    del r0, x0, r, x

T = 4096
rate, nxt = graphs.fsm_replay_batch(8, 16, T, seed=1236)
r0 = torch.from_numpy(rate).cuda(); x0 = torch.from_numpy(nxt).cuda()
r = torch.empty_like(r0); x = torch.empty_like(x0)
ms = best(lambda: dense.solve_batched_device(ctx, r, x), lambda: (r.copy_(r0), x.copy_(x0)))
out["C3_4096x128"] = {"device_ms": ms, "relax_per_s": T * 128.0 ** 3 / (ms * 1e-3)}
m3 = [torch.empty_like(x0) for _ in range(3)]
ms = best(lambda: dense.solve_batched_device(ctx, r, x, m3[0], m3[1], m3[2]), lambda: (r.copy_(r0), x.copy_(x0)))
out["C3_4096x128_paths"] = {"device_ms": ms, "relax_per_s": T * 128.0 ** 3 / (ms * 1e-3)}
print(json.dumps(out))
