"""Time device-resident solves with the per-phase breakdown (CUDA events)."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from floydwarshall_b200 import _lib, dense, graphs

sizes = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096, 8192]
ctx = _lib.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
for n in sizes:
    E = n // 16
    rate, nxt = graphs.exchange_graph(E, 16, seed=1234)
    r0 = torch.from_numpy(rate).cuda(); x0 = torch.from_numpy(nxt).cuda()
    r = torch.empty_like(r0); x = torch.empty_like(x0)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e30
    for it in range(3):
        r.copy_(r0); x.copy_(x0)
        ctx.set_profiling(it == 2)
        e0.record(); dense.solve_device(ctx, r, x); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    ms, cnt = ctx.phase_ms()
    spans = ctx.phase_spans(3)
    ctx.set_profiling(False)
    if spans:
        q = max(1, len(spans) // 8)
        print("  bulk ms per launch (every %d-th): " % q + " ".join(f"{v:.3f}" for v in spans[::q]))
    print(json.dumps({"n": n, "ms": round(best, 3), "relax_per_s": f"{n**3 / (best * 1e-3):.4e}",
                      "phase_ms": [round(m, 3) for m in ms], "phase_launches": cnt,
                      "bulk_relax_per_s": f"{(n // 128) * (n - 128) ** 2 * 128 / (ms[3] * 1e-3 + 1e-12):.4e}"}))
