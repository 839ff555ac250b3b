"""Solve, then solve the RESULT again: the second solve fires nothing (fixed point), so the difference
is the cost of leaving the fast path (exact replays + write-back), and with a -DFW_BULK_STATS build the
second solve's slow fraction is the filter's false-candidate rate (exact a*b > o but RN(a*b) <= o)."""
import ctypes
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from floydwarshall_b200 import _lib, dense, graphs

sizes = [int(a) for a in sys.argv[1:]] or [8192]
ctx = _lib.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
L = _lib.load()
stats = hasattr(L, "fw_debug_bulk_stats")
out = (ctypes.c_ulonglong * 4)()
for n in sizes:
    rate, nxt = graphs.exchange_graph(n // 16, 16, seed=1234)
    r = torch.from_numpy(rate).cuda(); x = torch.from_numpy(nxt).cuda()
    res = {"n": n}
    for tag in ("first", "second", "third"):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        if stats:
            L.fw_debug_bulk_stats(out, 1)
        before = r.clone()
        e0.record(); dense.solve_device(ctx, r, x); e1.record(); torch.cuda.synchronize()
        res[tag + "_ms"] = round(e0.elapsed_time(e1), 3)
        res[tag + "_changed_entries"] = int((before.view(torch.int64) != r.view(torch.int64)).sum().item())
        if stats:
            L.fw_debug_bulk_stats(out, 1)
            res[tag + "_slow_frac"] = int(out[1]) / max(int(out[0]), 1)
            res[tag + "_fires"] = int(out[3])
        del before
    print(json.dumps(res))
