"""How often does the bulk kernel leave its fast path?  Needs the stats build:
   make -C floydwarshall_b200/csrc VARIANT=_stats EXTRA=-DFW_BULK_STATS
   FWGPU_LIB=floydwarshall_b200/libfwgpu_stats.so python tools/bulk_stats.py 4096 8192"""
import ctypes
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from floydwarshall_b200 import _lib, dense, graphs

sizes = [int(a) for a in sys.argv[1:]] or [4096]
ctx = _lib.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
L = _lib.load()
out = (ctypes.c_ulonglong * 4)()
for n in sizes:
    rate, nxt = graphs.exchange_graph(n // 16, 16, seed=1234)
    r = torch.from_numpy(rate).cuda(); x = torch.from_numpy(nxt).cuda()
    L.fw_debug_bulk_stats(out, 1)
    dense.solve_device(ctx, r, x)
    L.fw_debug_bulk_stats(out, 1)
    ws, slow, rows, fires = [int(v) for v in out]
    print(json.dumps({"n": n, "warp_steps": ws, "slow_warp_steps": slow, "slow_frac": slow / max(ws, 1),
                      "rows_replayed_per_slow_step": rows / max(slow, 1), "fires": fires,
                      "fires_per_entry": fires / (n * n), "fires_per_slow_step": fires / max(slow, 1)}))
