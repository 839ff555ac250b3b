"""One device-resident solve of the synthetic N x N graph (default 32768) and nothing else: the target of the ncu
captures (launch list, --set full of one fused bulk launch).  usage: one_solve.py [N] [paths]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import device_graph, SEED
from floydwarshall_b200 import _lib, dense

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
paths = len(sys.argv) > 2 and sys.argv[2] == "paths"
dev = torch.device("cuda", 0)
ctx = _lib.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
r, x = device_graph(n, SEED, dev)
tabs = [torch.empty_like(x) for _ in range(3)] if paths else [None, None, None]
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); dense.solve_device(ctx, r, x, *tabs); e1.record()
torch.cuda.synchronize()
print(f"n={n} paths={paths} ms={e0.elapsed_time(e1):.1f} launches={ctx.last_launches}")
