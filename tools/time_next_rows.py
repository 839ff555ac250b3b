"""Timings of SURVEY.md 8(f)'s widened rows: (1) exact `_path` materialisation (solve with the
mid/csT/rs side tables + fw_paths expansion), (2)+(3) fw_state_sync (map in COO form in, buildMatrix +
runAlgo on the device, matrix kept in HBM) and the per-query read-out fw_state_optimum."""
import ctypes
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from floydwarshall_b200 import _lib, dense, graphs, paths

out = {}
ctx = _lib.Context(0)
L = _lib.load()
vp = lambda a: ctypes.c_void_p(a.ctypes.data)

# (1) N=4096 with side tables, then 65536 random (src, dst) expansions on the device
n = 4096
rate, nxt = graphs.exchange_graph(n // 16, 16, seed=1234)
t0 = time.perf_counter(); plain = dense.solve(rate, nxt, ctx=ctx); t_plain = time.perf_counter() - t0
t0 = time.perf_counter(); res = dense.solve(rate, nxt, paths=True, ctx=ctx); t_paths = time.perf_counter() - t0
rng = np.random.default_rng(7)
q = rng.integers(0, n, size=(65536, 2)).astype(np.int32)
q = q[q[:, 0] != q[:, 1]]
pairs = [tuple(x) for x in q.tolist()]
t0 = time.perf_counter(); pl = paths.expand(nxt, res.mid, res.csT, res.rs, pairs, ctx=ctx); t_exp = time.perf_counter() - t0
hops = sum(len(p) for p in pl)
out["paths_n4096"] = {"host_solve_ms": t_plain * 1e3, "host_solve_with_side_tables_ms": t_paths * 1e3,
                      "queries": len(pairs), "hops_total": hops, "expand_ms_incl_table_upload": t_exp * 1e3,
                      "paths_per_s": len(pairs) / t_exp}

# (2)+(3) resident state at N=8192: sync from the COO map, then single-pair look-ups
n = 8192
E, C = n // 16, 16
blocks = graphs.exchange_blocks(E, C, 1234)
ei, ai, bi = np.nonzero(blocks)
src = (ei * C + ai).astype(np.int32); dst = (ei * C + bi).astype(np.int32)
val = np.ascontiguousarray(blocks[ei, ai, bi], dtype=np.float64)
ccy = (np.arange(n) % C).astype(np.int32)
h = ctypes.c_void_p()
_lib.check(L.fw_state_create(ctx.handle, ctypes.byref(h)))
ts = []
for it in range(3):
    t0 = time.perf_counter()
    _lib.check(L.fw_state_sync(h, n, vp(ccy), len(src), vp(src), vp(dst), vp(val)))
    ts.append(time.perf_counter() - t0)
rate_o = ctypes.c_double(); plen = ctypes.c_int32(); path = np.empty(4096, dtype=np.int32)
lat = []
for i, j in q[:2000].tolist():
    i %= n; j %= n
    if i == j:
        continue
    t0 = time.perf_counter()
    rc = L.fw_state_optimum(h, i, j, ctypes.byref(rate_o), vp(path), 4096, ctypes.byref(plen))
    lat.append(time.perf_counter() - t0)
    assert rc == 0, rc
L.fw_state_destroy(h)
out["state_n8192"] = {"edges": int(len(src)), "sync_ms_best": min(ts) * 1e3, "sync_relax_per_s": n ** 3 / min(ts),
                      "optimum_queries": len(lat), "optimum_us_median": float(np.median(lat)) * 1e6,
                      "optimum_us_p99": float(np.quantile(lat, 0.99)) * 1e6}
print(json.dumps(out, indent=1))
