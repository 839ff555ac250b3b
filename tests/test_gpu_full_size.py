"""Size-independent properties at BASELINE.json's full single-GPU size (N=32768), where the CPU
oracle would need over an hour:

  * monotone: no rate decreases, the diagonal is untouched;
  * closure: a second solve on the solved matrix changes no rate by more than 1e-12 relative
    (the max-times triangle inequality holds up to rounding);
  * next-hop walks: for sampled (i, j), following `next` from i reaches j and the product of the
    INITIAL rates along the walk equals rate[i][j] to 1e-12 relative (consistent graphs are
    arbitrage-free, so walks terminate);
  * sub-problem agreement: the leading 1024 x 1024 block of the N=32768 generator is the N=1024
    generator's graph (same seed), and solving it alone can only give rates <= the full solve's.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 32768
CCY = 16
SEED = 1237


def test_full_size_properties():
    import torch
    from bench import device_graph
    from floydwarshall_b200 import _lib, dense

    dev = torch.device("cuda", 0)
    free, _total = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~40 GB of free HBM")
    ctx = _lib.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    r0, x0 = device_graph(N, SEED, dev)
    r = r0.clone()
    x = x0.clone()
    dense.solve_device(ctx, r, x)
    torch.cuda.synchronize()
    # monotone + diagonal untouched
    assert bool((r >= r0).all())
    assert bool((r.diagonal() == r0.diagonal()).all()) and bool((x.diagonal() == -1).all())
    # closure under one more full solve
    r2 = r.clone()
    x2 = x.clone()
    dense.solve_device(ctx, r2, x2)
    torch.cuda.synchronize()
    rel = ((r2 - r) / r.clamp_min(1e-300)).abs().max().item()
    assert rel <= 1e-12, rel
    del r2, x2
    # next-hop walks on sampled pairs (host side, a few rows fetched on demand)
    rng = np.random.default_rng(0)
    pairs = rng.integers(0, N, size=(300, 2))
    xs = x.cpu().numpy()
    r0h_rows = {}
    for i, j in pairs:
        i, j = int(i), int(j)
        if i == j:
            continue
        want = r[i, j].item()
        if xs[i, j] < 0:
            assert want == 0.0
            continue
        cur, prod, hops = i, 1.0, 0
        while cur != j:
            nx = int(xs[cur, j])
            assert nx >= 0 and hops < 64
            if cur not in r0h_rows:
                r0h_rows[cur] = r0[cur].cpu().numpy()
            prod *= r0h_rows[cur][nx]
            cur = nx
            hops += 1
        assert abs(prod - want) <= 1e-12 * want, (i, j, prod, want)
    # sub-problem: the leading block solved alone never beats the full solve
    n1 = 1024
    r1 = r0[:n1, :n1].contiguous()
    x1 = x0[:n1, :n1].contiguous()
    dense.solve_device(ctx, r1, x1)
    torch.cuda.synchronize()
    assert bool((r1 <= r[:n1, :n1] * (1 + 1e-12)).all())
    ctx.close()
