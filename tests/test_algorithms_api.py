"""The reference's own Algorithms tests (src/test/AlgorithmsTest.hs) replayed against the
host-side mirror floydwarshall_b200.algorithms (buildMatrix / floydWarshall / optimum)."""
import json
import os

import numpy as np
import pytest

from floydwarshall_b200 import algorithms as A
from floydwarshall_b200.types import RateEntry, Vertex, isolated_entry

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def V(p):
    return Vertex(p[0], p[1])


def mock_map(*names):
    return {(V(GOLD["mock_rates"][k][0]), V(GOLD["mock_rates"][k][1])): GOLD["mock_rates"][k][2] for k in names}


ALL4 = ("gdax_btc_usd", "kraken_btc_usd", "gdax_usd_btc", "kraken_usd_btc")


def rate_matrix_for_test(vertices, rows):
    """TestUtils.hs:11-21 rateMatrixForTest."""
    return [[RateEntry(float(r), vertices[i], [vertices[p] for p in path]) for (r, path) in row]
            for i, row in enumerate(rows)]


# ---------------------------------------------------------------- CPU-only (host logic)
def test_buildMatrix_emptyMatrix():
    assert A.build_matrix({}) == []                      # AlgorithmsTest.hs:45-47


def test_buildMatrix_4x4Matrix():
    vs = [V(p) for p in GOLD["vertices_4x4"]]
    assert A.sorted_vertices(mock_map(*ALL4)) == vs
    expected = rate_matrix_for_test(vs, GOLD["G2_buildMatrix_4x4"]["matrix"])
    assert A.build_matrix(mock_map(*ALL4)) == expected   # AlgorithmsTest.hs:49-60


def test_buildMatrix_same_currency_beats_map_entry():
    """Algorithms.hs:35 is checked before the lookup at :36."""
    a, b = Vertex("X", "BTC"), Vertex("Y", "BTC")
    m = A.build_matrix({(a, b): 7.0, (b, a): 0.1})
    assert m[0][1].best_rate == 1.0 and m[1][0].best_rate == 1.0


def test_optimum_matrixMaybeEmpty():
    """AlgorithmsTest.hs:112-134 with MockData.genRateMatrix, enumerated."""
    sv = [V(p) for p in GOLD["G5_all_ones"]["sample_vertices"]]
    rng = np.random.default_rng(17)
    for _ in range(80):
        k = int(rng.integers(0, len(sv) // 2 + 2))
        verts = sorted(set(sv[int(x)] for x in rng.integers(0, len(sv), size=k)))
        if bool(rng.integers(0, 2)):
            matrix = [[] for _ in verts]
        else:
            matrix = [[isolated_entry(s) if s == d else RateEntry(1.0, s, [d]) for d in verts] for s in verts]
        src, dest = sv[int(rng.integers(0, len(sv)))], sv[int(rng.integers(0, len(sv)))]
        try:
            got, err = A.optimum(src, dest, matrix), None
        except A.AlgoOptimumError as ex:
            got, err = None, ex.msg
        if len(matrix) == 0:
            assert err == f"{src.show()} is not entered before"
        elif any(len(r) == 0 for r in matrix):
            assert err == "The matrix is empty"
        elif src not in verts:
            assert err == f"{src.show()} is not entered before"
        elif dest not in verts:
            assert err == f"{dest.show()} is not entered before"
        elif src == dest:
            assert err == f"There is no exchange between {src.show()} and {dest.show()}"
        else:
            assert got == RateEntry(1.0, src, [dest])


def test_floydWarshall_emptyMatrix_needs_no_gpu():
    assert A.floyd_warshall({}) == []                    # AlgorithmsTest.hs:62-64


def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly (no oracle / CPU fallback)."""
    from floydwarshall_b200 import _lib
    if _lib.load().fw_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(_lib.FwError):
        A.floyd_warshall(mock_map(*ALL4))


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_floydWarshall_4x4Matrix():
    vs = [V(p) for p in GOLD["vertices_4x4"]]
    expected = rate_matrix_for_test(vs, GOLD["G1_floydWarshall_4x4"]["matrix"])
    result = A.floyd_warshall(mock_map(*ALL4))
    assert result == expected                            # AlgorithmsTest.hs:66-77
    assert result.to_lists() == expected


@pytest.mark.gpu
def test_optimum_srcOrDestNotExist():
    m = A.floyd_warshall(mock_map(*ALL4))
    g = GOLD["G4_optimum"]["unknown_vertex"]
    for s, d in g["queries"]:
        with pytest.raises(A.AlgoOptimumError) as ei:
            A.optimum(V(s), V(d), m)
        assert ei.value.msg == g["error"]                # AlgorithmsTest.hs:82-91


@pytest.mark.gpu
def test_optimum_reachability():
    m = A.floyd_warshall(mock_map(*ALL4)).to_lists()
    g = GOLD["G4_optimum"]["reachability"]
    i, j = g["isolate"]
    m[i][j] = isolated_entry(m[i][0].start)              # AlgorithmsTest.hs:99-102
    with pytest.raises(A.AlgoOptimumError) as ei:
        A.optimum(V(g["unreachable"]["query"][0]), V(g["unreachable"]["query"][1]), m)
    assert ei.value.msg == g["unreachable"]["error"]
    for a in g["answers"]:
        e = A.optimum(V(a["query"][0]), V(a["query"][1]), m)
        assert e == RateEntry(a["rate"], V(a["query"][0]), [V(p) for p in a["path"]])


@pytest.mark.gpu
def test_readme_session_answers():
    """README.md:210-246 and ProcessRequestsTest.hs:83-95."""
    g = GOLD["G7_readme_session"]
    m2 = A.floyd_warshall(mock_map("kraken_btc_usd", "kraken_usd_btc"))
    for a in g["kraken_only"]:
        e = A.optimum(V(a["query"][0]), V(a["query"][1]), m2)
        assert e.best_rate == a["rate"] and e.path == [V(p) for p in a["path"]]
    m4 = A.floyd_warshall(mock_map(*ALL4))
    for a in g["four_vertices"]:
        e = A.optimum(V(a["query"][0]), V(a["query"][1]), m4)
        assert e.best_rate == a["rate"] and e.path == [V(p) for p in a["path"]]
    e = A.optimum(V(GOLD["G6_end_to_end"]["query"][0]), V(GOLD["G6_end_to_end"]["query"][1]), m4)
    assert e.best_rate == GOLD["G6_end_to_end"]["rate"]
    assert e.path == [V(p) for p in GOLD["G6_end_to_end"]["path"]]


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["consistent", "arbitrage", "ones", "pow2"])
def test_exact_paths_vs_literal_twin(mode):
    """Every `_path` of a 3-k-block... no: of a 24-vertex graph equals the literal Python twin's list."""
    from floydwarshall_b200 import graphs
    from oracle import fw_oracle as O
    blocks = graphs.exchange_blocks(4, 4, seed=5, density=0.9, mode=mode)
    rmap = graphs.rates_map_from_blocks(blocks)
    ours = A.floyd_warshall({(Vertex(*s), Vertex(*d)): r for (s, d), r in rmap.items()})
    lit = O.floyd_warshall({(O.Vertex(*s), O.Vertex(*d)): r for (s, d), r in rmap.items()})
    assert len(ours) == len(lit)
    got = ours.to_lists()
    for i in range(len(lit)):
        for j in range(len(lit)):
            assert got[i][j].best_rate == lit[i][j].best_rate
            assert [(v.exch, v.ccy) for v in got[i][j].path] == [(v.exch, v.ccy) for v in lit[i][j].path], (i, j)


@pytest.mark.gpu
def test_paths_blocked_graph_vs_oracle_reconstruction():
    """fw_paths on a multi-k-block solve (n = 320) against the oracle's reconstruction."""
    from floydwarshall_b200 import dense, graphs, paths
    from oracle import fw_oracle as O
    rate, nxt = graphs.exchange_graph(20, 16, seed=8, density=0.6)
    res = dense.solve(rate, nxt, paths=True)
    ref = O.solve_dense(rate, nxt, paths=True, threads=0)
    rng = np.random.default_rng(0)
    pairs = [(int(a), int(b)) for a, b in rng.integers(0, 320, size=(500, 2))]
    got = paths.expand(nxt, res.mid, res.csT, res.rs, pairs)
    for (i, j), p in zip(pairs, got):
        assert p == O.reconstruct_path(i, j, nxt, ref.mid, ref.csT, ref.rs)
