"""Pin the CPU oracle against every golden vector the reference's own tests hold
for the matrix-optimisation path."""
import json
import os

import numpy as np
import pytest

from oracle import fw_oracle as O
from floydwarshall_b200 import graphs

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))


def V(p):
    return O.Vertex(p[0], p[1])


def mock_map(*names):
    m = {}
    for nme in names:
        s, d, r = GOLD["mock_rates"][nme]
        m[(V(s), V(d))] = r
    return m


ALL4 = ("gdax_btc_usd", "kraken_btc_usd", "gdax_usd_btc", "kraken_usd_btc")


def check_matrix(matrix, expected, vertices):
    assert len(matrix) == len(expected)
    for i, row in enumerate(expected):
        assert len(matrix[i]) == len(row)
        for j, (rate, path) in enumerate(row):
            e = matrix[i][j]
            assert e.best_rate == float(rate), (i, j, e.best_rate, rate)
            assert e.start == vertices[i]
            assert e.path == [vertices[p] for p in path], (i, j, e.path, path)


def test_G2_build_matrix_4x4():
    vs = [V(p) for p in GOLD["vertices_4x4"]]
    m = O.build_matrix(mock_map(*ALL4))
    check_matrix(m, GOLD["G2_buildMatrix_4x4"]["matrix"], vs)


def test_G1_floyd_warshall_4x4_literal():
    vs = [V(p) for p in GOLD["vertices_4x4"]]
    m = O.floyd_warshall(mock_map(*ALL4))
    check_matrix(m, GOLD["G1_floydWarshall_4x4"]["matrix"], vs)


def test_G3_empty():
    assert O.build_matrix({}) == []
    assert O.floyd_warshall({}) == []
    r = O.solve_dense(np.zeros((0, 0)), np.zeros((0, 0), dtype=np.int32))
    assert r.rate.shape == (0, 0)


@pytest.mark.parametrize("literal", [True, False])
def test_G1_dense_oracle_matches_golden(literal):
    """The C loop on the dense encoding reproduces G1 rates, next-hops AND exact paths."""
    m0 = O.build_matrix(mock_map(*ALL4))
    _, rate0, next0, _ = O.dense_from_entries(m0)
    res = O.solve_dense(rate0, next0, paths=True, literal=literal)
    exp = GOLD["G1_floydWarshall_4x4"]["matrix"]
    for i in range(4):
        for j in range(4):
            rate, path = exp[i][j]
            assert res.rate[i, j] == float(rate)
            assert res.next[i, j] == (path[0] if path else -1)
            assert O.reconstruct_path(i, j, next0, res.mid, res.csT, res.rs) == path


def test_readme_graph_generator_is_G2():
    rate, nxt = graphs.readme_graph()
    exp = GOLD["G2_buildMatrix_4x4"]["matrix"]
    for i in range(4):
        for j in range(4):
            assert rate[i, j] == float(exp[i][j][0])
            assert nxt[i, j] == (exp[i][j][1][0] if exp[i][j][1] else -1)


def test_G4_optimum_unknown_vertex():
    m = O.floyd_warshall(mock_map(*ALL4))
    g = GOLD["G4_optimum"]["unknown_vertex"]
    for s, d in g["queries"]:
        with pytest.raises(O.AlgoOptimumError) as ei:
            O.optimum(V(s), V(d), m)
        assert str(ei.value) == g["error"]


def test_G4_optimum_reachability():
    m = O.floyd_warshall(mock_map(*ALL4))
    g = GOLD["G4_optimum"]["reachability"]
    i, j = g["isolate"]
    m[i][j] = O.isolated_entry(m[i][0].start)
    s, d = g["unreachable"]["query"]
    with pytest.raises(O.AlgoOptimumError) as ei:
        O.optimum(V(s), V(d), m)
    assert str(ei.value) == g["unreachable"]["error"]
    for a in g["answers"]:
        s, d = a["query"]
        e = O.optimum(V(s), V(d), m)
        assert e.best_rate == a["rate"] and e.start == V(s)
        assert e.path == [V(p) for p in a["path"]]


def test_G5_all_ones_generative():
    """AlgorithmsTest.hs:112-134 with MockData.genRateMatrix, enumerated instead of sampled."""
    sv = [V(p) for p in GOLD["G5_all_ones"]["sample_vertices"]]
    rng = np.random.default_rng(5)
    for trial in range(60):
        k = int(rng.integers(0, len(sv) // 2 + 2))
        verts = sorted(set(sv[int(x)] for x in rng.integers(0, len(sv), size=k)))
        empty_row = bool(rng.integers(0, 2))
        if empty_row:
            matrix = [[] for _ in verts]
        else:
            matrix = [[O.isolated_entry(s) if s == d else O.RateEntry(1.0, s, [d]) for d in verts]
                      for s in verts]
        src = sv[int(rng.integers(0, len(sv)))]
        dest = sv[int(rng.integers(0, len(sv)))]
        try:
            got = O.optimum(src, dest, matrix)
            err = None
        except O.AlgoOptimumError as ex:
            got, err = None, str(ex)
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
            assert err is None and got.best_rate == 1.0 and got.start == src and got.path == [dest]


def test_G6_end_to_end_answer():
    m = O.floyd_warshall(mock_map(*ALL4))
    g = GOLD["G6_end_to_end"]
    s, d = g["query"]
    e = O.optimum(V(s), V(d), m)
    assert e.best_rate == g["rate"]
    assert e.path == [V(p) for p in g["path"]]


def test_G7_readme_session():
    g = GOLD["G7_readme_session"]
    m2 = O.floyd_warshall(mock_map("kraken_btc_usd", "kraken_usd_btc"))
    for a in g["kraken_only"]:
        e = O.optimum(V(a["query"][0]), V(a["query"][1]), m2)
        assert e.best_rate == a["rate"] and e.path == [V(p) for p in a["path"]]
    m4 = O.floyd_warshall(mock_map(*ALL4))
    for a in g["four_vertices"]:
        e = O.optimum(V(a["query"][0]), V(a["query"][1]), m4)
        assert e.best_rate == a["rate"] and e.path == [V(p) for p in a["path"]]


@pytest.mark.parametrize("mode", graphs.MODES)
@pytest.mark.parametrize("E,C", [(2, 3), (3, 4), (4, 4)])
def test_dense_oracle_equals_literal_twin(E, C, mode):
    """C loop (both forms) == literal Python twin incl. exact paths, on tie-heavy and arbitrage graphs."""
    blocks = graphs.exchange_blocks(E, C, seed=11 + E * C, density=0.8, mode=mode)
    rmap = {(O.Vertex(*s), O.Vertex(*d)): r for (s, d), r in graphs.rates_map_from_blocks(blocks).items()}
    lit = O.floyd_warshall(rmap)
    verts, lrate, lnext, lpaths = O.dense_from_entries(lit)
    names = graphs.vertex_names(E, C)
    # vertices that appear in no map entry are absent from the reference matrix; keep only full graphs
    if len(verts) != E * C:
        pytest.skip("sparse draw dropped a vertex")
    assert [(v.exch, v.ccy) for v in verts] == names
    rate0, next0 = graphs.dense_from_blocks(blocks)
    m0 = O.build_matrix(rmap)
    _, brate, bnext, _ = O.dense_from_entries(m0)
    assert np.array_equal(brate, rate0) and np.array_equal(bnext, next0)
    for literal in (True, False):
        res = O.solve_dense(rate0, next0, paths=True, literal=literal, threads=2)
        assert np.array_equal(res.rate, lrate)
        assert np.array_equal(res.next, lnext)
        n = E * C
        for i in range(n):
            for j in range(n):
                try:
                    p = O.reconstruct_path(i, j, next0, res.mid, res.csT, res.rs, cap=50000)
                except OverflowError:
                    continue
                assert p == lpaths[i][j], (i, j)


@pytest.mark.parametrize("mode", graphs.MODES)
def test_inplace_omp_equals_generations(mode):
    rate0, next0 = graphs.exchange_graph(12, 8, seed=3, density=0.9, mode=mode)
    a = O.solve_dense(rate0, next0, paths=True, literal=True)
    b = O.solve_dense(rate0, next0, paths=True, literal=False, threads=4)
    for f in ("rate", "next", "mid", "csT", "rs"):
        assert np.array_equal(getattr(a, f), getattr(b, f), equal_nan=True), f
    assert a.updates == b.updates
