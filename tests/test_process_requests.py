"""src/test/ProcessRequestsTest.hs and the README session replayed against the mirror of the FSM
(floydwarshall_b200.process_requests).  updateRates rules are CPU-only; anything that triggers
floydWarshall needs the GPU."""
import json
import os
from datetime import datetime, timezone

import pytest

from floydwarshall_b200 import process_requests as PR
from floydwarshall_b200.types import AlgoOptimumError, RateEntry, Vertex

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
kraken_btc, kraken_usd = Vertex("KRAKEN", "BTC"), Vertex("KRAKEN", "USD")
gdax_btc, gdax_usd = Vertex("GDAX", "BTC"), Vertex("GDAX", "USD")
t0942_23, t0943_23, t0942_24 = (datetime.fromtimestamp(s, tz=timezone.utc) for s in (1509529343, 1509529403, 1509529344))

ex_rates1 = {(kraken_btc, kraken_usd): (1000.0, t0942_23), (kraken_usd, kraken_btc): (0.0009, t0942_23)}
ex_rates2 = dict(ex_rates1)
ex_rates2.update({(gdax_btc, gdax_usd): (1001.0, t0943_23), (gdax_usd, gdax_btc): (0.0008, t0943_23)})


# ---------------------------------------------------------------- CPU only
def test_updateRates_addRateToEmptyState():
    s = PR.update_rates("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.0009", PR.blank_state())
    assert isinstance(s, PR.OutSync) and s.ex_rates == ex_rates1            # ProcessRequestsTest.hs:103-106


def test_updateRates_turnStateOutSync():
    for orig in (PR.InSync(ex_rates1, object()), PR.OutSync(ex_rates1)):    # :108-115
        s = PR.update_rates("2017-11-01T09:43:23+00:00 GDAX BTC USD 1001.0 0.0008", orig)
        assert isinstance(s, PR.OutSync) and s.ex_rates == ex_rates2


def test_updateRates_onlyUpdateByNewerTs():
    exp = dict(ex_rates2)
    exp[(kraken_btc, kraken_usd)] = (1001.1, t0942_24)
    exp[(kraken_usd, kraken_btc)] = (0.00089, t0942_24)
    for line in ("2017-11-01T09:42:24+00:00 KRAKEN USD BTC 0.00089 1001.1",
                 "2017-11-01T09:42:24+00:00 kraken usd btc 0.00089 1001.1"):   # :117-129
        s = PR.update_rates(line, PR.OutSync(ex_rates2))
        assert isinstance(s, PR.OutSync) and s.ex_rates == exp


def test_updateRates_notNewerTs():
    for line in ("2017-11-01T09:42:20+00:00 KRAKEN USD BTC 0.00089 1001.1",
                 "2017-11-01T09:42:23+00:00 KRAKEN USD BTC 0.00089 1001.1"):   # :131-137
        for orig in (PR.InSync(ex_rates1, "m"), PR.OutSync(ex_rates1)):
            assert PR.update_rates(line, orig) is orig


def test_serveReq_updateRates():
    s, m = PR.serve_req("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.0009", PR.blank_state())
    assert isinstance(s, PR.OutSync) and s.ex_rates == ex_rates1
    assert m.err == [] and m.res == [
        "(KRAKEN, BTC) -- 1000.0 2017-11-01 09:42:23 UTC --> (KRAKEN, USD)",
        "(KRAKEN, USD) -- 9.0e-4 2017-11-01 09:42:23 UTC --> (KRAKEN, BTC)"]   # :73-80


def test_serveReq_bothInvalid_on_empty_state():
    s, m = PR.serve_req("2017-11-0109:42:23+00:00 KRAKEN BTC USD 1000.0 0.0009", PR.OutSync(ex_rates2))
    assert m.res == [] and m.err == [
        'Failed reading: parseTimeM: no parse of "2017-11-0109:42:23+00:00"',
        "Invalid request to update rates, probably a request for best rate",
        "letter: Failed reading: satisfy"]                                      # :62-71


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
def test_serveReq_findBestRate(resident):
    s, m = PR.serve_req("KRAKEN BTC KRAKEN USD", PR.OutSync(ex_rates2), resident=resident)
    assert isinstance(s, PR.InSync) and s.ex_rates == ex_rates2
    assert m.err == ['Failed reading: parseTimeM: no parse of "KRAKEN"',
                     "Invalid request to update rates, probably a request for best rate"]
    assert m.res == GOLD["G6_end_to_end"]["display"]                            # :82-95


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
def test_findBestRate_unknown_vertices_and_same_result(resident):
    for line, v in (("KRAKEN STC GDAX USD", "(KRAKEN, STC)"), ("KRAKEN USD GDAX STC", "(GDAX, STC)")):
        with pytest.raises(AlgoOptimumError) as ei:
            PR.find_best_rate(line, PR.OutSync(ex_rates2), resident=resident)
        assert ei.value.msg == f"{v} is not entered before"                      # :142-152
    e1, s1 = PR.find_best_rate("KRAKEN BTC KRAKEN USD", PR.OutSync(ex_rates2), resident=resident)
    e2, s2 = PR.find_best_rate("KRAKEN BTC KRAKEN USD", s1, resident=resident)   # InSync: no recompute
    exp = RateEntry(1001.0, kraken_btc, [gdax_btc, gdax_usd, kraken_usd])
    assert e1 == exp and e2 == exp and s2 is s1                                  # :154-162


@pytest.mark.gpu
@pytest.mark.parametrize("resident", [False, True])
def test_readme_session(resident):
    """README.md:163-247, line by line (the invalid lines included)."""
    s = PR.blank_state()
    say = lambda line: PR.user_prompt_lines(line, s, resident=resident)
    s, out = say("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.434d")
    assert out[0] == "Failed reading: Product of 1000.0 and 0.434 must be <= 1.0" and out[2] == "letter: Failed reading: satisfy"
    s, out = say("2017-11-01T09:42:23+00:00 KRAKEN BTC USD -1 0.0")
    assert out[0] == "Failed reading: Rate must be > 0"
    s, out = say("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.0009")
    assert out[:2] == ["(KRAKEN, BTC) -- 1000.0 2017-11-01 09:42:23 UTC --> (KRAKEN, USD)",
                       "(KRAKEN, USD) -- 9.0e-4 2017-11-01 09:42:23 UTC --> (KRAKEN, BTC)"]
    s, out = say("KRAKEN BTC KRAKEN BTC")
    assert out[2] == "Failed reading: source must be different from destination"
    s, out = say("KRAKEN BTC GDAX BTC")
    assert out[2] == "(GDAX, BTC) is not entered before"
    s, out = say("KRAKEN BTC KRAKEN USD")
    assert out[:4] == ["BEST_RATES_BEGIN KRAKEN BTC KRAKEN USD 1000.0", "(KRAKEN, BTC)", "(KRAKEN, USD)", "BEST_RATES_END"]
    s, out = say("KRAKEN USD KRAKEN BTC")
    assert out[0] == "BEST_RATES_BEGIN KRAKEN USD KRAKEN BTC 9.0e-4"
    s, out = say("2017-11-01T09:43:23+00:00 GDAX BTC USD 1001.0 0.0008")
    assert out[:4] == ["(GDAX, BTC) -- 1001.0 2017-11-01 09:43:23 UTC --> (GDAX, USD)",
                       "(GDAX, USD) -- 8.0e-4 2017-11-01 09:43:23 UTC --> (GDAX, BTC)",
                       "(KRAKEN, BTC) -- 1000.0 2017-11-01 09:42:23 UTC --> (KRAKEN, USD)",
                       "(KRAKEN, USD) -- 9.0e-4 2017-11-01 09:42:23 UTC --> (KRAKEN, BTC)"]
    s, out = say("KRAKEN BTC GDAX BTC")
    assert out[:4] == ["BEST_RATES_BEGIN KRAKEN BTC GDAX BTC 1.0", "(KRAKEN, BTC)", "(GDAX, BTC)", "BEST_RATES_END"]
    s, out = say("KRAKEN BTC GDAX USD")          # config C1
    assert out[:5] == ["BEST_RATES_BEGIN KRAKEN BTC GDAX USD 1001.0", "(KRAKEN, BTC)", "(GDAX, BTC)", "(GDAX, USD)",
                       "BEST_RATES_END"]
    s, out = say("GDAX USD KRAKEN BTC")
    assert out[:5] == ["BEST_RATES_BEGIN GDAX USD KRAKEN BTC 9.0e-4", "(GDAX, USD)", "(KRAKEN, USD)", "(KRAKEN, BTC)",
                       "BEST_RATES_END"]


@pytest.mark.gpu
def test_resident_matrix_equals_host_path():
    """fw_state_* (device buildMatrix + resident solve + device query) == pack + fw_solve + fw_paths."""
    import numpy as np
    from floydwarshall_b200 import algorithms as A, graphs
    blocks = graphs.exchange_blocks(12, 12, seed=3, density=0.6)      # n = 144: blocked path, padded
    rmap = {(Vertex(*s), Vertex(*d)): r for (s, d), r in graphs.rates_map_from_blocks(blocks).items()}
    host = A.floyd_warshall(rmap)
    res = PR.ResidentMatrix(rmap)
    r, x = res.download()
    assert np.array_equal(r.view(np.uint64), host.rate.view(np.uint64)) and np.array_equal(x, host.next)
    rng = np.random.default_rng(1)
    for i, j in rng.integers(0, len(host), size=(40, 2)):
        assert res.lookup(int(i), int(j)) == host.entry(int(i), int(j))
    res.close()
