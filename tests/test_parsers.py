"""src/test/ParserTest.hs replayed against floydwarshall_b200.parsers (host logic, CPU only)."""
from datetime import datetime, timezone

import pytest

from floydwarshall_b200.parsers import ParseInputError, parse_exch_pair, parse_rates, show_double
from floydwarshall_b200.types import Vertex

kraken_btc, kraken_usd, gdax_usd = Vertex("KRAKEN", "BTC"), Vertex("KRAKEN", "USD"), Vertex("GDAX", "USD")


def ts(sec):
    return datetime.fromtimestamp(sec, tz=timezone.utc)


@pytest.mark.parametrize("line,err", [
    ("2017-11-01T09:42:3+00:00 KRAKEN BTC USD 1000.0 0.0009",
     'Failed reading: parseTimeM: no parse of "2017-11-01T09:42:3+00:00"'),            # ParserTest.hs:46-49
    ("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.0091",
     "Failed reading: Product of 1000.0 and 9.1e-3 must be <= 1.0"),                   # :51-54
    ("2017-11-01T09:42:23+00:00 KRAKEN BTC BTC 1000.0 0.0009",
     "Failed reading: The currencies must be different"),                              # :56-59
    ("2017-11-01T09:42:23+00:00 KRAKEN btc BTC 1000.0 0.0009",
     "Failed reading: The currencies must be different"),                              # :61-64
    ("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 0 0.0009", "Failed reading: Rate must be > 0"),       # :66-69
    ("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000 -0.0009", "Failed reading: Rate must be > 0"),   # :71-74
    ("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0x 0.0009", "Failed reading: takeWhile1"),       # :76-79
])
def test_parseRates_invalid(line, err):
    with pytest.raises(ParseInputError) as ei:
        parse_rates(line)
    assert ei.value.msg == err


def test_parseRates_valid():
    exp = (ts(1509529343), kraken_btc, kraken_usd, 1000.0, 0.0009)
    assert parse_rates("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.0009") == exp              # :81-84
    assert parse_rates("  2017-11-01T09:42:23+00:00    KRAKEN   BTC  USD   1000.0  0.0009  ") == exp  # :86-89
    assert parse_rates("2017-11-01T09:42:24+00:00 kraken usd btc 1000.0 0.0009") == \
        (ts(1509529344), kraken_usd, kraken_btc, 1000.0, 0.0009)                                       # :91-94


def test_parseExchPair():
    with pytest.raises(ParseInputError) as ei:
        parse_exch_pair("KRAKEN BTC KRAKEN BTC")
    assert ei.value.msg == "Failed reading: source must be different from destination"               # :99-102
    assert parse_exch_pair("KRAKEN BTC GDAX USD") == (kraken_btc, gdax_usd)                           # :104-106
    assert parse_exch_pair("    KRAKEN BTC   GDAX    USD  ") == (kraken_btc, gdax_usd)                # :108-110
    assert parse_exch_pair("kraken btc gdax usd") == (kraken_btc, gdax_usd)                           # :112-114
    with pytest.raises(ParseInputError) as ei:
        parse_exch_pair("2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.0009")
    assert ei.value.msg == "letter: Failed reading: satisfy"                                          # ProcessRequestsTest.hs:66-68


def test_show_double_matches_haskell_show():
    for x, s in [(1000.0, "1000.0"), (0.0009, "9.0e-4"), (0.0008, "8.0e-4"), (1001.0, "1001.0"), (1.0, "1.0"),
                 (0.0091, "9.1e-3"), (0.434, "0.434"), (0.1, "0.1"), (1e7, "1.0e7"), (1234567.5, "1234567.5"),
                 (1001.1, "1001.1"), (0.00089, "8.9e-4"), (5e-324, "5.0e-324")]:
        assert show_double(x) == s
