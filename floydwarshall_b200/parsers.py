"""Mirror of the reference's Parsers module (/root/reference/src/lib/Parsers.hs): the two
request grammars and their error texts (attoparsec `parseOnly` messages as the reference's
tests pin them, src/test/ParserTest.hs).  Text-only host logic -- not on the GPU path; it
exists so that the callers of the hot path (`serveReq`) can be replayed here without GHC."""
from __future__ import annotations

import re
from datetime import datetime, timezone
from decimal import Decimal
from typing import Tuple

from .types import Vertex


class ParseInputError(Exception):
    """Types.hs:59-60  data ParseError = ParseInputError Text."""

    def __init__(self, msg: str):
        super().__init__(msg)
        self.msg = msg


def show_double(x: float) -> str:
    """Haskell `show :: Double -> String` (shortest digits; fixed for 0.1 <= |x| < 10^7)."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign = "-" if x < 0 else ""
    t = Decimal(repr(abs(x))).as_tuple()
    digits = "".join(map(str, t.digits)).rstrip("0") or "0"
    e = len(t.digits) + t.exponent          # value = 0.d1d2... * 10^e
    if 0.1 <= abs(x) < 1e7:
        if e <= 0:
            return sign + "0." + "0" * (-e) + digits
        ip = digits[:e].ljust(e, "0")
        fp = digits[e:] or "0"
        return sign + ip + "." + fp
    return sign + digits[0] + "." + (digits[1:] or "0") + "e" + str(e - 1)


def show_utc(t: datetime) -> str:
    """Haskell `show :: UTCTime` as the session log prints it (README.md:204)."""
    return t.astimezone(timezone.utc).strftime("%Y-%m-%d %H:%M:%S UTC")


class _Cursor:
    def __init__(self, s: str):
        self.s, self.i = s, 0

    def skip_space(self):                       # attoparsec skipSpace
        while self.i < len(self.s) and self.s[self.i].isspace():
            self.i += 1

    def alphabets(self) -> str:                 # Parsers.hs:57-58  many1 letter
        j = self.i
        while j < len(self.s) and self.s[j].isalpha():
            j += 1
        if j == self.i:
            raise ParseInputError("letter: not enough input" if self.i >= len(self.s)
                                  else "letter: Failed reading: satisfy")
        tok, self.i = self.s[self.i:j], j
        return tok

    _DBL = re.compile(r"[+-]?\d+(?:\.\d+)?(?:[eE][+-]?\d+)?")

    def double(self) -> float:                  # attoparsec `double`
        m = self._DBL.match(self.s, self.i)
        if not m:
            k = self.i + 1 if self.i < len(self.s) and self.s[self.i] in "+-" else self.i
            raise ParseInputError("not enough input" if k >= len(self.s) else "Failed reading: takeWhile1")
        self.i = m.end()
        return float(m.group(0))

    def non_space_token(self) -> str:           # Parsers.hs:27  many1 (satisfy (/= ' '))
        j = self.i
        while j < len(self.s) and self.s[j] != " ":
            j += 1
        if j == self.i:
            raise ParseInputError("not enough input" if self.i >= len(self.s) else "Failed reading: satisfy")
        tok, self.i = self.s[self.i:j], j
        return tok


_TS = re.compile(r"^(\d{4})-(\d{2})-(\d{2})T(\d{2}):(\d{2}):(\d{2})([+-])(\d{2}):?(\d{2})$")


def _parse_timestamp(tok: str) -> datetime:
    """Parsers.hs:39  parseTimeM True defaultTimeLocale "%Y-%m-%dT%H:%M:%S%z"."""
    m = _TS.match(tok)
    try:
        if not m:
            raise ValueError
        return datetime.strptime(tok, "%Y-%m-%dT%H:%M:%S%z").astimezone(timezone.utc)
    except ValueError:
        raise ParseInputError(f'Failed reading: parseTimeM: no parse of "{tok}"') from None


def parse_rates(line: str) -> Tuple[datetime, Vertex, Vertex, float, float]:
    """Parsers.hs:25-40,60-63  exchRatesParser: (time, src vertex, dest vertex, fwd, bkd)."""
    c = _Cursor(line)
    c.skip_space()
    time = _parse_timestamp(c.non_space_token())
    c.skip_space(); exch = c.alphabets()
    c.skip_space(); src = c.alphabets()
    c.skip_space(); dest = c.alphabets()
    c.skip_space(); fwd = c.double()
    if fwd <= 0:                                 # :40 positiveCheck
        raise ParseInputError("Failed reading: Rate must be > 0")
    c.skip_space(); bkd = c.double()
    if bkd <= 0:
        raise ParseInputError("Failed reading: Rate must be > 0")
    if fwd * bkd > 1.0:                          # :34
        raise ParseInputError(f"Failed reading: Product of {show_double(fwd)} and {show_double(bkd)} must be <= 1.0")
    exch, src, dest = exch.upper(), src.upper(), dest.upper()      # :35
    if src == dest:                              # :36
        raise ParseInputError("Failed reading: The currencies must be different")
    return time, Vertex(exch, src), Vertex(exch, dest), fwd, bkd


def parse_exch_pair(line: str) -> Tuple[Vertex, Vertex]:
    """Parsers.hs:46-55,65-68  exchPairParser."""
    c = _Cursor(line)
    toks = []
    for _ in range(4):
        c.skip_space()
        toks.append(c.alphabets().upper())
    src, dest = Vertex(toks[0], toks[1]), Vertex(toks[2], toks[3])
    if src == dest:                              # :54
        raise ParseInputError("Failed reading: source must be different from destination")
    return src, dest
