"""`python -m floydwarshall_b200` (the mirror of src/app/Main.hs:10-37) against the reference's own README
transcript of `cabal run` (tests/golden/readme_session.json, generated from README.md:163-247 by
tests/golden/make_readme_session.py): byte for byte, config C1's query `KRAKEN BTC GDAX USD` included."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SESSION = json.load(open(os.path.join(ROOT, "tests", "golden", "readme_session.json")))["session"]


def run_repl(requests, *flags):
    text = "".join(r["input"] + "\n" for r in requests)
    out = subprocess.run([sys.executable, "-m", "floydwarshall_b200", *flags], input=text, capture_output=True,
                         text=True, cwd=ROOT, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout


def expected(requests):
    return "".join(ln + "\n" for r in requests for ln in r["output"])


def test_rejected_lines_need_no_gpu():
    """The first three requests of the session are rejected by the parsers: the exact texts, no device touched."""
    assert run_repl(SESSION[:3]) == expected(SESSION[:3])


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [(), ("--resident",)])
def test_readme_session_byte_for_byte(flags):
    assert "KRAKEN BTC GDAX USD" in [r["input"] for r in SESSION]      # config C1
    assert run_repl(SESSION, *flags) == expected(SESSION)
