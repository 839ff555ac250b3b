"""`python -m floydwarshall_b200` -- the reference's REPL (src/app/Main.hs:10-37) over the CUDA path.

Reads request lines from stdin until EOF, exactly like `cabal run` of the reference: a line is first tried as
a rate update (`2017-11-01T09:42:23+00:00 KRAKEN BTC USD 1000.0 0.0009`), then as a best-rate request
(`KRAKEN BTC GDAX USD`); the output lines are those of Main.userPrompt (README.md:163-247 is the sample
session, replayed by tests/test_process_requests.py).  Every best-rate request on an OutSync state runs
floydWarshall on the GPU (ProcessRequests.hs:82-84); there is no CPU fallback.

    --resident   keep the optimised matrix in HBM (fw_state_*) and read only the requested entry back
"""
import sys

from .process_requests import blank_state, user_prompt_lines


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    resident = "--resident" in argv
    state = blank_state()                       # Main.hs:11  userPrompt blankState
    for raw in sys.stdin:                       # Main.hs:21  r <- getLine
        line = raw.rstrip("\n")
        state, out = user_prompt_lines(line, state, resident=resident)   # Main.hs:22  run serveReq r
        for ln in out:                          # Main.hs:23  traverse_ putStrLn msgs
            print(ln)
        sys.stdout.flush()
    return 0


if __name__ == "__main__":
    sys.exit(main())
