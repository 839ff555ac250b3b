"""Generates tests/golden/readme_session.json from the reference's README sample session
(/root/reference/README.md:163-247, the transcript of `cabal run`).  Run in the build container, where the
reference is mounted; the GPU box only has the committed JSON.

The transcript interleaves what the user typed with what Main.userPrompt printed.  A line is an INPUT line if
it starts with a timestamp or consists of four alphabetic words; everything up to the next input line is the
expected output of that request (every request's output ends with one empty line, Main.hs:30-37)."""
import json
import os
import re

SRC = "/root/reference/README.md"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    lines = open(SRC).read().split("\n")
    start = next(i for i, ln in enumerate(lines) if ln.startswith("Running floydWarshall...")) + 1
    end = next(i for i in range(start, len(lines)) if lines[i].startswith("```"))
    body = lines[start:end]
    is_input = lambda ln: bool(re.match(r"^\d{4}-\d\d-\d\dT", ln)) or bool(re.match(r"^[A-Za-z]+ [A-Za-z]+ [A-Za-z]+ [A-Za-z]+$", ln))
    session = []
    for ln in body:
        if is_input(ln):
            session.append({"input": ln, "output": []})
        else:
            session[-1]["output"].append(ln)
    # the last request's trailing blank line is cut off by the closing fence of the README block
    if not session[-1]["output"] or session[-1]["output"][-1] != "":
        session[-1]["output"].append("")
    json.dump({"source": "README.md:163-247 of jinilover/floydWarshall", "session": session},
              open(os.path.join(HERE, "readme_session.json"), "w"), indent=1)
    print(len(session), "requests")


if __name__ == "__main__":
    main()
