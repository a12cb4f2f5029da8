#!/usr/bin/env python3
"""Extract the reference's committed golden residual traces into small text fixtures.

Source (read-only, only available in the build container, NOT on the GPU box):
  /root/reference/examples/server_heat/output/residuals.tikz
    lines 29-3187   : CP   xi per iteration (3159 values), produced by
                      examples/server_heat/residuals.jl (N=7, d=2, nx=nu=5, x0=0.1, tol 1e-5, LOG)
    lines 3195-3765 : SPOCK xi, one row per "call of L" (each outer iteration is repeated
                      backtrack_count+2 times; residuals.jl:45-53), 571 rows = 245 iterations.

Outputs (committed):
  tests/golden/residuals_cp.txt   one xi per line (repr round-trip exact)
  tests/golden/residuals_sp.txt   "<xi> <backtrack_count>" per outer iteration

Run:  python tests/golden/make_golden.py
"""
import os
import re
import sys

SRC = "/root/reference/examples/server_heat/output/residuals.tikz"
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_series(path):
    series, cur = [], None
    row = re.compile(r"^\s*([0-9.eE+-]+)\s+([0-9.eE+-]+)\s*\\\\\s*$")
    with open(path) as fh:
        for line in fh:
            if "table[row sep" in line:
                cur = []
                series.append(cur)
                continue
            if cur is not None:
                m = row.match(line)
                if m:
                    cur.append((float(m.group(1)), m.group(2)))
                elif line.strip().startswith("}"):
                    cur = None
    return series


def main():
    if not os.path.exists(SRC):
        sys.exit("reference not mounted; fixtures are already committed")
    cp, sp = parse_series(SRC)
    assert len(cp) == 3159 and len(sp) == 571, (len(cp), len(sp))
    with open(os.path.join(HERE, "residuals_cp.txt"), "w") as fh:
        for _, val in cp:
            fh.write(val + "\n")
    # run-length decode the SP series: each outer iteration appears (backtrack_count + 2) times
    out, i = [], 0
    while i < len(sp):
        j = i
        while j < len(sp) and sp[j][1] == sp[i][1]:
            j += 1
        run = j - i
        # run lengths are 2 (bc=0) or 11 (bc=9); equal consecutive xi never occur in this trace
        assert run in (2, 11), run
        out.append((sp[i][1], run - 2))
        i = j
    assert len(out) == 245, len(out)
    with open(os.path.join(HERE, "residuals_sp.txt"), "w") as fh:
        for val, bc in out:
            fh.write(f"{val} {bc}\n")
    print("wrote", len(cp), "CP values and", len(out), "SP iterations")


if __name__ == "__main__":
    main()
