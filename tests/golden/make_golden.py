#!/usr/bin/env python
"""Generates tests/golden/*.json from the reference tree (run in the build container only).

  matr33.json : the reference's 3x3-blocked known-answer system (tests/matr33.txt, tests/rhs3.txt,
                both '% ISTL_STRUCT blocked' MatrixMarket files) converted to raw BSR arrays in
                the layout BdaBridge hands to a backend (BdaBridge.cpp:167-189,231-232), with
                  x_golden = the vector tests/test_flexiblesolver.cpp:114-116 and
                             tests/test_preconditionerfactory.cpp:137-139 expect (1e-3 %),
                  x_direct = numpy.linalg.solve of the same system (agrees with x_golden).
  levels_*.json : toOrder / fromOrder / rowsPerColor returned by the reference's OWN compiled
                findLevelScheduling (oracle/_ref, bda/Reorder.cpp:266-318) on small patterns.

/root/reference does not exist on the GPU box, hence the committed fixtures.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = "/root/reference"


def read_istl_mm_matrix(path):
    with open(path) as f:
        lines = f.read().split("\n")
    assert lines[0].startswith("%%MatrixMarket matrix coordinate real general")
    m = re.match(r"% ISTL_STRUCT blocked (\d+) (\d+)", lines[1])
    br, bc = int(m.group(1)), int(m.group(2))
    n, mcols, nnz = (int(t) for t in lines[2].split())
    dense = {}
    for ln in lines[3:3 + nnz]:
        i, j, v = ln.split()
        dense[(int(i) - 1, int(j) - 1)] = float(v)
    Nb = n // br
    blocks = sorted({(i // br, j // bc) for (i, j) in dense})
    rows = [0]
    cols, vals = [], []
    for I in range(Nb):
        for (bi, bj) in blocks:
            if bi != I:
                continue
            cols.append(bj)
            vals.append([[dense.get((bi * br + r, bj * bc + c), 0.0) for c in range(bc)] for r in range(br)])
        rows.append(len(cols))
    return rows, cols, vals


def read_istl_mm_vector(path):
    with open(path) as f:
        lines = [ln for ln in f.read().split("\n") if ln.strip()]
    assert lines[0].startswith("%%MatrixMarket matrix array real general")
    n = int(lines[2].split()[0])
    return [float(t) for t in lines[3:3 + n]]


def main():
    rows, cols, vals = read_istl_mm_matrix(os.path.join(REF, "tests/matr33.txt"))
    b = read_istl_mm_vector(os.path.join(REF, "tests/rhs3.txt"))
    # tests/test_flexiblesolver.cpp:114-116
    x_golden = [-1.62493, -1.76435e-06, 1.86991e-10, -458.542, 2.28308e-06, -2.45341e-07,
                -1.48005, -5.02264e-07, -1.049e-05]
    Nb = len(rows) - 1
    A = np.zeros((3 * Nb, 3 * Nb))
    for i in range(Nb):
        for k in range(rows[i], rows[i + 1]):
            A[3 * i:3 * i + 3, 3 * cols[k]:3 * cols[k] + 3] = np.array(vals[k])
    x_direct = np.linalg.solve(A, np.array(b))
    out = {"source": "reference tests/matr33.txt + tests/rhs3.txt; golden tests/test_flexiblesolver.cpp:114-116",
           "Nb": Nb, "rows": rows, "cols": cols, "vals": vals, "b": b,
           "x_golden": x_golden, "x_golden_rel_tol": 1e-5, "x_direct": x_direct.tolist()}
    with open(os.path.join(HERE, "matr33.json"), "w") as f:
        json.dump(out, f, indent=1)

    # level-set fixtures from the compiled reference function
    from oracle import oracle
    from tests.patterns import grid_pattern
    cases = {"grid_4x3x1": (4, 3, 1, 0), "grid_5x4x3": (5, 4, 3, 0), "grid_6x5x4_nnc": (6, 5, 4, 2)}
    for name, (nx, ny, nz, nnc) in cases.items():
        r, c = grid_pattern(nx, ny, nz, nnc)
        to, fr, lp = oracle.ref_level_schedule(r, c)
        with open(os.path.join(HERE, "levels_%s.json" % name), "w") as f:
            json.dump({"source": "bda::findLevelScheduling (Reorder.cpp:266-318) compiled unmodified",
                       "grid": [nx, ny, nz], "nnc_planes": nnc, "rows": r.tolist(), "cols": c.tolist(),
                       "toOrder": to.tolist(), "fromOrder": fr.tolist(),
                       "rowsPerColor": np.diff(lp).tolist()}, f)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
