"""Blocked MatrixMarket files as Flow dumps them.

With a linear-solver verbosity above 10 Flow writes every linear system it solves as
``reports/prob_<episode>_time_<t>__nit_<n>_matrix_istl.mm`` / ``..._rhs_istl.mm``
(ISTLSolverEbos.hpp:245-252 -> WriteSystemMatrixHelper.hpp:31-84); the header carries
``% ISTL_STRUCT blocked <rows> <cols>`` (MatrixMarketSpecializations.hpp:26-59) and the reference's own
fixtures use the same format (tests/matr33.txt:1-3, tests/rhs3.txt:1-3).  ``read_matrix`` turns such a
file into the raw BSR arrays a BdaBridge hands its backend (BdaBridge.cpp:167-189,231-232: columns
ascending per row, row-major 3x3 blocks), so that systems dumped by a real Flow run elsewhere can be
replayed through ``bench.py --workload mm:<matrix>,<rhs>`` and the parity tests.
"""
from __future__ import annotations

import numpy as np


def _header(f, kind):
    first = f.readline()
    if not first.startswith("%%MatrixMarket matrix " + kind + " real general"):
        raise ValueError("not a %s real general MatrixMarket file" % kind)
    br = bc = 1
    line = f.readline()
    while line.startswith("%"):
        t = line.split()
        if len(t) >= 5 and t[1] == "ISTL_STRUCT" and t[2] == "blocked":
            br, bc = int(t[3]), int(t[4])
        line = f.readline()
    return br, bc, line


def read_matrix(path, block=3):
    """-> (rows int32 [Nb+1], cols int32 [nnzb], vals float64 [nnzb, bs, bs]).  A file without the ISTL_STRUCT line is
    taken as a scalar matrix and cut into `block` x `block` blocks (every block touched by an entry is stored)."""
    with open(path) as f:
        br, bc, size = _header(f, "coordinate")
        n, m, nnz = (int(t) for t in size.split())
        data = np.loadtxt(f, dtype=np.float64, ndmin=2) if nnz else np.zeros((0, 3))
    if data.shape[0] != nnz:
        raise ValueError("truncated MatrixMarket file %s" % path)
    if br == 1 and bc == 1:
        br = bc = block
    if br != bc or n != m or n % br:
        raise ValueError("only square matrices of square blocks are supported")
    bs, Nb = br, n // br
    i = data[:, 0].astype(np.int64) - 1
    j = data[:, 1].astype(np.int64) - 1
    key = (i // bs) * Nb + (j // bs)
    ukey, inv = np.unique(key, return_inverse=True)
    vals = np.zeros((len(ukey), bs, bs))
    vals[inv, i % bs, j % bs] = data[:, 2]
    brow = (ukey // Nb).astype(np.int64)
    rows = np.zeros(Nb + 1, np.int32)
    np.add.at(rows, brow + 1, 1)
    rows = np.cumsum(rows).astype(np.int32)
    return rows, (ukey % Nb).astype(np.int32), vals


def read_vector(path):
    with open(path) as f:
        _, _, size = _header(f, "array")
        t = size.split()
        n, m = int(t[0]), int(t[1]) if len(t) > 1 else 1
        v = np.loadtxt(f, dtype=np.float64).reshape(-1)
    if v.size != n * m:
        raise ValueError("truncated MatrixMarket file %s" % path)
    return v


def write_matrix(path, rows, cols, vals):
    vals = np.asarray(vals)
    bs = vals.shape[1]
    Nb = len(rows) - 1
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("% ISTL_STRUCT blocked {0} {0}\n".format(bs))
        f.write("%d %d %d\n" % (bs * Nb, bs * Nb, bs * bs * len(cols)))
        for r in range(Nb):
            for k in range(rows[r], rows[r + 1]):
                for a in range(bs):
                    for b in range(bs):
                        f.write("%d %d %.17g\n" % (bs * r + a + 1, bs * cols[k] + b + 1, vals[k, a, b]))


def write_vector(path, v, bs=3):
    v = np.asarray(v).reshape(-1)
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n")
        f.write("% ISTL_STRUCT blocked {0} 1\n".format(bs))
        f.write("%d 1\n" % len(v))
        for x in v:
            f.write("%.17g\n" % x)
