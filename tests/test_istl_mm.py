"""Blocked MatrixMarket reader/writer (the format Flow dumps its linear systems in)."""
import numpy as np

from tests.helpers import relerr


def test_round_trip_and_oracle_solve(built, matr33, tmp_path):
    from opm_autodiff_b200 import istl_mm
    from oracle import oracle
    g = matr33
    pm, pb = str(tmp_path / "m.mm"), str(tmp_path / "b.mm")
    istl_mm.write_matrix(pm, g["rows"], g["cols"], g["vals"].reshape(-1, 3, 3))
    istl_mm.write_vector(pb, g["b"])
    with open(pm) as f:
        assert f.readline().startswith("%%MatrixMarket matrix coordinate real general") and f.readline() == "% ISTL_STRUCT blocked 3 3\n"
    rows, cols, vals = istl_mm.read_matrix(pm)
    b = istl_mm.read_vector(pb)
    assert np.array_equal(rows, g["rows"]) and np.array_equal(cols, g["cols"])
    assert np.array_equal(vals.reshape(-1), g["vals"].reshape(-1)) and np.array_equal(b, g["b"])
    ref = oracle.solve(rows, cols, vals, b, tol=1e-10, maxit=50)
    assert np.max(np.abs(ref.x / g["x_golden"] - 1.0)) < 1e-5       # tests/test_flexiblesolver.cpp:114-116


def test_grid_system_round_trip(built, tmp_path):
    from opm_autodiff_b200 import istl_mm, synth
    s = synth.small(5, 4, 3, faults=((2, 1),))
    pm, pb = str(tmp_path / "m.mm"), str(tmp_path / "b.mm")
    istl_mm.write_matrix(pm, s.rows, s.cols, s.vals)
    istl_mm.write_vector(pb, s.b)
    rows, cols, vals = istl_mm.read_matrix(pm)
    assert np.array_equal(rows, s.rows) and np.array_equal(cols, s.cols) and relerr(vals, s.vals) == 0.0
    assert np.array_equal(istl_mm.read_vector(pb), s.b)
