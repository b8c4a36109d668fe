"""The C++ host shim (bda::b200SolverBackend<3> behind the reference's BdaSolver<3> interface) end to end on a GPU:
the reference's boundary test tests/test_cusparseSolver.cpp:49-112 restated without Dune/Boost
(opm-autodiff_b200/hostcpp/test_b200Solver.cpp) reads the blocked MatrixMarket files, calls solve_system +
get_result with an empty WellContributions and prints x; expected values: the CPU golden vector of
tests/test_flexiblesolver.cpp:114-116 (tests/golden/matr33.json)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "opm-autodiff_b200", "hostcpp", "test_b200Solver")


def _write_mm(path_m, path_b, rows, cols, vals, b):
    """Blocked MatrixMarket, as MatrixMarketSpecializations.hpp:26-59 writes it."""
    vals = np.asarray(vals).reshape(-1, 3, 3)
    Nb = len(rows) - 1
    with open(path_m, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% ISTL_STRUCT blocked 3 3\n")
        f.write("%d %d %d\n" % (3 * Nb, 3 * Nb, 9 * len(cols)))
        for i in range(Nb):
            for k in range(rows[i], rows[i + 1]):
                for r in range(3):
                    for c in range(3):
                        f.write("%d %d %.17g\n" % (3 * i + r + 1, 3 * cols[k] + c + 1, vals[k, r, c]))
    with open(path_b, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n% ISTL_STRUCT blocked 3 1\n")
        f.write("%d 1\n" % len(b))
        for v in b:
            f.write("%.17g\n" % v)


def _run(tmp_path, rows, cols, vals, b, tol, maxit, env=None):
    if not os.path.exists(BIN):
        pytest.fail("C++ shim test binary not built (run __graft_entry__.build())")
    pm, pb = str(tmp_path / "m.mm"), str(tmp_path / "b.mm")
    _write_mm(pm, pb, rows, cols, vals, b)
    out = subprocess.run([BIN, pm, pb, repr(tol), str(maxit)], capture_output=True, text=True, timeout=120,
                         env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    return lines[0], np.array([float(t) for t in lines[1:]])


def test_matr33_through_the_cpp_backend(built, matr33, tmp_path):
    g = matr33
    head, x = _run(tmp_path, g["rows"], g["cols"], g["vals"], g["b"], 0.5, 20)       # options_flexiblesolver.json: tol 0.5, maxiter 20
    assert "converged 1" in head
    assert np.max(np.abs(x / g["x_golden"] - 1.0)) < 1e-5                             # BOOST_CHECK_CLOSE(.., 1e-3) percent


def test_grid_system_through_the_cpp_backend(built, tmp_path):
    from opm_autodiff_b200 import synth
    from oracle import oracle
    s = synth.small(9, 7, 5, faults=((4, 1),))
    head, x = _run(tmp_path, s.rows, s.cols, s.vals, s.b, 1e-10, 200)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, None, tol=1e-10, maxit=200)
    assert "converged 1" in head
    assert np.linalg.norm(x - ref.x) / np.linalg.norm(ref.x) < 1e-6


def test_multisegment_well_through_the_cpp_surface(built, tmp_path):
    """Opm::WellContributions::addMultisegmentWellContribution of the C++ stand-in (bda_compat.hpp) reaches the device path:
    a one-segment well with B = 0 leaves the solution unchanged."""
    from opm_autodiff_b200 import synth
    from oracle import oracle
    s = synth.small(9, 7, 5)
    head, x = _run(tmp_path, s.rows, s.cols, s.vals, s.b, 1e-10, 200, env={"B200_TEST_MSWELL": "1"})
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, None, tol=1e-10, maxit=200)
    assert "converged 1" in head
    assert np.linalg.norm(x - ref.x) / np.linalg.norm(ref.x) < 1e-6
