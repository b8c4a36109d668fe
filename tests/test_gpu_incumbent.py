"""Parity against the reference's OWN GPU backend (bda::cusparseSolverBackend<3> + Opm::WellContributions, compiled
unmodified from /root/reference into oracle/_ref/libref_cusparse.so by oracle/Makefile; the prebuilt library travels to
the GPU box).  This pins what the reference's CPU tests cannot: a full ILU0-BiCGSTAB solve on a grid system, and the
standard-well apply y -= C^T D^-1 B x through apply_well_contributions (bda/WellContributions.cu:36-126) -- on wells of
at most 10 perforations, the range in which that kernel applies every perforation (it runs 32 threads, :115-124,192)."""
import os

import numpy as np
import pytest

from tests.helpers import bridge_wells, oracle_wells, relerr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_cusparse.so")


@pytest.fixture(scope="module")
def mods(built):
    from opm_autodiff_b200 import bridge, synth
    from oracle import oracle
    if not bridge.device_available():
        pytest.fail("GPU tests need a B200; the product has no CPU fallback")
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libref_cusparse.so was not built (needs /root/reference at build time)")
    from tests import incumbent_cusparse as inc
    return bridge, synth, oracle, inc


@pytest.mark.parametrize("shape,faults,nwells,nperf", [((12, 10, 8), (), 0, 0), ((20, 16, 12), ((7, 2),), 0, 0),
                                                       ((12, 10, 8), ((6, 1),), 3, 4), ((24, 20, 16), (), 6, 9)])
def test_solution_matches_the_reference_cusparse_backend(mods, shape, faults, nwells, nperf):
    bridge, synth, oracle, inc = mods
    s = synth.small(*shape, faults=faults, nwells=nwells, nperf=nperf)
    L = inc.ref_lib()
    xi, wall, iters, red, conv = inc.run_incumbent(L, s, s.wells, 1e-10, 200, 1)
    assert conv[0] == 1
    be = bridge.B200SolverBackend(0, 200, 1e-10, 0)
    res = bridge.BdaResult()
    be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells), res)
    x = np.zeros(3 * s.Nb)
    be.get_result(x)
    assert res.converged
    assert relerr(x, xi) <= 1e-6                                    # BASELINE.json's bar, against the reference itself
    assert abs(res.it - iters[0]) <= max(1.0, 0.1 * iters[0])       # the reference reports floor(it) (cusparseSolverBackend.cu:172)
    # and the reference agrees with the oracle on the same system (the oracle is what the other parity tests use)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200)
    assert relerr(xi, ref.x) <= 1e-6
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, xi, oracle_wells(s.wells)) < 2e-10
