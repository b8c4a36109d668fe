"""CPU tests of the ORACLE against the reference's golden vectors / known-answer properties."""
import glob
import json
import os

import numpy as np
import pytest

from oracle import oracle
from tests.helpers import dense_from_bsr, dense_well_operator, oracle_wells, relerr
from tests.patterns import grid_pattern

HERE = os.path.dirname(os.path.abspath(__file__))


def test_golden_matr33_solution(matr33, built):
    """tests/test_flexiblesolver.cpp:114-116 expects x to 1e-3 % (BOOST_CHECK_CLOSE); ILU0 is exact on
    this block-tridiagonal system so BiCGSTAB stops in the first half step at any tolerance."""
    g = matr33
    for tol in (0.5, 1e-2, 1e-10):
        r = oracle.solve(g["rows"], g["cols"], g["vals"], g["b"], tol=tol, maxit=20)
        assert r.converged and r.it == 0.5
        assert np.max(np.abs(r.x / g["x_golden"] - 1.0)) < 1e-5      # 1e-3 percent
        assert np.max(np.abs(r.x / g["x_direct"] - 1.0)) < 1e-6


def test_inv3_matches_numpy(built):
    rng = np.random.default_rng(0)
    for _ in range(50):
        m = rng.normal(size=(3, 3)) * np.array([1e-7, 1.0, 1.0])
        inv = np.empty(9)
        assert oracle.lib().orc_inv3(np.ascontiguousarray(m.reshape(-1)), inv) == 0
        assert relerr(inv.reshape(3, 3), np.linalg.inv(m)) < 1e-10


def _laplace2d_blocks(n, seed=1):
    """Pattern of tests/test_milu.cpp:102-130 (5-point, N x N) with diagonally dominant 3x3 blocks."""
    rows, cols = grid_pattern(n, n, 1)
    rng = np.random.default_rng(seed)
    nnzb = rows[-1]
    vals = np.zeros((nnzb, 3, 3))
    for i in range(len(rows) - 1):
        for k in range(rows[i], rows[i + 1]):
            vals[k] = (4.0 * np.eye(3) + 0.3 * rng.normal(size=(3, 3))) if cols[k] == i \
                else (-1.0 * np.eye(3) + 0.1 * rng.normal(size=(3, 3)))
    return rows, cols, vals


def test_ilu0_factor_property(built):
    """Restated tests/test_milu.cpp:42-99 for plain ILU0: on a 5-point pattern ILU0 has no fill to drop
    only inside the pattern, so (L U) agrees with A on the pattern; and (LU)^-1 (L U e) = e."""
    rows, cols, vals = _laplace2d_blocks(6)
    LU, diag, st = oracle.ilu0(rows, cols, vals)
    assert st == 0
    Nb = len(rows) - 1
    # dense L (unit) and U (pivot stored inverted)
    L = np.eye(3 * Nb)
    U = np.zeros((3 * Nb, 3 * Nb))
    for i in range(Nb):
        for k in range(rows[i], rows[i + 1]):
            j = cols[k]
            blk = LU[k]
            if j < i:
                L[3 * i:3 * i + 3, 3 * j:3 * j + 3] = blk
            elif j == i:
                U[3 * i:3 * i + 3, 3 * j:3 * j + 3] = np.linalg.inv(blk)
            else:
                U[3 * i:3 * i + 3, 3 * j:3 * j + 3] = blk
    A = dense_from_bsr(rows, cols, vals)
    P = (L @ U)
    mask = A != 0
    assert np.max(np.abs(P[mask] - A[mask])) < 1e-12                 # ILU0 reproduces A on its pattern
    e = np.ones(3 * Nb)
    d = P @ e
    v = oracle.ilu0_apply(rows, cols, diag, LU, d)
    assert relerr(v, e) < 1e-12                                       # (LU)^-1 LU e = e
    v09 = oracle.ilu0_apply(rows, cols, diag, LU, d, w=0.9)
    assert relerr(v09, 0.9 * e) < 1e-12                               # relaxation :899-901


def test_ilu0_is_exact_lu_for_block_tridiagonal(matr33, built):
    g = matr33
    LU, diag, st = oracle.ilu0(g["rows"], g["cols"], g["vals"])
    assert st == 0
    x = oracle.ilu0_apply(g["rows"], g["cols"], diag, LU, g["b"])
    assert np.max(np.abs(x / g["x_direct"] - 1.0)) < 1e-6


def test_ilu0_errors(built):
    rows = np.array([0, 1, 3], np.int32)
    cols = np.array([0, 0, 1], np.int32)
    vals = np.zeros((3, 3, 3))
    vals[0] = np.eye(3); vals[1] = np.eye(3); vals[2] = np.zeros((3, 3))
    _, _, st = oracle.ilu0(rows, cols, vals)
    assert st == 2                                                    # "ILU failed to invert matrix block"
    rows2 = np.array([0, 1, 2], np.int32)
    cols2 = np.array([0, 0], np.int32)
    _, _, st = oracle.ilu0(rows2, cols2, np.stack([np.eye(3)] * 2))
    assert st == 1                                                    # "diagonal entry missing"


def test_check_zero_diagonal(built):
    rows, cols, vals = _laplace2d_blocks(3)
    d0 = int(np.nonzero(cols[rows[0]:rows[1]] == 0)[0][0])
    vals[d0, 1, 1] = 0.0
    vals[d0, 0, 1] = 0.0          # off-diagonal scalar zero must stay
    n = oracle.check_zero_diagonal(rows, cols, vals)
    assert n == 1 and vals[d0, 1, 1] == 1e-15 and vals[d0, 0, 1] == 0.0
    assert oracle.check_zero_diagonal(rows, cols, vals) == 0


def test_spmv_and_wells_vs_dense(built):
    from opm_autodiff_b200 import synth
    s = synth.small(5, 4, 3, faults=((2, 1),), nwells=2, nperf=3)
    A = dense_from_bsr(s.rows, s.cols, s.vals)
    W = dense_well_operator(s.wells, s.Nb)
    rng = np.random.default_rng(3)
    x = rng.normal(size=3 * s.Nb)
    y = oracle.spmv(s.rows, s.cols, s.vals, x)
    assert relerr(y, A @ x) < 1e-13
    y2 = oracle.well_apply(oracle_wells(s.wells), x, y)
    assert relerr(y2, (A - W) @ x) < 1e-12
    # b was generated as (A - C^T D^-1 B) x_true
    assert relerr((A - W) @ s.x_true, s.b) < 1e-11


def test_well_apply_more_than_ten_perforations_and_shared_cell(built):
    """The reference GPU kernels silently drop perforations beyond the 10th (WellContributions.cu:115-124);
    the CPU operator (StandardWell_impl.hpp:1251-1277) does not -- the oracle follows the CPU one."""
    rng = np.random.default_rng(5)
    Nb, P = 40, 23
    cells = rng.permutation(Nb)[:P].astype(np.int32)
    cells2 = np.concatenate([cells[:3], rng.permutation(Nb)[:4]]).astype(np.int32)   # shares cells with well 0
    w = oracle.Wells(np.array([0, P, P + len(cells2)], np.uint32), np.concatenate([cells, cells2]),
                     np.concatenate([cells, cells2]), rng.normal(size=(P + len(cells2), 4, 3)),
                     rng.normal(size=(P + len(cells2), 4, 3)), rng.normal(size=(2, 4, 4)))
    x = rng.normal(size=3 * Nb)
    y0 = rng.normal(size=3 * Nb)
    y = oracle.well_apply(w, x, y0)
    assert relerr(y, y0 - dense_well_operator(w, Nb) @ x) < 1e-12


def test_bicgstab_matches_direct_solve(built):
    from opm_autodiff_b200 import synth
    s = synth.small(6, 5, 4, faults=((3, 1),), nwells=2, nperf=3)
    A = dense_from_bsr(s.rows, s.cols, s.vals) - dense_well_operator(s.wells, s.Nb)
    xd = np.linalg.solve(A, s.b)
    r = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-12, maxit=200)
    assert r.converged and relerr(r.x, xd) < 1e-8
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, r.x, oracle_wells(s.wells)) < 1e-11
    assert len(r.history) == int(2 * r.it) + 1 and r.history[-1] < 1e-12 * r.history[0]
    # non-convergence is reported, not raised
    r2 = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-14, maxit=2)
    assert (not r2.converged) and r2.it == 2.0


def test_block_jacobi_partitions(built):
    """G > 1 oracle: ILU0 per row slab with couplings across slabs dropped
    (PreconditionerFactory.hpp:218-252, ParallelOverlappingILU0.hpp:857-897)."""
    from opm_autodiff_b200 import synth
    s = synth.small(6, 5, 8)
    plane = 30
    part = np.array([0, 4 * plane, 8 * plane], np.int32)
    r1 = oracle.solve(s.rows, s.cols, s.vals, s.b, tol=1e-10, maxit=400)
    r2 = oracle.solve(s.rows, s.cols, s.vals, s.b, tol=1e-10, maxit=400, part_ptr=part)
    assert r1.converged and r2.converged
    assert relerr(r2.x, r1.x) < 1e-6 and r2.it >= r1.it
    # explicit check of the partitioned preconditioner: equals ILU0 of the block-diagonal part
    keep = (np.repeat(np.arange(s.Nb), np.diff(s.rows)) // (4 * plane)) == (s.cols // (4 * plane))
    rows_f = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), s.rows[:-1]))]).astype(np.int32)
    LUf, diagf, st = oracle.ilu0(rows_f, s.cols[keep], s.vals[keep])
    assert st == 0
    d = np.random.default_rng(1).normal(size=3 * s.Nb)
    vf = oracle.ilu0_apply(rows_f, s.cols[keep], diagf, LUf, d)
    LU = s.vals.copy().reshape(-1)
    diag = np.zeros(s.Nb, np.int32)
    L = oracle.lib()
    v = np.zeros(3 * s.Nb)
    for p in range(2):
        assert L.orc_ilu0_decompose_range(s.rows, s.cols.astype(np.int32), LU, diag, int(part[p]), int(part[p + 1])) == 0
    for p in range(2):
        L.orc_ilu0_apply_range(s.rows, s.cols.astype(np.int32), diag, LU, d, v, 1.0, int(part[p]), int(part[p + 1]))
    assert relerr(v, vf) < 1e-13


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(HERE, "golden", "levels_*.json"))))
def test_level_schedule_matches_reference_fixture(fixture, built):
    """Level sets / permutation pinned by the reference's own compiled findLevelScheduling."""
    with open(fixture) as f:
        g = json.load(f)
    rows, cols = np.array(g["rows"], np.int32), np.array(g["cols"], np.int32)
    to, fr, lp = oracle.level_schedule(rows, cols)
    assert to.tolist() == g["toOrder"] and fr.tolist() == g["fromOrder"]
    assert np.diff(lp).tolist() == g["rowsPerColor"]


def test_level_schedule_vs_live_reference(built):
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    for (nx, ny, nz, nnc) in ((4, 3, 1, 0), (7, 5, 4, 0), (8, 6, 5, 2), (1, 1, 9, 0), (3, 3, 3, 1)):
        rows, cols = grid_pattern(nx, ny, nz, nnc)
        to, fr, lp = oracle.level_schedule(rows, cols)
        rto, rfr, rlp = oracle.ref_level_schedule(rows, cols)
        assert np.array_equal(to, rto) and np.array_equal(fr, rfr) and np.array_equal(lp, rlp)
        if nnc == 0:
            assert len(lp) - 1 == nx + ny + nz - 2


def test_multisegment_well_apply_against_dense_algebra():
    """The oracle's restatement of MultisegmentWellContribution::apply (bda/MultisegmentWellContribution.cpp:70-110, dense
    LU with partial pivoting in place of UMFPACK) against C^T D^-1 B built densely with numpy; and the operator of the
    oracle's solve applies multisegment wells before the standard ones (bda/WellContributions.cu:167-193).
    No reference test builds a MultisegmentWellContribution; the next test pins the restatement against the class itself."""
    from opm_autodiff_b200 import synth
    from tests.helpers import dense_mswell_operator, dense_well_operator, dense_from_bsr, oracle_mswells, oracle_wells, relerr
    s = synth.small(7, 6, 5, nwells=2, nperf=3)
    ms = synth.add_mswells(s, 3, 6, seed=5)
    assert any(len(set(a.BcolIndices.tolist()) & set(b.BcolIndices.tolist())) for a in ms for b in ms if a is not b)
    om = oracle_mswells(ms)
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal(3 * s.Nb), rng.standard_normal(3 * s.Nb)
    Mop = dense_mswell_operator(ms, s.Nb)
    assert relerr(om.apply(x, y), y - Mop @ x) < 1e-14
    # the generator keeps x_true the solution of (A - std wells - ms wells) x = b
    A = dense_from_bsr(s.rows, s.cols, s.vals) - dense_well_operator(s.wells, s.Nb) - Mop
    assert relerr(A @ s.x_true, s.b) < 1e-12
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, s.x_true, oracle_wells(s.wells), om) < 1e-13
    r = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200, mswells=om)
    assert r.converged and relerr(r.x, s.x_true) < 1e-5
    r0 = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200)
    assert relerr(r0.x, s.x_true) > 1e-3            # the wells matter


def test_multisegment_well_apply_against_the_reference_class():
    """Pins orc_ms_apply against the reference's own Opm::MultisegmentWellContribution::apply
    (bda/MultisegmentWellContribution.cpp:32-110), compiled unmodified into oracle/_ref/libref_mswell.so; only UMFPACK's five
    entry points are replaced (dense LU), so the block layouts and loops are the reference's."""
    from opm_autodiff_b200 import synth
    from tests.helpers import oracle_mswells, relerr
    s = synth.small(9, 8, 7)
    ms = synth.add_mswells(s, 4, 9, seed=17)
    rng = np.random.default_rng(4)
    x, y = rng.standard_normal(3 * s.Nb), rng.standard_normal(3 * s.Nb)
    om = oracle_mswells(ms)
    ref = y.copy()
    for w in om.wells:
        ref = oracle.ref_mswell_apply(w, x, ref)
        if ref is None:
            pytest.skip("oracle/_ref/libref_mswell.so was not built (needs /root/reference at build time)")
    got = om.apply(x, y)
    assert np.linalg.norm(ref - y) > 1e-3
    assert relerr(got - y, ref - y) < 1e-12
