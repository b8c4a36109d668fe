"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI vs the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import bridge_wells, oracle_wells, relerr
from tests.patterns import grid_pattern

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    from opm_autodiff_b200 import bridge, synth
    from oracle import oracle
    if not bridge.device_available():
        pytest.fail("GPU tests need a B200; the product has no CPU fallback")
    return bridge, synth, oracle


def _solve(bridge, s, tol=1e-10, maxit=200, wells=True, relaxation=1.0, opts=None):
    be = bridge.B200SolverBackend(0, maxit, tol, 0)
    be.set_option("relaxation", relaxation)
    for k, v in (opts or {}).items():
        be.set_option(k, v)
    res = bridge.BdaResult()
    wc = bridge_wells(s.wells if wells else None)
    st = be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc, res)
    x = np.zeros(3 * s.Nb)
    be.get_result(x)
    return be, st, res, x


def test_matr33_golden_through_bridge(mods, matr33):
    """The reference's own boundary test (tests/test_cusparseSolver.cpp:49-112): same call sequence,
    tol/maxiter from options_flexiblesolver.json (0.5 / 20), empty WellContributions; expected x is the
    CPU golden of tests/test_flexiblesolver.cpp:114-116 (the GPU-test goldens are not solutions)."""
    bridge, synth, oracle = mods
    g = matr33
    br = bridge.BdaBridge("b200", "empty", 10 * 0, 20, 0.5, 0, 0, "none")
    wc = bridge.WellContributions("b200", False)
    res = bridge.InverseOperatorResult()
    mat = bridge.BsrMatrix(g["rows"], g["cols"], g["vals"].copy())
    br.solve_system(mat, g["b"].copy(), wc, res)
    x = np.zeros(9)
    br.get_result(x)
    assert res.converged and br.last_result.it == 0.5
    assert np.max(np.abs(x / g["x_golden"] - 1.0)) < 1e-5           # BOOST_CHECK_CLOSE(.., 1e-3) percent
    ref = oracle.solve(g["rows"], g["cols"], g["vals"], g["b"], tol=0.5, maxit=20)
    assert relerr(x, ref.x) < 1e-9


@pytest.mark.parametrize("shape,faults", [((7, 6, 5), ()), ((12, 10, 8), ((6, 1),)), ((1, 1, 17), ()), ((33, 1, 1), ()),
                                           ((20, 16, 12), ((7, 2), (13, 1)))])
def test_kernels_vs_oracle(mods, shape, faults):
    bridge, synth, oracle = mods
    s = synth.small(*shape, faults=faults, nwells=2 if shape[0] >= 7 else 0, nperf=3)
    be = bridge.B200SolverBackend(0, 10, 1e-2, 0)
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells))
    rng = np.random.default_rng(11)
    x = rng.normal(size=3 * s.Nb) * np.tile([1e5, 1.0, 1.0], s.Nb)
    # SpMV: same products, different summation order inside a row -> a few ulp
    y = be.spmv(x)
    yo = oracle.spmv(s.rows, s.cols, s.vals, x)
    assert relerr(y, yo) < 1e-14
    # wells
    if s.wells is not None:
        y0 = rng.normal(size=3 * s.Nb)
        assert relerr(be.well_apply(x, y0), oracle.well_apply(oracle_wells(s.wells), x, y0)) < 1e-13
    # ILU0 factors in the caller's pattern
    assert be.ilu0_factorize() == bridge.SolverStatus.BDA_SOLVER_SUCCESS
    LU = be.get_ilu0(s.nnzb)
    LUo, diag, st = oracle.ilu0(s.rows, s.cols, s.vals)
    assert st == 0
    scale = np.abs(LUo).reshape(-1, 9).max(axis=1)[:, None, None]
    assert np.max(np.abs(LU - LUo) / scale) < 1e-9
    # ILU0 apply
    d = rng.normal(size=3 * s.Nb)
    v = be.ilu0_apply(d)
    vo = oracle.ilu0_apply(s.rows, s.cols, diag, LUo, d)
    assert relerr(v, vo) < 1e-9
    # relaxed apply v = w (LU)^-1 d: the back-substitution runs unrelaxed, then `v *= w` (ParallelOverlappingILU0.hpp:897-901)
    be.set_option("relaxation", 0.9)
    assert be.ilu0_factorize() == bridge.SolverStatus.BDA_SOLVER_SUCCESS
    v9 = be.ilu0_apply(d)
    assert relerr(v9, oracle.ilu0_apply(s.rows, s.cols, diag, LUo, d, w=0.9)) < 1e-9
    assert relerr(v9, 0.9 * vo) < 1e-9
    be.set_option("relaxation", 1.0)
    # level schedule of the uploaded pattern == oracle restatement of Reorder.cpp:266-318
    to, fr, rpl = be.get_level_schedule()
    oto, ofr, olp = oracle.level_schedule(s.rows, s.cols)
    assert np.array_equal(to, oto) and np.array_equal(fr, ofr) and np.array_equal(rpl, np.diff(olp))


@pytest.mark.parametrize("shape,faults,nwells,nperf", [((12, 10, 8), ((6, 1),), 3, 4), ((24, 20, 16), (), 0, 0),
                                                        ((30, 24, 20), ((10, 1), (20, 2)), 6, 15)])
@pytest.mark.parametrize("big", [0, 2])
def test_solve_parity_small(mods, shape, faults, nwells, nperf, big):
    """north_star parity: ||x - x_ref|| / ||x_ref|| <= 1e-6 at 1e-10 relative residual, iterations +-10%.
    big = 2 forces the features that are automatic from 100 000 rows (SELL SpMV, SpMV inside the upper sweep, deferred solution
    update, early helpers); big = 0 switches them off."""
    bridge, synth, oracle = mods
    s = synth.small(*shape, faults=faults, nwells=nwells, nperf=nperf)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200)
    feats = {k: big for k in ("spmv_sell", "fuse_spmv", "defer_x", "sweep_early")}
    be, st, res, x = _solve(bridge, s, opts=feats)
    assert st == bridge.SolverStatus.BDA_SOLVER_SUCCESS and res.converged and ref.converged
    assert relerr(x, ref.x) <= 1e-6
    assert abs(res.it - ref.it) <= max(1.0, 0.1 * ref.it)
    assert res.reduction < 1e-10
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, x, oracle_wells(s.wells)) < 2e-10
    assert res.iterations == int(res.it)                              # cusparseSolverBackend.cu:172
    # relaxation 0.9 (Flow's CPU default, FlowLinearSolverParameters.hpp:146-150): same solution
    ref9 = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200, relaxation=0.9)
    _, _, res9, x9 = _solve(bridge, s, relaxation=0.9, opts=feats)
    assert res9.converged and relerr(x9, ref9.x) <= 1e-6 and abs(res9.it - ref9.it) <= max(1.0, 0.1 * ref9.it)


def test_solve_parity_c2_norne_size(mods):
    """BASELINE.json configs[1]: Norne-sized 36x56x22 with NNC faults, full size (oracle: < 1 s)."""
    bridge, synth, oracle = mods
    s = synth.full_system("c2")
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, None, tol=1e-10, maxit=2000)
    be, st, res, x = _solve(bridge, s, maxit=2000)
    assert res.converged and relerr(x, ref.x) <= 1e-6
    assert abs(res.it - ref.it) <= 0.1 * ref.it
    # production setting (reduction 1e-2, maxit 200; FlowLinearSolverParameters.hpp:141-154)
    refp = oracle.solve(s.rows, s.cols, s.vals, s.b, None, tol=1e-2, maxit=200)
    _, _, resp, xp = _solve(bridge, s, tol=1e-2, maxit=200)
    assert resp.converged and resp.it == refp.it and relerr(xp, refp.x) < 1e-6


def test_full_size_c3_properties(mods):
    """BASELINE.json configs[2] at full size (1M cells, 50 wells x 20 perforations): size-independent
    checks -- true residual of the returned x (oracle operator), error against the generator's x_true,
    repeatability, and iteration count against the oracle's partial history."""
    bridge, synth, oracle = mods
    s = synth.full_system("c3")
    be, st, res, x = _solve(bridge, s, maxit=2000)
    assert res.converged and res.reduction < 1e-10
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, x, oracle_wells(s.wells)) < 2e-10
    assert relerr(x, s.x_true) < 1e-5
    res2 = bridge.BdaResult()
    be.solve_resident(res2)
    x2 = np.zeros_like(x)
    be.get_result(x2)
    assert res2.it == res.it and np.array_equal(x, x2)                # deterministic reductions
    # first 6 half steps of the oracle (about 2 s of CPU): same residual history to 1e-6 relative
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-30, maxit=3, threads=oracle.max_threads())
    be3 = bridge.B200SolverBackend(0, 3, 1e-30, 0)
    r3 = bridge.BdaResult()
    be3.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells), r3)
    assert not r3.converged and r3.it == 3.0
    assert abs(r3.norm / ref.norm - 1.0) < 1e-5 and abs(r3.norm0 / ref.norm0 - 1.0) < 1e-12


def test_solve_parity_c3_full_oracle_solve(mods):
    """north_star parity on the HEADLINE configuration: the full one-partition oracle solve of C3 (1 M cells, 50 wells x 20
    perforations; ~10 s on one host core) against the device solve: ||x - x_ref|| / ||x_ref|| <= 1e-6 at 1e-10 relative
    residual, iterations within +-10 %; the same at the production setting (reduction 1e-2, maxit 200:
    FlowLinearSolverParameters.hpp:141-154; FlexibleSolver_impl.hpp:145-157)."""
    bridge, synth, oracle = mods
    s = synth.full_system("c3")
    ow = oracle_wells(s.wells)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, ow, tol=1e-10, maxit=2000, threads=1)
    be, st, res, x = _solve(bridge, s, maxit=2000)
    assert ref.converged and res.converged and st == bridge.SolverStatus.BDA_SOLVER_SUCCESS
    assert relerr(x, ref.x) <= 1e-6
    assert abs(res.it - ref.it) <= 0.1 * ref.it
    refp = oracle.solve(s.rows, s.cols, s.vals, s.b, ow, tol=1e-2, maxit=200, threads=1)
    _, _, resp, xp = _solve(bridge, s, tol=1e-2, maxit=200)
    assert refp.converged and resp.converged and abs(resp.it - refp.it) <= max(0.5, 0.1 * refp.it)
    assert relerr(xp, refp.x) <= 1e-6


def test_nan_and_inf_input_end_without_convergence(mods):
    """A right-hand side or a matrix with NaN / Inf must end as converged = false (or CREATE_PRECONDITIONER_FAILED for a
    non-finite pivot), never in the dataflow time-out of the sweeps: ordinary NaNs are data, only the sentinel payload
    means "not computed yet"."""
    bridge, synth, oracle = mods
    s = synth.small(12, 10, 8, nwells=2, nperf=3)
    for what in ("rhs_nan", "rhs_inf", "offdiag_nan", "pivot_nan"):
        vals, b = s.vals.copy(), s.b.copy()
        if what == "rhs_nan": b[7] = np.nan
        if what == "rhs_inf": b[11] = np.inf
        if what == "offdiag_nan": vals.reshape(-1, 3, 3)[s.rows[5] + 1 if s.cols[s.rows[5]] == 5 else s.rows[5], 1, 2] = np.nan
        if what == "pivot_nan":
            k = [k for k in range(s.rows[9], s.rows[10]) if s.cols[k] == 9][0]
            vals.reshape(-1, 3, 3)[k] = np.nan
        be = bridge.B200SolverBackend(0, 50, 1e-10, 0)
        res = bridge.BdaResult()
        st = be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, vals, s.rows, s.cols, b, bridge_wells(s.wells), res)
        x = np.zeros(3 * s.Nb)
        be.get_result(x)
        if what in ("pivot_nan", "offdiag_nan"):       # a NaN in the matrix reaches a pivot: the factorisation reports it (the reference's
            # inverter only throws on det == 0, MatrixBlock.hpp:735, and goes on with a NaN preconditioner: no convergence either)
            assert st == bridge.SolverStatus.BDA_SOLVER_CREATE_PRECONDITIONER_FAILED or \
                (st == bridge.SolverStatus.BDA_SOLVER_SUCCESS and not res.converged), what
        else:
            assert st == bridge.SolverStatus.BDA_SOLVER_SUCCESS and not res.converged, what
        # and the solver object is still usable afterwards
        st = be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells), res)
        assert st == bridge.SolverStatus.BDA_SOLVER_SUCCESS and res.converged, what


def test_wells_edge_cases(mods):
    """> 10 perforations (the reference GPU kernels truncate there, WellContributions.cu:115-124), two
    wells sharing cells, B and C with different column lists, zero wells."""
    bridge, synth, oracle = mods
    s = synth.small(10, 8, 6)
    rng = np.random.default_rng(9)
    P = 23
    c0 = rng.permutation(s.Nb)[:P].astype(np.int32)
    c1 = np.concatenate([c0[:4], rng.permutation(s.Nb)[:5]]).astype(np.int32)
    cb = np.concatenate([c0, c1])
    cc = np.concatenate([c0[::-1], c1])                               # C columns differ from B columns
    sc = 0.05
    cs = np.array([1e-7, 1.0, 1.0])                                   # pressure column scaled like the matrix blocks
    from opm_autodiff_b200.synth import WellData
    w = WellData(np.array([0, P, P + len(c1)], np.uint32), cb, cc, sc * rng.normal(size=(len(cb), 4, 3)) * cs,
                 sc * rng.normal(size=(len(cb), 4, 3)), np.stack([np.eye(4) + 0.1 * rng.normal(size=(4, 4))] * 2))
    be = bridge.B200SolverBackend(0, 200, 1e-10, 0)
    wc = bridge_wells(w)
    assert wc.getNumWells() == 2
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
    x = rng.normal(size=3 * s.Nb)
    y0 = rng.normal(size=3 * s.Nb)
    assert relerr(be.well_apply(x, y0), oracle.well_apply(oracle_wells(w), x, y0)) < 1e-13
    res = bridge.BdaResult()
    be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc, res)
    xs = np.zeros(3 * s.Nb)
    be.get_result(xs)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(w), tol=1e-10, maxit=200)
    assert res.converged and relerr(xs, ref.x) <= 1e-6 and abs(res.it - ref.it) <= max(1.0, 0.1 * ref.it)
    # NULL wells and an empty container give the plain-matrix solve
    res0 = bridge.BdaResult()
    be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None, res0)
    x0 = np.zeros(3 * s.Nb)
    be.get_result(x0)
    ref0 = oracle.solve(s.rows, s.cols, s.vals, s.b, None, tol=1e-10, maxit=200)
    assert relerr(x0, ref0.x) <= 1e-6


def test_state_machine_and_errors(mods):
    """cusparseSolverBackend.cu:480-499: pattern analysed once, values/rhs change per call; dim != 3 and a
    changed pattern are errors; singular pivot -> CREATE_PRECONDITIONER_FAILED; maxit -> converged False."""
    bridge, synth, oracle = mods
    s = synth.small(9, 8, 7)
    be = bridge.B200SolverBackend(0, 200, 1e-10, 0)
    res = bridge.BdaResult()
    be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None, res)
    assert res.converged and res.t_analysis > 0 and res.num_levels == 9 + 8 + 7 - 2
    # second Newton step: new values and rhs, rows/cols not needed any more
    s2 = synth.small(9, 8, 7, seed=8)
    be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s2.vals, None, None, s2.b, None, res)
    x = np.zeros(3 * s.Nb)
    be.get_result(x)
    ref = oracle.solve(s2.rows, s2.cols, s2.vals, s2.b, None, tol=1e-10, maxit=200)
    assert res.converged and res.t_analysis == 0 and relerr(x, ref.x) <= 1e-6
    with pytest.raises(RuntimeError, match="pattern changed"):
        be.solve_system(3 * (s.Nb - 1), 9 * s.nnzb, 3, s.vals, None, None, s.b, None, res)
    with pytest.raises(RuntimeError, match="3x3"):
        bridge.B200SolverBackend(0, 10, 1e-2, 0).solve_system(2 * s.Nb, 4 * s.nnzb, 2, s.vals, s.rows, s.cols, s.b, None, res)
    # non-convergence
    be2 = bridge.B200SolverBackend(0, 2, 1e-14, 0)
    st = be2.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None, res)
    assert st == bridge.SolverStatus.BDA_SOLVER_SUCCESS and not res.converged and res.it == 2.0 and res.iterations == 2
    # singular pivot block
    bad = s.vals.copy()
    bad[0] = 0.0                      # row 0 has no lower entries: its pivot is never updated, so it stays singular
    assert s.cols[0] == 0 and oracle.ilu0(s.rows, s.cols, bad)[2] == 2
    be3 = bridge.B200SolverBackend(0, 5, 1e-2, 0)
    st = be3.solve_system(3 * s.Nb, 9 * s.nnzb, 3, bad, s.rows, s.cols, s.b, None, res)
    assert st == bridge.SolverStatus.BDA_SOLVER_CREATE_PRECONDITIONER_FAILED
    # missing diagonal -> analysis failure (BdaSolver.hpp:34)
    rows = np.array([0, 1, 2], np.int32); cols = np.array([0, 0], np.int32)
    be4 = bridge.B200SolverBackend(0, 5, 1e-2, 0)
    st = be4.solve_system(6, 18, 3, np.tile(np.eye(3), (2, 1, 1)), rows, cols, np.ones(6), None, res)
    assert st == bridge.SolverStatus.BDA_SOLVER_ANALYSIS_FAILED


def test_zero_rhs_and_single_row(mods):
    bridge, synth, oracle = mods
    s = synth.small(5, 4, 3)
    be, st, res, x = _solve(bridge, s)
    be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, None, None, np.zeros(3 * s.Nb), None, res)
    be.get_result(x)
    assert res.converged and res.it == 0.0 and not np.any(x)          # Dune: norm0 < 1e-30 -> 0 iterations
    one = bridge.B200SolverBackend(0, 5, 1e-10, 0)
    blk = np.array([[[2.0, 1, 0], [0, 3, 1], [1, 0, 4]]])
    r = bridge.BdaResult()
    one.solve_system(3, 9, 3, blk, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([1.0, 2, 3]), None, r)
    x1 = np.zeros(3)
    one.get_result(x1)
    assert r.converged and relerr(x1, np.linalg.solve(blk[0], [1.0, 2, 3])) < 1e-12


def test_bridge_zero_diagonal_fixup_feeds_backend(mods):
    """BdaBridge::checkZeroDiagonal (BdaBridge.cpp:125-161) mutates the caller's matrix before the solve."""
    bridge, synth, oracle = mods
    s = synth.small(6, 5, 4)
    vals = s.vals.copy()
    d = int(np.nonzero(s.cols[s.rows[7]:s.rows[8]] == 7)[0][0]) + s.rows[7]
    vals[d, 0, 0] = 0.0
    vo = vals.copy()
    oracle.check_zero_diagonal(s.rows, s.cols, vo)
    ref = oracle.solve(s.rows, s.cols, vo, s.b, None, tol=1e-10, maxit=200)
    br = bridge.BdaBridge("b200", "", 0, 200, 1e-10, 0, 0, "none")
    mat = bridge.BsrMatrix(s.rows, s.cols, vals)
    res = bridge.InverseOperatorResult()
    br.solve_system(mat, s.b, None, res)
    x = np.zeros(3 * s.Nb)
    br.get_result(x)
    assert mat.vals[d, 0, 0] == 1e-15
    assert res.converged == ref.converged and relerr(x, ref.x) <= 1e-6


def _irregular_system(synth, shape, extra_frac, seed, drop_upper=0.0):
    """A grid system with random long-range symmetric couplings added (NNC-like; rows get up to ~10 blocks, i.e. more than
    the three dependency slots of a sweep record and longer elimination plans), optionally made structurally non-symmetric."""
    s = synth.small(*shape)
    rng = np.random.default_rng(seed)
    Nb = s.Nb
    nb = [dict() for _ in range(Nb)]
    for i in range(Nb):
        for k in range(s.rows[i], s.rows[i + 1]):
            nb[i][int(s.cols[k])] = s.vals[k].copy()
    for _ in range(int(extra_frac * Nb)):
        a, b = (int(t) for t in rng.integers(0, Nb, 2))
        if a == b or b in nb[a]:
            continue
        t = 0.05 * rng.random()
        blk = -t * (np.eye(3) + 0.2 * rng.uniform(-1, 1, (3, 3))) * np.array([1e-7, 1.0, 1.0])[None, :]
        nb[a][b] = blk
        nb[b][a] = blk.T.copy() * np.array([1e-7, 1.0, 1.0])[None, :] / np.array([1e-7, 1.0, 1.0])[:, None]
        for r in (a, b):
            nb[r][r] = nb[r][r] + t * np.eye(3) * np.array([1e-7, 1.0, 1.0])[None, :]
    if drop_upper > 0.0:
        for i in range(Nb):
            for c in [c for c in nb[i] if c > i + 1 and rng.random() < drop_upper]:
                del nb[i][c]
    rows, cols, vals = [0], [], []
    for i in range(Nb):
        for c in sorted(nb[i]):
            cols.append(c); vals.append(nb[i][c])
        rows.append(len(cols))
    rows, cols, vals = np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals)
    return rows, cols, vals


@pytest.mark.parametrize("shape,extra,drop,opts", [((9, 8, 7), 0.15, 0.0, {}), ((14, 9, 6), 0.3, 0.0, {"sweep_parts": 23, "sweep_stage_bytes": 4096}),
                                                   ((10, 10, 5), 0.1, 0.3, {"sweep_parts": 7})])
def test_irregular_patterns_vs_oracle(mods, shape, extra, drop, opts):
    """Long rows (continuation records in the sweeps, long elimination plans), long-range dependencies (rows parked from
    other parts or from beyond the window) and a structurally non-symmetric pattern through the real kernels."""
    bridge, synth, oracle = mods
    rows, cols, vals = _irregular_system(synth, shape, extra, seed=3, drop_upper=drop)
    Nb = len(rows) - 1
    assert np.max(np.diff(rows)) > 7
    rng = np.random.default_rng(5)
    xt = rng.uniform(-1, 1, 3 * Nb) * np.tile([1e5, 1.0, 1.0], Nb)
    b = oracle.spmv(rows, cols, vals, xt)
    be = bridge.B200SolverBackend(0, 300, 1e-10, 0)
    for k, v in opts.items():
        be.set_option(k, v)
    be.upload_system(3 * Nb, 9 * len(cols), 3, vals, rows, cols, b, None)
    assert be.ilu0_factorize() == bridge.SolverStatus.BDA_SOLVER_SUCCESS
    LU = be.get_ilu0(len(cols))
    LUo, diag, st = oracle.ilu0(rows, cols, vals)
    assert st == 0
    scale = np.abs(LUo).reshape(-1, 9).max(axis=1)[:, None, None]
    assert np.max(np.abs(LU - LUo) / scale) < 1e-9
    d = rng.normal(size=3 * Nb)
    assert relerr(be.ilu0_apply(d), oracle.ilu0_apply(rows, cols, diag, LUo, d)) < 1e-9
    assert relerr(be.spmv(d), oracle.spmv(rows, cols, vals, d)) < 1e-14
    res = bridge.BdaResult()
    be.solve_resident(res)
    x = np.zeros(3 * Nb)
    be.get_result(x)
    ref = oracle.solve(rows, cols, vals, b, None, tol=1e-10, maxit=300)
    assert res.converged and ref.converged and relerr(x, ref.x) <= 1e-6 and abs(res.it - ref.it) <= max(1.0, 0.1 * ref.it)


@pytest.mark.parametrize("sell", [2, 0])
def test_spmv_layouts_with_very_long_rows(mods, sell):
    """SpMV from the sliced-ELL copy (default) and from the BSR arrays: a few rows coupled to ~40 cells exceed the width cap of
    their slice, so their tail comes from the BSR arrays (k_spmv_sell overflow path); Nb is not a multiple of 32."""
    bridge, synth, oracle = mods
    s = synth.small(11, 7, 5)
    rng = np.random.default_rng(17)
    Nb = s.Nb
    nb = [dict() for _ in range(Nb)]
    for i in range(Nb):
        for k in range(s.rows[i], s.rows[i + 1]):
            nb[i][int(s.cols[k])] = s.vals[k].copy()
    for hub in (3, 200, Nb - 1):
        for c in rng.choice(Nb, 40, replace=False):
            if int(c) != hub:
                nb[hub].setdefault(int(c), rng.uniform(-1, 1, (3, 3)) * 1e-3)
    rows, cols, vals = [0], [], []
    for i in range(Nb):
        for c in sorted(nb[i]):
            cols.append(c); vals.append(nb[i][c])
        rows.append(len(cols))
    rows, cols, vals = np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals)
    assert Nb % 32 != 0 and np.max(np.diff(rows)) > 40
    be = bridge.B200SolverBackend(0, 10, 1e-2, 0)
    be.set_option("spmv_sell", sell)
    be.upload_system(3 * Nb, 9 * len(cols), 3, vals, rows, cols, np.ones(3 * Nb), None)
    for seed in (1, 2):
        x = np.random.default_rng(seed).normal(size=3 * Nb) * np.tile([1e5, 1.0, 1.0], Nb)
        assert relerr(be.spmv(x), oracle.spmv(rows, cols, vals, x)) < 1e-14


@pytest.mark.parametrize("opts", [{}, {"sweep_parts": 5}, {"sweep_parts": 37, "sweep_stage_bytes": 8192, "fuse_unit_slices": 1}])
def test_spmv_inside_the_upper_sweep_matches_separate_kernels(mods, opts):
    """The upper-sweep CTAs run the following SpMV as their parts finish, the lower-sweep CTAs apply the deferred solution
    update, helpers fetch a stage ahead (all on by default from 100 000 rows, forced here).  Same solution and
    iteration count as with separate kernels, on a faulted grid with wells and on a pattern with rows of ~45 blocks (tail of a
    row read from the BSR arrays inside the fused kernel)."""
    bridge, synth, oracle = mods
    s = synth.small(20, 16, 12, faults=((7, 2),), nwells=3, nperf=5)
    out = {}
    FORCE = {"spmv_sell": 2, "defer_x": 2, "sweep_early": 2}        # small systems: the automatic setting would switch them off
    for fuse in (2, 0):
        _, st, res, x = _solve(bridge, s, tol=1e-10, maxit=300, opts=dict(opts, fuse_spmv=fuse, **(FORCE if fuse else {})))
        assert st == bridge.SolverStatus.BDA_SOLVER_SUCCESS and res.converged
        out[fuse] = (res.it, x)
    assert out[2][0] == out[0][0] and relerr(out[2][1], out[0][1]) < 1e-9
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=300)
    assert relerr(out[2][1], ref.x) <= 1e-6 and abs(out[2][0] - ref.it) <= max(1.0, 0.1 * ref.it)
    # long rows
    rows, cols, vals = _irregular_system(synth, (12, 9, 7), 0.2, seed=9)
    Nb = len(rows) - 1
    rng = np.random.default_rng(4)
    nb = {}
    hub_rows = {5: rng.choice(Nb, 40, replace=False), Nb - 3: rng.choice(Nb, 40, replace=False)}
    r2, c2, v2 = [0], [], []
    for i in range(Nb):
        ent = {int(cols[k]): vals[k] for k in range(rows[i], rows[i + 1])}
        for c in hub_rows.get(i, ()):
            ent.setdefault(int(c), rng.uniform(-1, 1, (3, 3)) * 1e-4 * np.array([1e-7, 1.0, 1.0])[None, :])
        for c in sorted(ent):
            c2.append(c); v2.append(ent[c])
        r2.append(len(c2))
    rows, cols, vals = np.array(r2, np.int32), np.array(c2, np.int32), np.array(v2)
    xt = rng.uniform(-1, 1, 3 * Nb) * np.tile([1e5, 1.0, 1.0], Nb)
    b = oracle.spmv(rows, cols, vals, xt)
    sol = {}
    for fuse in (2, 0):
        be = bridge.B200SolverBackend(0, 300, 1e-10, 0)
        for k, v in dict(opts, fuse_spmv=fuse, **(FORCE if fuse else {})).items():
            be.set_option(k, v)
        be.upload_system(3 * Nb, 9 * len(cols), 3, vals, rows, cols, b, None)
        res = bridge.BdaResult()
        be.solve_resident(res)
        x = np.zeros(3 * Nb); be.get_result(x)
        assert res.converged
        sol[fuse] = (res.it, x)
    assert sol[2][0] == sol[0][0] and relerr(sol[2][1], sol[0][1]) < 1e-9
    ref = oracle.solve(rows, cols, vals, b, None, tol=1e-10, maxit=300)
    assert relerr(sol[2][1], ref.x) <= 1e-6


def test_repeated_solves_reproduce_bit_for_bit(mods):
    """Race hunting without a sanitizer: repeated solves with every size-dependent feature forced on (dataflow sweeps, SpMV and
    solution updates in the sweep tails, per-unit dot partials) must reproduce the first solution bit for bit.
    tools/stress_determinism.py is the long version."""
    bridge, synth, oracle = mods
    for shape, nw, opts in (((33, 17, 29), 0, {"sweep_parts": 37}), ((25, 25, 25), 5, {"sweep_parts": 9, "fuse_unit_slices": 1})):
        s = synth.small(*shape, nwells=nw, nperf=6)
        be = bridge.B200SolverBackend(0, 300, 1e-10, 0)
        for k in ("spmv_sell", "fuse_spmv", "defer_x", "sweep_early"):
            be.set_option(k, 2)
        for k, v in opts.items():
            be.set_option(k, v)
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells))
        res = bridge.BdaResult()
        first = None
        for rep in range(8):
            be.solve_resident(res)
            x = np.zeros(3 * s.Nb)
            be.get_result(x)
            assert res.converged
            if first is None:
                first, it0 = x.copy(), res.it
            else:
                assert np.array_equal(x, first) and res.it == it0
        assert relerr(first, s.x_true) < 1e-5


def test_wells_flat_and_general_kernels_agree(mods):
    """Standard wells run through k_wells_flat when they fit one CTA (<= 1024 perforations, <= 128 wells) and through
    k_wells otherwise (option wells_flat = 0 forces the general kernel): both against the oracle, on wells that share
    cells and whose B and C columns differ, in the apply alone and inside a solve; and a container beyond the flat
    limits (1500 perforations) still takes the general kernel."""
    bridge, synth, oracle = mods
    from opm_autodiff_b200.synth import WellData
    s = synth.small(14, 12, 10)
    rng = np.random.default_rng(21)
    cs = np.array([1e-7, 1.0, 1.0])

    def wells(nw, nperf, shared):
        ptr, cb, cc = [0], [], []
        for w in range(nw):
            c = rng.permutation(s.Nb)[:nperf].astype(np.int32)
            if shared and w > 0:
                c[:3] = cb[-1][:3]                               # three cells shared with the previous well
            cb.append(c)
            cc.append(c[::-1].copy() if w % 2 else c.copy())     # C columns differ from B columns on odd wells
            ptr.append(ptr[-1] + nperf)
        n = ptr[-1]
        return WellData(np.array(ptr, np.uint32), np.concatenate(cb), np.concatenate(cc), 0.05 * rng.normal(size=(n, 4, 3)) * cs,
                        0.05 * rng.normal(size=(n, 4, 3)), np.stack([np.eye(4) + 0.1 * rng.normal(size=(4, 4)) for _ in range(nw)]))

    x = rng.normal(size=3 * s.Nb)
    y0 = rng.normal(size=3 * s.Nb)
    for nw, nperf, shared in ((1, 1, False), (7, 37, True), (40, 25, True), (3, 500, False), (130, 2, False)):
        w = wells(nw, nperf, shared)
        ref = oracle.well_apply(oracle_wells(w), x, y0)
        got = {}
        for flat in (1, 0):
            be = bridge.B200SolverBackend(0, 200, 1e-10, 0)
            be.set_option("wells_flat", flat)
            be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(w))
            got[flat] = be.well_apply(x, y0)
            assert relerr(got[flat] - y0, ref - y0) < 1e-12, (nw, nperf, flat)
            assert np.array_equal(be.well_apply(x, y0), got[flat])
        assert relerr(got[1], got[0]) < 1e-13
    # inside a solve: same iterates within rounding, same iteration count
    w = wells(6, 30, True)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(w), tol=1e-10, maxit=300)
    its = []
    for flat in (1, 0):
        be = bridge.B200SolverBackend(0, 300, 1e-10, 0)
        be.set_option("wells_flat", flat)
        res = bridge.BdaResult()
        be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(w), res)
        xs = np.zeros(3 * s.Nb)
        be.get_result(xs)
        assert res.converged and relerr(xs, ref.x) <= 1e-6 and abs(res.it - ref.it) <= max(1.0, 0.1 * ref.it)
        its.append(res.it)
    assert its[0] == its[1]


def test_well_values_change_under_a_cached_structure(mods):
    """WellContributions is rebuilt by the caller for every solve (ISTLSolverEbos.hpp:265-272) with new values on, usually,
    the same perforation pattern: the library keeps the index side on the device and refreshes the values only.  Same
    pattern / new values, then a new pattern, then the first one again, each against the oracle; both well kernels."""
    bridge, synth, oracle = mods
    import copy
    s = synth.small(12, 10, 8, nwells=4, nperf=5)
    rng = np.random.default_rng(77)
    x = rng.normal(size=3 * s.Nb)
    y0 = rng.normal(size=3 * s.Nb)
    w1 = s.wells
    w2 = copy.copy(w1)                                       # same pattern, other values
    w2.B = w1.B * (1.0 + 0.3 * rng.normal(size=w1.B.shape))
    w2.C = w1.C * (1.0 + 0.3 * rng.normal(size=w1.C.shape))
    w2.Dinv = w1.Dinv * (1.0 + 0.1 * rng.normal(size=w1.Dinv.shape))
    w3 = copy.copy(w2)                                       # other pattern (columns reversed per container), same sizes
    w3.Bcols = w2.Bcols[::-1].copy()
    w3.Ccols = w2.Ccols[::-1].copy()
    for flat in (1, 0):
        be = bridge.B200SolverBackend(0, 200, 1e-10, 0)
        be.set_option("wells_flat", flat)
        for w in (w1, w2, w3, w1, w2):
            be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(w))
            ref = oracle.well_apply(oracle_wells(w), x, y0)
            assert relerr(be.well_apply(x, y0) - y0, ref - y0) < 1e-12
        res = bridge.BdaResult()
        for w in (w2, w3):
            be.solve_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(w), res)
            xs = np.zeros(3 * s.Nb)
            be.get_result(xs)
            ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(w), tol=1e-10, maxit=200)
            assert res.converged == ref.converged and relerr(xs, ref.x) <= 1e-6


@pytest.mark.parametrize("shape,faults,nwells,nperf,sell", [((12, 10, 8), ((6, 1),), 3, 4, 0), ((30, 24, 20), ((10, 1),), 4, 6, 2)])
def test_opt_in_graph_colouring_matches_the_oracle_on_the_permuted_system(mods, shape, faults, nwells, nperf, sell):
    """SURVEY a19, opt-in: BdaBridge(..., opencl_ilu_reorder="graph_coloring") (BdaBridge.cpp:72-80, BILU0.cpp:86-91) takes ILU0
    of the colour-permuted matrix -- another preconditioner, so the parity target is the oracle's solve of P A P^T with the
    ordering the solver reports (b200_get_reorder): x to 1e-6 in the caller's order, half-step iterations within 10 %, and
    the colouring needs MORE iterations than natural order (why it is not the default).  vals / b / x stay in the caller's
    order; the sweeps run level by level (one launch per colour)."""
    bridge, synth, oracle = mods
    from tests.helpers import permute_system
    s = synth.small(*shape, faults=faults, nwells=nwells, nperf=nperf)
    br = bridge.BdaBridge("b200", "", 0, 400, 1e-10, 0, 0, "graph_coloring")
    br.backend.set_option("spmv_sell", sell)
    res = bridge.InverseOperatorResult()
    wc = bridge_wells(s.wells)
    br.solve_system(bridge.BsrMatrix(s.rows, s.cols, s.vals.copy()), s.b, wc, res)
    x = np.zeros(3 * s.Nb)
    br.get_result(x)
    to, fr, nc = br.backend.get_reorder()
    assert 2 <= nc <= 256 and np.array_equal(fr[to], np.arange(s.Nb))
    prow, pcol, pval, pb, pw = permute_system(s.rows, s.cols, s.vals, s.b, to, fr, s.wells)
    ref = oracle.solve(prow, pcol, pval, pb, oracle_wells(pw), tol=1e-10, maxit=400)
    xr = np.zeros(3 * s.Nb).reshape(-1, 3)
    xr[fr] = ref.x.reshape(-1, 3)
    assert res.converged and ref.converged
    assert relerr(x, xr.reshape(-1)) <= 1e-6
    nat = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=400)
    assert abs(res.iterations - int(ref.it)) <= max(1, 0.1 * ref.it) and ref.it >= nat.it
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, x, oracle_wells(s.wells)) < 2e-10
    # the plain ILU0 apply in that ordering: v = (LU)^-1 d of the permuted matrix, in the caller's order
    d = np.random.default_rng(5).normal(size=3 * s.Nb)
    v = br.backend.ilu0_apply(d)
    LUo, diag, st = oracle.ilu0(prow, pcol, pval)
    assert st == 0
    vr = oracle.ilu0_apply(prow, pcol, diag, LUo, d.reshape(-1, 3)[fr].reshape(-1))
    vv = np.zeros_like(xr)
    vv[fr] = np.asarray(vr).reshape(-1, 3)
    assert relerr(v, vv.reshape(-1)) < 1e-9
