"""CPU checks against the reference's OWN reorder code (bda/Reorder.cpp + bda/BlockedMatrix.cpp, compiled unmodified into
oracle/_ref/libref_reorder.so):

* blockMult / blockMultSub (BlockedMatrix.cpp:69-100) carry a left-looking block ILU0 written as the OpenCL kernels state it
  (L_ij = A_ij invD_j, A_ik -= L_ij U_jk, invD_i = inv(A_ii)) -- must equal the oracle's factorisation;
* reorderBlockedMatrixByPattern (Reorder.cpp:179-208) with the level-scheduling order: ILU0-BiCGSTAB of P A P^T is the
  natural-order solve (same iterates) -- the property the device schedule relies on (any topological order gives the
  sequential result);
* findGraphColoring (Reorder.cpp:322-330): a valid colouring, but a different preconditioner -- more iterations than the
  +-10 % bound of BASELINE.json allows, which is why the backend does not colour (ILUReorder.hpp:25-26)."""
import numpy as np
import pytest

from oracle import oracle
from tests.helpers import relerr


@pytest.fixture(scope="module")
def R():
    lib = oracle.ref_lib()
    if lib is None:
        pytest.skip("oracle/_ref/libref_reorder.so was not built (needs /root/reference at build time)")
    return lib


def _reorder(R, rows, cols, vals, to, fr):
    Nb, nnzb = len(rows) - 1, len(cols)
    rv = np.zeros(nnzb * 9)
    rc = np.zeros(nnzb, np.int32)
    rr = np.zeros(Nb + 1, np.int32)
    R.ref_reorder_matrix(Nb, nnzb, np.ascontiguousarray(vals, np.float64).reshape(-1).copy(), np.ascontiguousarray(cols, np.int32).copy(),
                         np.ascontiguousarray(rows, np.int32).copy(), np.ascontiguousarray(to, np.int32), np.ascontiguousarray(fr, np.int32),
                         rv, rc, rr)
    return rr, rc, rv.reshape(-1, 3, 3)


def test_ilu0_built_from_the_reference_block_routines(R):
    from opm_autodiff_b200 import synth
    s = synth.small(5, 4, 3, faults=((2, 1),))
    rows, cols = s.rows, s.cols
    Nb = s.Nb
    LU = np.array(s.vals, dtype=np.float64).reshape(-1, 9).copy()
    pos = [{int(cols[k]): k for k in range(rows[i], rows[i + 1])} for i in range(Nb)]
    invD = np.zeros((Nb, 9))
    tmp = np.zeros(9)
    for i in range(Nb):
        for k in range(rows[i], rows[i + 1]):
            j = int(cols[k])
            if j >= i:
                break
            R.ref_block_mult(LU[k].copy(), invD[j], tmp)            # L_ij = A_ij * invD_j
            LU[k] = tmp
            for kk, q in pos[j].items():                            # A_ik -= L_ij * U_jk for k > j present in row i
                if kk > j and kk in pos[i]:
                    R.ref_block_mult_sub(LU[pos[i][kk]], LU[k], LU[q])
        inv = np.zeros(9)
        assert oracle.lib().orc_inv3(LU[pos[i][i]], inv) == 0
        invD[i] = inv
        LU[pos[i][i]] = inv                                          # inverse pivot stored in the diagonal slot
    ref, diag, st = oracle.ilu0(rows, cols, s.vals)
    assert st == 0
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True) + 1e-300
    assert np.max(np.abs(LU.reshape(-1, 3, 3) - ref) / scale) < 1e-12


def test_level_order_permutation_keeps_the_solve(R):
    from opm_autodiff_b200 import synth
    s = synth.small(10, 8, 6, faults=((5, 1),))
    to, fr, lp = oracle.ref_level_schedule(s.rows, s.cols)
    rr, rc, rv = _reorder(R, s.rows, s.cols, s.vals, to, fr)
    rb = s.b.reshape(-1, 3)[fr].reshape(-1)                         # reorderBlockedVectorByPattern (Reorder.cpp:231-237)
    nat = oracle.solve(s.rows, s.cols, s.vals, s.b, tol=1e-10, maxit=200)
    per = oracle.solve(rr, rc, rv, rb, tol=1e-10, maxit=200)
    x = np.zeros_like(per.x).reshape(-1, 3)
    x[fr] = per.x.reshape(-1, 3)
    assert per.it == nat.it
    assert relerr(x.reshape(-1), nat.x) < 1e-9
    assert relerr(per.history, nat.history) < 1e-6                  # same iterates, not just the same answer


def test_graph_colouring_is_valid_but_a_weaker_preconditioner(R):
    from opm_autodiff_b200 import synth
    s = synth.small(16, 14, 12)
    Nb = s.Nb
    to, fr, rpc = np.zeros(Nb, np.int32), np.zeros(Nb, np.int32), np.zeros(Nb, np.int32)
    nc = R.ref_graph_coloring(Nb, np.ascontiguousarray(s.rows, np.int32).copy(), np.ascontiguousarray(s.cols, np.int32).copy(), to, fr, rpc)
    assert 2 <= nc <= 256 and rpc[:nc].sum() == Nb                  # MAX_COLORS (Reorder.hpp:30)
    colour = np.repeat(np.arange(nc), rpc[:nc])[to]                 # colour of natural row i
    for i in range(Nb):
        for k in range(s.rows[i], s.rows[i + 1]):
            assert s.cols[k] == i or colour[s.cols[k]] != colour[i]
    rr, rc, rv = _reorder(R, s.rows, s.cols, s.vals, to, fr)
    rb = s.b.reshape(-1, 3)[fr].reshape(-1)
    nat = oracle.solve(s.rows, s.cols, s.vals, s.b, tol=1e-10, maxit=400)
    col = oracle.solve(rr, rc, rv, rb, tol=1e-10, maxit=400)
    x = np.zeros_like(col.x).reshape(-1, 3)
    x[fr] = col.x.reshape(-1, 3)
    assert col.converged and relerr(x.reshape(-1), nat.x) < 1e-5    # same system, same answer (two 1e-10 residual solves) ...
    assert col.it > 1.1 * nat.it                                    # ... but outside the +-10 % iteration bound


def test_own_graph_colouring_against_the_compiled_reference(R):
    """coloring.hpp (the opt-in ordering of option reorder = 1) restates colorBlockedNodes + colorsToReordering with a seed
    instead of std::random_device: valid (connected rows differ), colour-major with natural order inside a colour, the
    reference's per-colour limit (scalar rows per colour < Nb, BILU0.cpp:89), reproducible, and as many colours as the
    compiled reference needs on the same grid, give or take the randomness of either."""
    from opm_autodiff_b200 import bridge, synth
    s = synth.small(16, 14, 12, faults=((7, 1),))
    Nb = s.Nb
    to, fr, rpc = bridge.graph_coloring_host(s.rows, s.cols, seed=3)
    nc = len(rpc)
    assert 2 <= nc <= 256 and rpc.sum() == Nb
    assert np.array_equal(np.sort(to), np.arange(Nb)) and np.array_equal(fr[to], np.arange(Nb))
    colour = np.repeat(np.arange(nc), rpc)[to]
    for i in range(Nb):
        for k in range(s.rows[i], s.rows[i + 1]):
            assert s.cols[k] == i or colour[s.cols[k]] != colour[i]
    for c in range(nc):                                               # natural order inside a colour
        members = fr[rpc[:c].sum():rpc[:c + 1].sum()]
        assert np.all(np.diff(members) > 0)
    assert 3 * rpc.max() + 2 <= Nb + 3                                # rowsInColor + block_size - 1 >= maxRowsPerColor stops a colour
    to2, fr2, rpc2 = bridge.graph_coloring_host(s.rows, s.cols, seed=3)
    assert np.array_equal(to, to2) and np.array_equal(rpc, rpc2)      # seeded: reproducible
    to3, _, _ = bridge.graph_coloring_host(s.rows, s.cols, seed=4)
    assert not np.array_equal(to, to3)
    rto, rfr, rrpc = np.zeros(Nb, np.int32), np.zeros(Nb, np.int32), np.zeros(Nb, np.int32)
    rnc = R.ref_graph_coloring(Nb, np.ascontiguousarray(s.rows, np.int32).copy(), np.ascontiguousarray(s.cols, np.int32).copy(), rto, rfr, rrpc)
    assert 0.6 * rnc <= nc <= 1.6 * rnc


def test_own_colouring_gives_the_reference_permuted_solve(R):
    """The permutation helpers used by the GPU test agree with the compiled reorderBlockedMatrixByPattern, and the oracle's
    solve of the colour-permuted system -- the parity target of the opt-in GPU path -- needs more iterations than natural
    order but converges to the same solution."""
    from opm_autodiff_b200 import bridge, synth
    from tests.helpers import permute_system
    s = synth.small(12, 10, 8)
    to, fr, rpc = bridge.graph_coloring_host(s.rows, s.cols, seed=1)
    prow, pcol, pval, pb, _ = permute_system(s.rows, s.cols, s.vals, s.b, to, fr)
    rr, rc, rv = _reorder(R, s.rows, s.cols, s.vals, to, fr)
    assert np.array_equal(prow, rr) and np.array_equal(pcol, rc) and np.array_equal(pval.reshape(-1), np.asarray(rv).reshape(-1))
    nat = oracle.solve(s.rows, s.cols, s.vals, s.b, tol=1e-10, maxit=400)
    col = oracle.solve(prow, pcol, pval, pb, tol=1e-10, maxit=400)
    x = np.zeros_like(col.x).reshape(-1, 3)
    x[fr] = col.x.reshape(-1, 3)
    assert col.converged and relerr(x.reshape(-1), nat.x) < 1e-5 and col.it >= nat.it
