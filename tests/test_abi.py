"""CPU tests: the C-ABI library loads, exports every symbol include/b200bda.h declares, and its
host-only logic (well container, level scheduling, error paths) behaves like the reference."""
import ctypes
import glob
import json
import os
import re

import numpy as np
import pytest

from tests.patterns import grid_pattern

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from opm_autodiff_b200 import bridge
    with open(os.path.join(ROOT, "include", "b200bda.h")) as f:
        hdr = f.read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(bridge.LIB_PATH)
    for sym in declared:
        assert hasattr(L, sym), "missing export %s" % sym
    assert declared == set(bridge.EXPORTED_SYMBOLS)


def test_no_cpu_fallback_without_device(built):
    from opm_autodiff_b200 import bridge
    if bridge.device_available():
        pytest.skip("a B200 is present")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        bridge.B200SolverBackend(0, 10, 1e-2, 0)
    with pytest.raises(RuntimeError):
        bridge.BdaBridge("b200", "", 0, 10, 1e-2, 0, 0, "none")


def test_product_does_not_touch_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "opm-autodiff_b200")
    for path in glob.glob(os.path.join(pkg, "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".c", ".cpp", "Makefile")):
            with open(path, errors="ignore") as f:
                txt = f.read()
            for needle in ("import oracle", "from oracle", "oracle/", "oracle.py", "liboracle", "orc_"):
                assert needle not in txt, "%s references the oracle (%r)" % (path, needle)


def test_well_contributions_error_behaviour(built):
    """WellContributions.cpp:31-49,152-259."""
    from opm_autodiff_b200.bridge import WellContributions as WC
    with pytest.raises(ValueError, match="Invalid accelerator mode"):
        WC("quantum", False)
    with pytest.raises(ValueError, match="amgcl requires"):
        WC("amgcl", False)
    w = WC("b200", False)
    assert w.getNumWells() == 0
    with pytest.raises(ValueError, match="must be equal to 3 and 4"):
        w.setBlockSize(3, 3)
    w.setBlockSize(3, 4)
    with pytest.raises(ValueError, match="before allocating"):
        w.addMatrix(WC.MatrixType.C, np.zeros(1, np.int32), np.zeros(12), 1)
    w.addNumBlocks(2)
    w.addNumBlocks(1)
    w.alloc()
    assert w.getNumWells() == 2
    with pytest.raises(ValueError, match="after allocated"):
        w.addNumBlocks(1)
    w.addMatrix(WC.MatrixType.C, np.array([0, 1], np.int32), np.zeros(24), 2)
    w.addMatrix(WC.MatrixType.D, None, np.eye(4).reshape(-1), 1)
    w.addMatrix(WC.MatrixType.B, np.array([0, 1], np.int32), np.zeros(24), 2)
    with pytest.raises(ValueError, match="more blocks"):
        w.addMatrix(WC.MatrixType.C, np.array([0, 1], np.int32), np.zeros(24), 2)
    empty = WC("b200", False)
    empty.alloc()                      # zero wells: nothing allocated, no error (:236-259)
    assert empty.getNumWells() == 0


def test_multisegment_well_container(built):
    """addMultisegmentWellContribution (WellContributions.hpp:195-213, .cpp:261-271): host-side behaviour -- counts, block-size
    check, and the dense inverse of the CSC matrix D that stands in for the reference's UMFPACK factorisation
    (MultisegmentWellContribution.cpp:56-57)."""
    from opm_autodiff_b200 import synth
    from opm_autodiff_b200.bridge import WellContributions as WC
    from tests.helpers import add_bridge_mswells
    s = synth.small(6, 5, 4)
    ms = synth.add_mswells(s, 3, 7, seed=11)
    w = add_bridge_mswells(WC("b200", False), ms)
    assert w.getNumWells() == 3                      # num_std_wells + num_ms_wells (WellContributions.hpp:164-166)
    for i, m in enumerate(ms):
        D = m.dense_D()
        inv = w.multisegment_inverse(i, 4 * m.Mb)
        assert np.abs(inv @ D - np.eye(4 * m.Mb)).max() < 1e-12
    m = ms[0]
    with pytest.raises(ValueError, match="must be equal to 3 and 4"):
        w.addMultisegmentWellContribution(3, 3, m.Mb, m.Bvalues, m.BcolIndices, m.BrowPointers, m.DnumBlocks, m.Dvalues,
                                          m.DcolPointers, m.DrowIndices, m.Cvalues)
    with pytest.raises(ValueError, match="singular"):
        w.addMultisegmentWellContribution(3, 4, m.Mb, m.Bvalues, m.BcolIndices, m.BrowPointers, m.DnumBlocks,
                                          np.zeros_like(m.Dvalues), m.DcolPointers, m.DrowIndices, m.Cvalues)
    with pytest.raises(ValueError, match="array sizes"):
        w.addMultisegmentWellContribution(3, 4, m.Mb + 1, m.Bvalues, m.BcolIndices, m.BrowPointers, m.DnumBlocks, m.Dvalues,
                                          m.DcolPointers, m.DrowIndices, m.Cvalues)
    assert w.getNumWells() == 3
    # a pivoting case: D with a zero on the diagonal is still inverted
    P = WC("b200", False)
    Dd = np.array([[0.0, 2.0, 0, 0], [1.0, 0.0, 0, 0], [0, 0, 0.0, 4.0], [0, 0, 3.0, 1.0]])
    colptr = np.arange(0, 17, 4, dtype=np.int32)
    rowidx = np.tile(np.arange(4, dtype=np.int32), 4)
    P.addMultisegmentWellContribution(3, 4, 1, np.zeros(12), np.zeros(1, np.uint32), np.array([0, 1], np.uint32), 1,
                                      Dd.T.reshape(-1).copy(), colptr, rowidx, np.zeros(12))
    assert np.abs(P.multisegment_inverse(0, 4) @ Dd - np.eye(4)).max() < 1e-14


def test_bridge_mode_strings_and_zero_diagonal(built):
    from opm_autodiff_b200 import bridge
    with pytest.raises(ValueError, match="AcceleratorMode"):
        bridge.BdaBridge("cuda", "", 0, 10, 1e-2, 0, 0, "none")
    br = bridge.BdaBridge("none", "", 0, 10, 1e-2, 0, 0, "none")
    assert not br.getUseGpu()
    rows, cols = grid_pattern(3, 3, 1)
    vals = np.tile(np.eye(3), (rows[-1], 1, 1))
    m = bridge.BsrMatrix(rows, cols, vals)
    d0 = int(np.nonzero(cols[rows[4]:rows[5]] == 4)[0][0]) + rows[4]
    m.vals[d0, 2, 2] = 0.0
    m.vals[d0 + 1, 0, 0] = 0.0          # an off-diagonal BLOCK is never touched
    assert br.checkZeroDiagonal(m) == 1
    assert m.vals[d0, 2, 2] == 1e-15 and m.vals[d0 + 1, 0, 0] == 0.0
    assert br.checkZeroDiagonal(m) == 0
    res = bridge.InverseOperatorResult(converged=True)
    br.solve_system(m, np.zeros(27), None, res)
    assert res.converged is False       # BdaBridge.cpp:252-254: no accelerator -> converged = false
    # the oracle's restatement of the same fix-up agrees
    from oracle import oracle
    v2 = np.tile(np.eye(3), (rows[-1], 1, 1)); v2[d0, 2, 2] = 0.0
    assert oracle.check_zero_diagonal(rows, cols, v2) == 1 and v2[d0, 2, 2] == 1e-15


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "levels_*.json"))))
def test_product_level_schedule_matches_reference_fixture(fixture, built):
    """b200_level_schedule_host (host-only) reproduces the compiled reference's toOrder/fromOrder/levels."""
    from opm_autodiff_b200 import bridge
    with open(fixture) as f:
        g = json.load(f)
    to, fr, rpl = bridge.level_schedule_host(g["rows"], g["cols"])
    assert to.tolist() == g["toOrder"] and fr.tolist() == g["fromOrder"] and rpl.tolist() == g["rowsPerColor"]


def test_product_level_schedule_vs_oracle_and_live_reference(built):
    from opm_autodiff_b200 import bridge
    from oracle import oracle
    for (nx, ny, nz, nnc) in ((9, 7, 5, 0), (10, 6, 7, 2), (1, 1, 5, 0), (2, 1, 1, 0), (16, 16, 1, 0)):
        rows, cols = grid_pattern(nx, ny, nz, nnc)
        to, fr, rpl = bridge.level_schedule_host(rows, cols)
        oto, ofr, olp = oracle.level_schedule(rows, cols)
        assert np.array_equal(to, oto) and np.array_equal(fr, ofr) and np.array_equal(rpl, np.diff(olp))
        ref = oracle.ref_level_schedule(rows, cols)
        if ref is not None:
            assert np.array_equal(to, ref[0]) and np.array_equal(fr, ref[1]) and np.array_equal(rpl, np.diff(ref[2]))
    # missing diagonal / bad column -> analysis failure, not a crash
    rows = np.array([0, 1, 2], np.int32)
    with pytest.raises(RuntimeError):
        bridge.level_schedule_host(rows, np.array([0, 5], np.int32))


def test_synth_slabs_equal_full_system(built):
    from opm_autodiff_b200 import synth
    import dataclasses
    cfg = synth.GridConfig("t", 7, 6, 8, seed=3, faults=((3, 2),), nwells=4, nperf=3)
    full = synth.generate(cfg)
    off = 0
    for (k0, k1) in ((0, 3), (3, 8)):
        s = synth.generate(cfg, k0, k1)
        n = s.Nb
        r0 = full.rows[off]
        assert np.array_equal(s.rows, full.rows[off:off + n + 1] - r0)
        assert np.array_equal(s.cols, full.cols[r0:full.rows[off + n]])
        assert np.array_equal(s.vals, full.vals[r0:full.rows[off + n]])
        assert np.allclose(s.b, full.b[3 * off:3 * (off + n)], rtol=0, atol=0)
        off += n
    again = synth.generate(cfg)
    assert np.array_equal(again.vals, full.vals) and np.array_equal(again.b, full.b)
    c3 = synth.CONFIGS["c3"]
    assert c3.ncells == 1_000_000 and c3.nwells == 50
    c2 = synth.generate(synth.CONFIGS["c2"])
    assert c2.Nb == 44352 and c2.nnzb >= 302384        # + NNC blocks


def test_multisegment_inverse_on_hard_patterns(built):
    """invert_csc (extent-aware LU with partial pivoting) against numpy on matrices that need pivoting and whose pattern is
    not banded: a segment tree with shuffled numbering, weak diagonals, an arrow matrix, duplicates in the CSC input
    (UMFPACK sums them)."""
    from opm_autodiff_b200.bridge import WellContributions as WC
    rng = np.random.default_rng(123)

    def check(D, dup=False):
        M = D.shape[0]
        assert M % 4 == 0
        colptr, rowidx, vals = [0], [], []
        for c in range(M):
            nz = np.nonzero(D[:, c])[0]
            for r in nz:
                if dup:                                   # split the entry in two
                    rowidx += [int(r), int(r)]; vals += [0.25 * D[r, c], 0.75 * D[r, c]]
                else:
                    rowidx.append(int(r)); vals.append(D[r, c])
            colptr.append(len(rowidx))
        while len(vals) % 16:                             # DnumBlocks * 16 entries: pad with explicit zeros in the last column
            rowidx.append(M - 1); vals.append(0.0); colptr[-1] += 1
        w = WC("b200", False)
        Mb = M // 4
        w.addMultisegmentWellContribution(3, 4, Mb, np.zeros(12), np.zeros(1, np.uint32), np.array([0] + [1] * Mb, np.uint32),
                                          len(vals) // 16, np.array(vals), np.array(colptr, np.int32), np.array(rowidx, np.int32),
                                          np.zeros(12))
        inv = w.multisegment_inverse(0, M)
        ref = np.linalg.inv(D)
        assert np.abs(inv - ref).max() <= 1e-10 * max(1.0, np.abs(ref).max())

    # tree with shuffled segment numbering, weak (sometimes zero) diagonal entries
    nseg = 12
    order = rng.permutation(nseg)
    D = np.zeros((4 * nseg, 4 * nseg))
    for s_ in range(nseg):
        a = order[s_]
        D[4 * a:4 * a + 4, 4 * a:4 * a + 4] = rng.normal(size=(4, 4))
        D[4 * a, 4 * a] = 0.0                              # forces row exchanges
        if s_ > 0:
            q = order[rng.integers(0, s_)]
            D[4 * a:4 * a + 4, 4 * q:4 * q + 4] = rng.normal(size=(4, 4))
            D[4 * q:4 * q + 4, 4 * a:4 * a + 4] = rng.normal(size=(4, 4))
    check(D)
    check(D, dup=True)
    # arrow: dense first block row and column
    M = 32
    A = np.diag(2.0 + rng.random(M))
    A[:4, :] = rng.normal(size=(4, M)); A[:, :4] = rng.normal(size=(M, 4))
    check(A)
    # fully dense 8 x 8
    check(rng.normal(size=(8, 8)))
