"""Host-side checks of the pencil schedule of the triangular sweeps (no GPU).

The emulator (analysis.hpp emulate_sweep) walks the packed streams exactly as the device kernel does --
stage by stage, chunk by chunk, shared-memory window with wrap-around, NaN-sentinel waits on other parts --
and is compared with the sequential natural-order forward/backward substitution the reference performs
(ParallelOverlappingILU0.hpp:867-895).  A schedule that could deadlock on the device fails here.
"""
import numpy as np
import pytest

from tests.patterns import grid_pattern


@pytest.mark.parametrize("shape,parts,stage_bytes,window", [
    ((1, 1, 1), 148, 0, 0),
    ((3, 1, 1), 148, 0, 0),
    ((40, 1, 1), 4, 0, 0),
    ((12, 10, 8), 148, 0, 0),
    ((12, 10, 8), 7, 2048, 64),
    ((20, 18, 16), 148, 0, 0),
    ((20, 18, 16), 37, 4096, 256),
    ((30, 30, 1), 16, 0, 64),
    ((1, 25, 25), 148, 0, 0),
    # the stage sizes the automatic rule picks (a quarter of a part's factor bytes, 16-80 KB): Norne size and 110 k rows
    ((36, 56, 22), 148, 16384, 0),
    ((36, 56, 22), 148, 20480, 0),
    ((48, 48, 48), 148, 57344, 0),
])
def test_emulated_sweeps_match_sequential_substitution(built, shape, parts, stage_bytes, window):
    from opm_autodiff_b200 import bridge
    rows, cols = grid_pattern(*shape)
    err, st = bridge.sweep_schedule_check_host(rows, cols, parts, stage_bytes, window, seed=3)
    assert err < 1e-11
    assert 1 <= st["parts"] <= max(1, parts)
    assert st["levels"] == sum(shape) - 2


def test_fault_pattern_and_random_extra_connections(built):
    from opm_autodiff_b200 import bridge
    rows, cols = grid_pattern(14, 9, 11, nnc_planes=2)
    err, st = bridge.sweep_schedule_check_host(rows, cols, 40, 0, 128, seed=5)
    assert err < 1e-11
    # random long-range symmetric couplings (NNC-like): still a valid, deadlock-free schedule
    Nb = len(rows) - 1
    rng = np.random.default_rng(7)
    nb = [set(cols[rows[i]:rows[i + 1]]) for i in range(Nb)]
    for _ in range(Nb // 10):
        a, b = rng.integers(0, Nb, 2)
        nb[a].add(int(b)); nb[b].add(int(a))
    r2 = np.zeros(Nb + 1, np.int32)
    c2 = []
    for i in range(Nb):
        c2.extend(sorted(nb[i]))
        r2[i + 1] = len(c2)
    err, st = bridge.sweep_schedule_check_host(r2, np.array(c2, np.int32), 23, 3000, 64, seed=9)
    assert err < 1e-10


def test_structurally_nonsymmetric_pattern(built):
    from opm_autodiff_b200 import bridge
    # drop some upper entries: the symmetrised levels must still order both sweeps
    rows, cols = grid_pattern(9, 8, 7)
    Nb = len(rows) - 1
    rng = np.random.default_rng(11)
    r2 = np.zeros(Nb + 1, np.int32)
    c2 = []
    for i in range(Nb):
        for c in cols[rows[i]:rows[i + 1]]:
            if c > i and rng.random() < 0.3:
                continue
            c2.append(int(c))
        r2[i + 1] = len(c2)
    err, st = bridge.sweep_schedule_check_host(r2, np.array(c2, np.int32), 19, 0, 64, seed=13)
    assert err < 1e-11


def test_pencil_partition_of_a_structured_grid(built):
    from opm_autodiff_b200 import bridge
    rows, cols = grid_pattern(24, 24, 24)
    err, st = bridge.sweep_schedule_check_host(rows, cols, 36, 0, 0, seed=1)
    assert err < 1e-11
    assert st["lines"] == 24 * 24 and st["parts"] == 36
    # pencils: the bulk of the dependencies stays inside a part (shared memory)
    assert st["window_deps_L"] > 3 * st["global_deps_L"]


@pytest.mark.parametrize("shape,faults", [((7, 6, 5), ()), ((12, 10, 8), ((6, 1),)), ((1, 1, 17), ()), ((20, 16, 3), ((7, 2), (13, 1)))])
def test_factor_plan_replay_matches_oracle_ilu0(built, shape, faults):
    """The elimination plan k_ilu_factor_plan executes, replayed on the host, against the oracle's left-looking
    block ILU0 (ParallelOverlappingILU0.hpp:440-494)."""
    from opm_autodiff_b200 import bridge, synth
    from oracle import oracle
    s = synth.small(*shape, faults=faults)
    lu, max_row, max_ops = bridge.factor_plan_check_host(s.rows, s.cols, s.vals)
    ref, diag, st = oracle.ilu0(s.rows, s.cols, s.vals)
    assert st == 0
    scale = np.abs(ref).max(axis=(1, 2), keepdims=True) + 1e-300
    assert np.max(np.abs(lu - ref) / scale) < 1e-10
    assert max_row <= 16 and max_ops <= 48          # the planned kernel's buffers hold these patterns


# ---- round-2 schedule (csrc/sweep2.hpp): warp groups that stream their own records -------------------------------------------

@pytest.mark.parametrize("shape,parts,window,ext,cw,helpers,groups,wg", [
    ((1, 1, 1), 148, 0, 0, 0, 0, 0, 0),
    ((3, 1, 1), 148, 0, 0, 0, 0, 0, 0),
    ((40, 1, 1), 4, 0, 0, 0, 0, 0, 0),
    ((12, 10, 8), 148, 0, 0, 0, 0, 0, 0),
    ((12, 10, 8), 7, 64, 256, 6, 2, 3, 2),
    ((12, 10, 8), 7, 64, 256, 4, 1, 1, 1),           # one group: arrive and sync on the same barrier
    ((20, 18, 16), 148, 0, 0, 0, 0, 0, 0),
    ((20, 18, 16), 37, 256, 128, 12, 3, 0, 0),       # small external ring: slot reuse under flow control
    ((20, 18, 16), 5, 64, 1024, 12, 4, 0, 4),         # steps far larger than the group width: several records per warp and step
    ((30, 30, 1), 16, 64, 0, 0, 0, 0, 0),
    ((1, 25, 25), 148, 0, 0, 0, 0, 0, 0),
    ((36, 56, 22), 148, 0, 0, 0, 0, 0, 0),
    ((48, 48, 48), 148, 0, 0, 0, 0, 0, 0),
])
def test_round2_emulated_sweeps_match_sequential_substitution(built, shape, parts, window, ext, cw, helpers, groups, wg):
    from opm_autodiff_b200 import bridge
    rows, cols = grid_pattern(*shape)
    err, st = bridge.sweep2_schedule_check_host(rows, cols, parts, window, ext, cw, helpers, groups, wg, seed=3, relax=0.9)
    assert err < 1e-11
    assert 1 <= st["parts"] <= max(1, parts)
    assert st["lanes_L"] >= len(rows) - 1
    assert st["window_deps_L"] + st["external_deps_L"] == (len(cols) - (len(rows) - 1)) // 2


def test_round2_fault_pattern_long_rows_and_nonsymmetric_patterns(built):
    from opm_autodiff_b200 import bridge
    rows, cols = grid_pattern(14, 9, 11, nnc_planes=2)
    err, st = bridge.sweep2_schedule_check_host(rows, cols, 40, 128, 256, seed=5)
    assert err < 1e-11
    # random long-range symmetric couplings: rows with more than three lower / upper neighbours take continuation records
    Nb = len(rows) - 1
    rng = np.random.default_rng(7)
    nb = [set(cols[rows[i]:rows[i + 1]]) for i in range(Nb)]
    for _ in range(Nb // 5):
        a, b = rng.integers(0, Nb, 2)
        nb[a].add(int(b)); nb[b].add(int(a))
    hub = Nb // 2                                      # one very long row (a well folded into the matrix)
    for b in rng.integers(0, Nb, 40):
        nb[hub].add(int(b)); nb[int(b)].add(hub)
    r2 = np.zeros(Nb + 1, np.int32)
    c2 = []
    for i in range(Nb):
        c2.extend(sorted(nb[i]))
        r2[i + 1] = len(c2)
    err, st = bridge.sweep2_schedule_check_host(r2, np.array(c2, np.int32), 23, 64, 512, seed=9)
    assert err < 1e-10
    # structurally non-symmetric: drop some upper entries
    r3 = np.zeros(Nb + 1, np.int32)
    c3 = []
    for i in range(Nb):
        for c in cols[rows[i]:rows[i + 1]]:
            if c > i and rng.random() < 0.3:
                continue
            c3.append(int(c))
        r3[i + 1] = len(c3)
    err, st = bridge.sweep2_schedule_check_host(r3, np.array(c3, np.int32), 19, 64, 0, seed=13)
    assert err < 1e-11


def test_round2_diagonal_and_dense_corner_cases(built):
    from opm_autodiff_b200 import bridge
    # diagonal matrix: one step holding every row
    Nb = 300
    rows = np.arange(Nb + 1, dtype=np.int32)
    cols = np.arange(Nb, dtype=np.int32)
    err, st = bridge.sweep2_schedule_check_host(rows, cols, 5, 0, 0, seed=2)
    assert err < 1e-12
    # dense lower + upper triangle of 40 rows: one row per step, up to 39 dependencies per row
    n = 40
    rows = np.arange(0, n * n + 1, n, dtype=np.int32)
    cols = np.tile(np.arange(n, dtype=np.int32), n)
    err, st = bridge.sweep2_schedule_check_host(rows, cols, 3, 0, 0, seed=4)
    assert err < 1e-9
