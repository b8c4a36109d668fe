"""Multi-GPU parity (run with -m gpu): one process per GPU, halo over peer memory, NCCL all-reduce.
The world-1 case runs on any B200 box (it exercises the split reduce -> finish kernels); the world-2/4
cases need that many GPUs and are skipped otherwise."""
import os
import socket

import numpy as np
import pytest

from tests.helpers import oracle_wells, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(built):
    from opm_autodiff_b200 import bridge, dist, synth
    from oracle import oracle
    if not bridge.device_available():
        pytest.fail("GPU tests need a B200; the product has no CPU fallback")
    return bridge, dist, synth, oracle


def test_world1_dist_mode_matches_oracle(mods):
    bridge, dist, synth, oracle = mods
    s = synth.small(12, 10, 8, faults=((6, 1),), nwells=3, nperf=4)
    ranges = dist.slab_ranges(8, 120, 1)
    ls = dist.partition_global(s.rows, s.cols, s.vals, s.b, ranges, s.x_true, s.wells)[0]
    assert ls.n_ghost == 0
    ds = dist.DistSolver(ls, 0, maxit=200, tolerance=1e-10)
    res = ds.solve_system()
    x = ds.get_result()
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200)
    assert res.converged and abs(res.it - ref.it) <= max(1.0, 0.1 * ref.it)
    assert relerr(x, ref.x) < 1e-6


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _rank_main(rank, world, port, shape, faults, out, blocks=False):
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("gloo", rank=rank, world_size=world)      # set-up exchange only; data path is in the library
    try:
        from opm_autodiff_b200 import dist, synth
        cfg = synth.GridConfig("t", *shape, seed=5, faults=faults, nwells=4, nperf=3)
        ls = dist.block_system(cfg, rank, world) if blocks else dist.slab_system(cfg, rank, world)
        ds = dist.DistSolver(ls, rank, maxit=200, tolerance=1e-10)
        ds.upload()
        rng = np.random.default_rng(17)
        xg = rng.normal(size=3 * cfg.ncells)
        y = ds.spmv(xg[3 * ls.row0:3 * ls.row1])
        res = ds.solve_system()
        x = ds.get_result()
        res2 = ds.solve_resident()                      # second solve on the resident system: epochs keep advancing
        x2 = ds.get_result()
        out[rank] = (ls.row0, ls.row1, y, res.converged, res.it, x, res2.it, x2, ds.be.launch_count())
        td.barrier()
    finally:
        td.destroy_process_group()


# (tiny slabs of 2-4 planes, then realistic ones: 16 planes per rank at 40 x 36 x 32 / 2, 12 at 48^3 / 4, 8 at 64^3 / 8,
# wells inside the slabs, against the oracle's own partitioned solve = the reference's `mpirun -np N` semantics,
# PreconditionerFactory.hpp:237-252, WellOperators.hpp:200-214)
@pytest.mark.parametrize("world,shape,faults", [(2, (12, 10, 8), ((6, 1),)), (4, (9, 7, 12), ()),
                                                (2, (40, 36, 32), ((20, 1),)), (4, (48, 48, 48), ()), (8, (64, 64, 64), ()),
                                                (8, (20, 16, 24), ((9, 1),))])
def test_multi_gpu_solve_matches_partitioned_oracle(mods, world, shape, faults):
    bridge, dist, synth, oracle = mods
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main, args=(world, _free_port(), shape, faults, out), nprocs=world, join=True)
    cfg = synth.GridConfig("t", *shape, seed=5, faults=faults, nwells=4, nperf=3)
    s = synth.full_system(cfg)
    ranges = dist.slab_ranges(shape[2], shape[0] * shape[1], world)
    part_ptr = np.array([r[0] for r in ranges] + [s.Nb], np.int32)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200, part_ptr=part_ptr)
    rng = np.random.default_rng(17)
    xg = rng.normal(size=3 * s.Nb)
    y_ref = oracle.spmv(s.rows, s.cols, s.vals, xg)
    x = np.zeros(3 * s.Nb)
    x2 = np.zeros(3 * s.Nb)
    for rank in range(world):
        r0, r1, y, conv, it, xl, it2, xl2, launches = out[rank]
        assert relerr(y, y_ref[3 * r0:3 * r1]) < 1e-13           # halo exchange + owned/ghost SpMV
        assert conv and abs(it - ref.it) <= max(1.0, 0.1 * ref.it) and it2 == it
        assert launches > 0
        x[3 * r0:3 * r1] = xl
        x2[3 * r0:3 * r1] = xl2
    assert relerr(x, ref.x) < 1e-6 and relerr(x2, ref.x) < 1e-6
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, x, oracle_wells(s.wells)) < 1e-8


def test_four_gpu_block_partition_matches_partitioned_oracle(mods):
    """2 x 2 blocks in (y, z) with the rank-major renumbering (what bench.py --gpus 4 runs)."""
    bridge, dist, synth, oracle = mods
    import torch
    world, shape, faults = 4, (9, 8, 10), ((4, 1),)
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main, args=(world, _free_port(), shape, faults, out, True), nprocs=world, join=True)
    cfg = synth.GridConfig("t", *shape, seed=5, faults=faults, nwells=4, nperf=3)
    s = synth.full_system(cfg)
    bp = dist.block_partition(cfg.nx, cfg.ny, cfg.nz, world)
    rows, cols, vals, b, xt, wells, new_of = dist.permute_to_blocks(s, bp)
    part_ptr = np.array([r[0] for r in bp.ranges] + [s.Nb], np.int32)
    ref = oracle.solve(rows, cols, vals, b, oracle_wells(wells), tol=1e-10, maxit=200, part_ptr=part_ptr)
    rng = np.random.default_rng(17)
    xg = rng.normal(size=3 * s.Nb)                 # the ranks drew the same vector and used their [row0,row1) slice of it
    y_ref = oracle.spmv(rows, cols, vals, xg)
    x = np.zeros(3 * s.Nb)
    for rank in range(world):
        r0, r1, y, conv, it, xl, it2, xl2, launches = out[rank]
        assert relerr(y, y_ref[3 * r0:3 * r1]) < 1e-13
        assert conv and abs(it - ref.it) <= max(1.0, 0.1 * ref.it) and it2 == it
        x[3 * r0:3 * r1] = xl
    assert relerr(x, ref.x) < 1e-6


def _mswell_case(synth, dist, world, shape):
    """The whole system with standard wells + two multisegment wells inside every rank's rows (global column ids)."""
    cfg = synth.GridConfig("t", *shape, seed=5, faults=(), nwells=4, nperf=3)
    s = synth.full_system(cfg)
    ranges = dist.slab_ranges(shape[2], shape[0] * shape[1], world)
    per_rank = [synth.add_mswells(s, 2, 6, seed=31 + r, cell_range=ranges[r]) for r in range(world)]
    return s, ranges, per_rank


def _rank_main_ms(rank, world, port, shape, out):
    import copy
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from opm_autodiff_b200 import dist, synth
        from tests.helpers import add_bridge_mswells
        s, ranges, per_rank = _mswell_case(synth, dist, world, shape)
        ls = dist.partition_global(s.rows, s.cols, s.vals, s.b, ranges, s.x_true, s.wells)[rank]
        ds = dist.DistSolver(ls, rank, maxit=200, tolerance=1e-10)
        mine = []
        for m in per_rank[rank]:                        # this rank's multisegment wells with LOCAL column ids
            m = copy.copy(m)
            m.BcolIndices = (np.asarray(m.BcolIndices, np.int64) - ranges[rank][0]).astype(np.uint32)
            mine.append(m)
        add_bridge_mswells(ds.wc, mine)
        res = ds.solve_system()
        x = ds.get_result()
        res2 = ds.solve_resident()
        out[rank] = (ls.row0, ls.row1, res.converged, res.it, x, res2.it, ds.get_result())
        td.barrier()
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (12, 10, 8)), (4, (24, 20, 16))])
def test_multi_gpu_solve_with_multisegment_wells(mods, world, shape):
    """Standard + multisegment wells in the operator of a partitioned solve: every well lives inside one rank (the reference's
    default, AllowDistributedWells = false, ebos/eclbasevanguard.hh:148-150), its apply and the patch of the dot products stay
    rank-local, the all-reduce carries the patched sums.  Against the oracle's partitioned solve with the same wells."""
    bridge, dist, synth, oracle = mods
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    from tests.helpers import oracle_mswells
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main_ms, args=(world, _free_port(), shape, out), nprocs=world, join=True)
    s, ranges, per_rank = _mswell_case(synth, dist, world, shape)
    om = oracle_mswells([m for ms in per_rank for m in ms])
    part_ptr = np.array([r[0] for r in ranges] + [s.Nb], np.int32)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200, part_ptr=part_ptr, mswells=om)
    assert ref.converged
    x = np.zeros(3 * s.Nb)
    for rank in range(world):
        r0, r1, conv, it, xl, it2, xl2 = out[rank]
        assert conv and abs(it - ref.it) <= max(1.0, 0.1 * ref.it) and it2 == it
        assert np.array_equal(xl, xl2)
        x[3 * r0:3 * r1] = xl
    assert relerr(x, ref.x) < 1e-6
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, x, oracle_wells(s.wells), om) < 1e-8


def _rank_main_c4(rank, world, port, golden, out):
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from opm_autodiff_b200 import dist, synth
        cfg = synth.CONFIGS["c4"]
        ls = dist.slab_system(cfg, rank, world)                      # every rank generates only its own rows
        ds = dist.DistSolver(ls, rank, maxit=200, tolerance=1e-10)
        res = ds.solve_system()
        x = ds.get_result().reshape(-1, 3)
        rows = np.asarray(golden["sample_rows"], np.int64)
        mine = (rows >= ls.row0) & (rows < ls.row1)
        xs = np.asarray(golden["sample_x"], np.float64)[mine]
        got = x[rows[mine] - ls.row0]
        out[rank] = (bool(res.converged), float(res.it), float(np.sum((got - xs) ** 2)), float(np.sum(xs ** 2)), int(mine.sum()),
                     float(np.sum(x ** 2)))
        td.barrier()
    finally:
        td.destroy_process_group()


def test_c4_eight_gpu_solve_matches_the_partitioned_oracle_fingerprint(mods):
    """BASELINE.json's C4 (250 x 200 x 200 = 10 M cells, 200 wells) on 8 row slabs against the CPU oracle's 8-partition solve of
    the same system (tools/c4_golden.py, run once on a CPU box: tests/golden/c4_n8_oracle.json holds its iteration count, the
    norm of its solution and the solution at 4000 sampled block rows): iterations within 10 %, the sampled solution to 1e-6."""
    import json
    bridge, dist, synth, oracle = mods
    import torch
    if torch.cuda.device_count() < 8:
        pytest.skip("needs 8 GPUs")
    path = os.path.join(os.path.dirname(__file__), "golden", "c4_n8_oracle.json")
    golden = json.load(open(path))
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main_c4, args=(8, _free_port(), golden, out), nprocs=8, join=True)
    err2 = sum(out[r][2] for r in range(8))
    ref2 = sum(out[r][3] for r in range(8))
    assert sum(out[r][4] for r in range(8)) == len(golden["sample_rows"])
    assert all(out[r][0] for r in range(8))
    it = out[0][1]
    assert all(out[r][1] == it for r in range(8))
    assert abs(it - golden["iterations"]) <= max(1.0, 0.1 * golden["iterations"])
    assert np.sqrt(err2 / ref2) <= 1e-6
    assert abs(np.sqrt(sum(out[r][5] for r in range(8))) / golden["x_norm"] - 1.0) < 1e-6
    print("C4 on 8 GPUs: %.1f iterations (oracle %.1f), sampled |x - x_ref| / |x_ref| = %.2e" % (it, golden["iterations"], np.sqrt(err2 / ref2)))
