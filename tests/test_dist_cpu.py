"""Host logic of the multi-GPU path on the CPU: slab partition, ghost-last numbering, halo plan, and a
world_size-2 gloo run of the WHOLE distributed algorithm (block-Jacobi ILU0 + BiCGSTAB with halo exchange and
all-reduced dot products) with the oracle's kernels standing in for the CUDA ones -- compared with the
oracle's own partitioned solve.  The same dist.py functions drive the GPU ranks."""
import os
import socket

import numpy as np
import pytest

from tests.helpers import oracle_wells, relerr


@pytest.fixture(scope="module")
def mods(built):
    from opm_autodiff_b200 import dist, synth
    from oracle import oracle
    return dist, synth, oracle


@pytest.mark.parametrize("shape,world,faults", [((6, 5, 8), 2, ()), ((5, 4, 9), 3, ((2, 1),)), ((4, 3, 8), 8, ())])
def test_partition_halo_spmv_matches_global(mods, shape, world, faults):
    dist, synth, oracle = mods
    s = synth.small(*shape, faults=faults)
    ranges = dist.slab_ranges(shape[2], shape[0] * shape[1], world)
    assert ranges[0][0] == 0 and ranges[-1][1] == s.Nb and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    parts = dist.partition_global(s.rows, s.cols, s.vals, s.b, ranges, s.x_true)
    rng = np.random.default_rng(3)
    x = rng.normal(size=3 * s.Nb)
    y_ref = oracle.spmv(s.rows, s.cols, s.vals, x)
    xo = [x[3 * r0:3 * r1] for r0, r1 in ranges]
    ghosts = dist.halo_exchange_host(parts, xo)
    for ls, xg in zip(parts, ghosts):
        # ghost-last numbering, ghosts grouped by owner in ascending rank order
        assert ls.cols.min() >= 0 and ls.cols.max() < ls.n_owned + ls.n_ghost
        assert np.all(np.diff(ls.ghost_owner) >= 0)
        assert np.array_equal(xg, x.reshape(-1, 3)[ls.ghost_global].reshape(-1))
        y = dist.local_spmv_host(ls, xo[ls.rank], xg)
        assert relerr(y, y_ref[3 * ls.row0:3 * ls.row1]) < 1e-14
        # the plan is symmetric for a structurally symmetric pattern and only touches slab faces
        for n, peer in enumerate(ls.neigh_rank):
            assert abs(peer - ls.rank) == 1 or faults
            assert ls.send_ptr[n + 1] - ls.send_ptr[n] > 0 and ls.recv_ptr[n + 1] - ls.recv_ptr[n] > 0


@pytest.mark.parametrize("world,shape,faults", [(2, (6, 8, 9), ()), (4, (6, 8, 9), ((3, 1),)), (8, (5, 8, 8), ()), (6, (4, 9, 6), ())])
def test_block_partition_halo_spmv_matches_global(mods, world, shape, faults):
    """y-z block partition with the rank-major renumbering: every rank generates only its block; halo plan, local SpMV and
    right-hand sides must reproduce the whole system permuted into that numbering."""
    dist, synth, oracle = mods
    cfg = synth.GridConfig("t", *shape, seed=5, faults=faults, nwells=4, nperf=3)
    s = synth.full_system(cfg)
    bp = dist.block_partition(cfg.nx, cfg.ny, cfg.nz, world)
    assert bp.py * bp.pz == world and bp.ranges[0][0] == 0 and bp.ranges[-1][1] == s.Nb
    rows, cols, vals, b, xt, wells, new_of = dist.permute_to_blocks(s, bp)
    assert sorted(new_of.tolist()) == list(range(s.Nb))
    rng = np.random.default_rng(1)
    x = rng.normal(size=3 * s.Nb)
    xp = np.empty_like(x)
    xp.reshape(-1, 3)[new_of] = x.reshape(-1, 3)
    yp = oracle.spmv(rows, cols, vals, xp)
    assert relerr(yp.reshape(-1, 3)[new_of], oracle.spmv(s.rows, s.cols, s.vals, x).reshape(-1, 3)) < 1e-14
    parts = [dist.block_system(cfg, r, world) for r in range(world)]
    reqs = [dist.requests_of(ls) for ls in parts]
    for ls in parts:
        dist.plan_from_requests(ls, reqs)
    xo = [xp[3 * r0:3 * r1] for r0, r1 in bp.ranges]
    ghosts = dist.halo_exchange_host(parts, xo)
    nw = 0
    for ls, xg in zip(parts, ghosts):
        assert relerr(dist.local_spmv_host(ls, xo[ls.rank], xg), yp[3 * ls.row0:3 * ls.row1]) < 1e-14
        assert np.allclose(ls.b, b[3 * ls.row0:3 * ls.row1]) and np.allclose(ls.x_true, xt[3 * ls.row0:3 * ls.row1])
        nw += 0 if ls.wells is None else ls.wells.nwells
    assert nw == 4                                                   # horizontal wells never straddle a y or z cut
    # the partitioned oracle on the permuted system is the parity reference of the multi-GPU run: it must converge
    part_ptr = np.array([r[0] for r in bp.ranges] + [s.Nb], np.int32)
    ref = oracle.solve(rows, cols, vals, b, oracle_wells(wells), tol=1e-10, maxit=300, part_ptr=part_ptr)
    assert ref.converged and relerr(ref.x, xt) < 1e-5


def test_wells_must_not_span_ranks(mods):
    dist, synth, oracle = mods
    s = synth.small(6, 5, 8, nwells=3, nperf=3)
    ranges = dist.slab_ranges(8, 30, 2)
    parts = dist.partition_global(s.rows, s.cols, s.vals, s.b, ranges, s.x_true, s.wells)
    assert sum(0 if p.wells is None else p.wells.nwells for p in parts) == 3
    import copy
    w = copy.copy(s.wells)
    w.Bcols = np.array(w.Bcols).copy()
    w.Bcols[0] = 239 if w.Bcols[0] < 120 else 0     # force one perforation into the other rank
    with pytest.raises(ValueError):
        dist.partition_global(s.rows, s.cols, s.vals, s.b, ranges, s.x_true, w)


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _rank_main(rank, world, port, shape, out):
    """One rank of the emulated multi-GPU solve (CPU, gloo)."""
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from opm_autodiff_b200 import dist, synth
        from oracle import oracle
        oracle.lib().orc_set_threads(1)
        cfg = synth.GridConfig("t", *shape, seed=5, nwells=2, nperf=3)
        ls = dist.slab_system(cfg, rank, world)
        reqs = dist._all_gather_object(dist.requests_of(ls))
        dist.plan_from_requests(ls, reqs)
        n = ls.n_owned
        # block-Jacobi ILU0: owned x owned block only
        keep = ls.cols < n
        rid = np.repeat(np.arange(n), np.diff(ls.rows))
        sq_rows = np.concatenate([[0], np.cumsum(np.bincount(rid[keep], minlength=n))]).astype(np.int32)
        sq_cols, sq_vals = ls.cols[keep], ls.vals[keep]
        LU, diag, st = oracle.ilu0(sq_rows, sq_cols, sq_vals)
        assert st == 0
        wells = None if ls.wells is None else oracle.Wells(ls.wells.val_pointers, ls.wells.Bcols, ls.wells.Ccols, ls.wells.B,
                                                            ls.wells.C, ls.wells.Dinv)

        def halo(x):
            xg = np.zeros(3 * ls.n_ghost)
            ops, bufs = [], []
            for k, peer in enumerate(ls.neigh_rank):
                rows_ = ls.send_rows[ls.send_ptr[k]:ls.send_ptr[k + 1]]
                sb = torch.from_numpy(x.reshape(-1, 3)[rows_].reshape(-1).copy())
                rb = torch.zeros(3 * int(ls.recv_ptr[k + 1] - ls.recv_ptr[k]), dtype=torch.float64)
                bufs.append((k, rb))
                ops += [td.P2POp(td.isend, sb, peer), td.P2POp(td.irecv, rb, peer)]
            for w_ in td.batch_isend_irecv(ops):
                w_.wait()
            for k, rb in bufs:
                xg[3 * int(ls.recv_ptr[k]):3 * int(ls.recv_ptr[k + 1])] = rb.numpy()
            return xg

        def op(x):
            y = dist.local_spmv_host(ls, x, halo(x))
            return oracle.well_apply(wells, x, y) if wells is not None else y

        def prec(d):
            return oracle.ilu0_apply(sq_rows, sq_cols, diag, LU, d)

        def dot(a, b):
            t = torch.tensor([float(np.dot(a, b))], dtype=torch.float64)
            td.all_reduce(t)
            return float(t[0])

        # Dune BiCGSTAB, restated as in oracle_bda.c:orc_solve
        tol, maxit = 1e-10, 200
        x = np.zeros(3 * n); r = ls.b.copy(); rt = r.copy(); p = np.zeros_like(r); v = np.zeros_like(r)
        norm0 = np.sqrt(dot(r, r)); rho = alpha = omega = 1.0
        it, conv = 0.5, False
        while it < maxit:
            rho_new = dot(rt, r)
            p = r.copy() if it < 1.0 else (p - omega * v) * ((rho_new / rho) * (alpha / omega)) + r
            y = prec(p); v = op(y)
            alpha = rho_new / dot(rt, v)
            x += alpha * y; r -= alpha * v
            if np.sqrt(dot(r, r)) < tol * norm0:
                conv = True; break
            it += 0.5
            y = prec(r); t = op(y)
            omega = dot(t, r) / dot(t, t)
            x += omega * y; r -= omega * t
            rho = rho_new
            if np.sqrt(dot(r, r)) < tol * norm0:
                conv = True; break
            it += 0.5
        out[rank] = (conv, it, ls.row0, ls.row1, x)
    finally:
        td.destroy_process_group()


def test_gloo_world2_block_jacobi_bicgstab_matches_partitioned_oracle(mods):
    dist, synth, oracle = mods
    import torch.multiprocessing as mp
    shape, world = (6, 5, 8), 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_rank_main, args=(world, port, shape, out), nprocs=world, join=True)
    cfg = synth.GridConfig("t", *shape, seed=5, nwells=2, nperf=3)
    s = synth.full_system(cfg)
    ranges = dist.slab_ranges(shape[2], shape[0] * shape[1], world)
    part_ptr = np.array([r[0] for r in ranges] + [s.Nb], np.int32)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200, part_ptr=part_ptr)
    x = np.zeros(3 * s.Nb)
    for rank in range(world):
        conv, it, r0, r1, xl = out[rank]
        assert conv and it == ref.it
        x[3 * r0:3 * r1] = xl
    assert relerr(x, ref.x) < 1e-9
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, x, oracle_wells(s.wells)) < 1e-9


def test_rank_local_multisegment_wells_keep_x_true(mods):
    """synth.add_mswells(cell_range=...) -- the generator of the multi-GPU multisegment-well test -- perforates only the rows
    of one rank, keeps x_true the solution of the whole operator (A - sum C^T D^-1 B) x = b, and the oracle's partitioned
    solve with those wells converges to it."""
    dist, synth, oracle = mods
    from tests.helpers import oracle_mswells
    cfg = synth.GridConfig("t", 12, 10, 8, seed=5, faults=(), nwells=4, nperf=3)
    s = synth.full_system(cfg)
    world = 2
    ranges = dist.slab_ranges(8, 120, world)
    per_rank = [synth.add_mswells(s, 2, 6, seed=31 + r, cell_range=ranges[r]) for r in range(world)]
    for r, ms in enumerate(per_rank):
        for m in ms:
            c = np.asarray(m.BcolIndices, np.int64)
            assert c.size and c.min() >= ranges[r][0] and c.max() < ranges[r][1]
    om = oracle_mswells([m for ms in per_rank for m in ms])
    assert oracle.true_residual(s.rows, s.cols, s.vals, s.b, s.x_true, oracle_wells(s.wells), om) < 1e-12
    part_ptr = np.array([a for a, _ in ranges] + [s.Nb], np.int32)
    ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200, part_ptr=part_ptr, mswells=om)
    assert ref.converged and relerr(ref.x, s.x_true) < 1e-5
