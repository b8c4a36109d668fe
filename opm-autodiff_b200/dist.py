"""Multi-GPU host layer: row-slab partition with ghosts numbered last, halo plan, per-rank solver.

What the reference's MPI CPU path does with Dune index sets is restated here with plain arrays
(SURVEY.md 8e):

  partition     contiguous block-row slabs; local numbering = owned rows first, then the ghost cells
                (--owner-cells-first, ISTLSolverEbos.hpp:171-180; findOverlapRowsAndColumns.hpp:119-139)
  operator      owned rows x (owned + ghost) vector after the halo exchange
                (WellModelGhostLastMatrixAdapter::apply, WellOperators.hpp:200-214)
  precond.      ILU0 of the owned x owned block per rank, no communication
                (PreconditionerFactory.hpp:237-252, ParallelOverlappingILU0.hpp:440-494)
  dot products  owned entries, summed over ranks

The pure-array functions (``slab_ranges``, ``localize``, ``plan_from_requests``, ``partition_global``)
need neither a GPU nor torch and are what the gloo tests exercise; ``DistSolver`` binds one
``B200SolverBackend`` per rank: NCCL id and CUDA-IPC handles travel through ``torch.distributed``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class LocalSystem:
    rank: int
    world: int
    row0: int                      # first owned global block row
    row1: int                      # one past the last owned global block row
    rows: np.ndarray               # int32 [n_owned + 1]
    cols: np.ndarray               # int32 local columns: owned 0..n_owned-1, ghosts n_owned + g
    vals: np.ndarray               # [nnzb, 3, 3]
    b: np.ndarray                  # [3 n_owned]
    ghost_global: np.ndarray       # int64 [n_ghost] global ids, grouped by owner rank, ascending
    ghost_owner: np.ndarray        # int32 [n_ghost]
    x_true: Optional[np.ndarray] = None
    wells: object = None           # synth.WellData with LOCAL columns (all perforations owned) or None
    # halo plan (filled by plan_from_requests)
    neigh_rank: List[int] = field(default_factory=list)
    send_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))
    send_rows: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    recv_ptr: np.ndarray = field(default_factory=lambda: np.zeros(1, np.int32))

    @property
    def n_owned(self) -> int:
        return self.row1 - self.row0

    @property
    def n_ghost(self) -> int:
        return len(self.ghost_global)

    @property
    def nnzb(self) -> int:
        return int(self.rows[-1])


def slab_ranges(n_planes: int, plane_rows: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous slabs along the slowest grid index: rank g owns planes [n*g//G, n*(g+1)//G)."""
    if world > n_planes:
        raise ValueError("more ranks than grid planes")
    cuts = [(n_planes * g) // world for g in range(world + 1)]
    return [(plane_rows * cuts[g], plane_rows * cuts[g + 1]) for g in range(world)]


def owner_of(global_ids: np.ndarray, ranges: Sequence[Tuple[int, int]]) -> np.ndarray:
    starts = np.array([r[0] for r in ranges], dtype=np.int64)
    return (np.searchsorted(starts, global_ids, side="right") - 1).astype(np.int32)


def localize(rank: int, ranges: Sequence[Tuple[int, int]], rows, cols_global, vals, b, x_true=None, wells=None) -> LocalSystem:
    """Rows [row0,row1) with GLOBAL columns -> ghost-last local numbering (no halo plan yet)."""
    row0, row1 = ranges[rank]
    n_owned = row1 - row0
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    assert len(rows) == n_owned + 1
    cg = np.asarray(cols_global, dtype=np.int64)
    owned = (cg >= row0) & (cg < row1)
    gids = np.unique(cg[~owned])                       # ascending global id == grouped by owner for slabs
    gown = owner_of(gids, ranges)
    order = np.lexsort((gids, gown))
    gids, gown = gids[order], gown[order]
    local = np.empty(len(cg), dtype=np.int32)
    local[owned] = (cg[owned] - row0).astype(np.int32)
    # ghost id -> position in the (owner, id)-sorted ghost list
    sorter = np.argsort(gids, kind="stable")
    pos = sorter[np.searchsorted(gids, cg[~owned], sorter=sorter)]
    local[~owned] = (n_owned + pos).astype(np.int32)
    # columns stay ascending inside a row only if ghosts of lower slabs sort after owned ones: the solver
    # does not need ascending ghost columns (it splits them off), but the owned part must stay ascending
    lw = None
    if wells is not None and wells.nwells > 0:
        import copy
        lw = copy.copy(wells)
        for name in ("Bcols", "Ccols"):
            g = np.asarray(getattr(wells, name), dtype=np.int64)
            if np.any((g < row0) | (g >= row1)):
                raise ValueError("a standard well has perforations outside the rows of rank %d: wells must not span ranks" % rank)
            setattr(lw, name, (g - row0).astype(np.int32))
    return LocalSystem(rank, len(ranges), row0, row1, rows, local, np.ascontiguousarray(vals, dtype=np.float64).reshape(-1, 3, 3),
                       np.ascontiguousarray(b, dtype=np.float64), gids, gown, x_true, lw)


def requests_of(ls: LocalSystem) -> Dict[int, np.ndarray]:
    """owner rank -> global ids this rank needs from it (in its ghost order)."""
    return {int(o): ls.ghost_global[ls.ghost_owner == o] for o in np.unique(ls.ghost_owner)}


def plan_from_requests(ls: LocalSystem, all_requests: Sequence[Dict[int, np.ndarray]]) -> None:
    """Fill the halo plan of ``ls`` from every rank's request table (index = requesting rank)."""
    recv_from = sorted(int(o) for o in np.unique(ls.ghost_owner))
    send_to = sorted(r for r, req in enumerate(all_requests) if r != ls.rank and ls.rank in req and len(req[ls.rank]))
    neigh = sorted(set(recv_from) | set(send_to))
    send_ptr, send_rows, recv_ptr = [0], [], [0]
    for n in neigh:
        want = all_requests[n].get(ls.rank, np.zeros(0, np.int64)) if n in send_to else np.zeros(0, np.int64)
        want = np.asarray(want, dtype=np.int64)
        if len(want) and (want.min() < ls.row0 or want.max() >= ls.row1):
            raise ValueError("rank %d asked rank %d for rows it does not own" % (n, ls.rank))
        send_rows.append((want - ls.row0).astype(np.int32))
        send_ptr.append(send_ptr[-1] + len(want))
        recv_ptr.append(recv_ptr[-1] + int(np.count_nonzero(ls.ghost_owner == n)))
    assert recv_ptr[-1] == ls.n_ghost
    # ghosts are grouped by owner in ascending rank order == order of `neigh` restricted to recv_from
    ls.neigh_rank = neigh
    ls.send_ptr = np.array(send_ptr, np.int32)
    ls.send_rows = np.concatenate(send_rows).astype(np.int32) if send_rows else np.zeros(0, np.int32)
    ls.recv_ptr = np.array(recv_ptr, np.int32)


def partition_global(rows, cols, vals, b, ranges, x_true=None, wells=None) -> List[LocalSystem]:
    """Cut a whole system (global int columns) into every rank's LocalSystem, halo plans included.
    Single-process helper for tests and small runs; production ranks generate only their own slab."""
    rows = np.asarray(rows)
    vals = np.asarray(vals).reshape(-1, 3, 3)
    out = []
    for g, (r0, r1) in enumerate(ranges):
        k0, k1 = int(rows[r0]), int(rows[r1])
        lw = None
        if wells is not None and wells.nwells > 0:
            lw = _wells_of_range(wells, r0, r1)
        out.append(localize(g, ranges, rows[r0:r1 + 1] - rows[r0], np.asarray(cols)[k0:k1], vals[k0:k1], np.asarray(b)[3 * r0:3 * r1],
                            None if x_true is None else np.asarray(x_true)[3 * r0:3 * r1], lw))
    reqs = [requests_of(ls) for ls in out]
    for ls in out:
        plan_from_requests(ls, reqs)
    return out


def _wells_of_range(wells, r0, r1):
    """Wells whose perforations all lie in [r0,r1) (global ids kept); wells touching the range only partly raise."""
    import copy
    keep = []
    for w in range(wells.nwells):
        s, e = int(wells.val_pointers[w]), int(wells.val_pointers[w + 1])
        c = np.asarray(wells.Bcols[s:e], dtype=np.int64)
        inside = (c >= r0) & (c < r1)
        if inside.all():
            keep.append(w)
        elif inside.any():
            raise ValueError("standard well %d spans two ranks" % w)
    if not keep:
        return None
    lw = copy.copy(wells)
    idx = np.concatenate([np.arange(int(wells.val_pointers[w]), int(wells.val_pointers[w + 1])) for w in keep])
    lw.val_pointers = np.concatenate([[0], np.cumsum([int(wells.val_pointers[w + 1] - wells.val_pointers[w]) for w in keep])]).astype(np.uint32)
    lw.Bcols = np.asarray(wells.Bcols)[idx]
    lw.Ccols = np.asarray(wells.Ccols)[idx]
    lw.B = np.asarray(wells.B)[idx]
    lw.C = np.asarray(wells.C)[idx]
    lw.Dinv = np.asarray(wells.Dinv)[keep]
    return lw


def halo_exchange_host(locals_: Sequence[LocalSystem], x_owned: Sequence[np.ndarray]) -> List[np.ndarray]:
    """Reference (numpy) halo exchange: returns every rank's ghost vector [3 n_ghost]."""
    out = [np.zeros(3 * ls.n_ghost) for ls in locals_]
    for ls in locals_:
        for n, peer in enumerate(ls.neigh_rank):
            rows = ls.send_rows[ls.send_ptr[n]:ls.send_ptr[n + 1]]
            if len(rows) == 0:
                continue
            pl = locals_[peer]
            slot = pl.neigh_rank.index(ls.rank)
            off = int(pl.recv_ptr[slot])
            assert int(pl.recv_ptr[slot + 1]) - off == len(rows)
            out[peer][3 * off:3 * (off + len(rows))] = np.asarray(x_owned[ls.rank]).reshape(-1, 3)[rows].reshape(-1)
    return out


def local_spmv_host(ls: LocalSystem, x_owned: np.ndarray, x_ghost: np.ndarray) -> np.ndarray:
    """numpy y_owned = A_local [x_owned; x_ghost]."""
    x = np.concatenate([np.asarray(x_owned).reshape(-1, 3), np.asarray(x_ghost).reshape(-1, 3)])
    rowid = np.repeat(np.arange(ls.n_owned), np.diff(ls.rows))
    contrib = np.einsum("kij,kj->ki", ls.vals, x[ls.cols])
    y = np.zeros((ls.n_owned, 3))
    np.add.at(y, rowid, contrib)
    return y.reshape(-1)


# ---- per-rank solver over torch.distributed ---------------------------------------------------------

def _all_gather_object(obj, group=None):
    import torch.distributed as td
    out = [None] * td.get_world_size(group)
    td.all_gather_object(out, obj, group=group)
    return out


def slab_system(cfg, rank: int, world: int) -> LocalSystem:
    """This rank's slab of a synthetic configuration, generated locally (no global matrix anywhere)."""
    from . import synth
    ranges = slab_ranges(cfg.nz, cfg.nx * cfg.ny, world)
    k0, k1 = ranges[rank][0] // (cfg.nx * cfg.ny), ranges[rank][1] // (cfg.nx * cfg.ny)
    s = synth.generate(cfg, k0, k1)
    return localize(rank, ranges, s.rows, s.cols, s.vals, s.b, s.x_true, s.wells)


# ---- 2-D block partition (y x z), x never cut -------------------------------------------------------------
#
# The triangular sweeps of a rank cost (levels of its sub-grid) x (time per level): a slab along z keeps nx + ny of the
# nx + ny + nz levels whatever the rank count, a y-z block keeps nx + ny/py + nz/pz.  Lines along x (and with them the
# horizontal wells of the synthetic configurations) stay inside one rank.  Cells are renumbered rank by rank, natural
# (i fastest, then j, then k) inside a rank: the new numbering plays the role of Dune's owner-first local ordering, and
# block-Jacobi ILU0 on it is what the reference's MPI run does with this partition.

@dataclass
class BlockPartition:
    nx: int
    ny: int
    nz: int
    py: int
    pz: int
    jcuts: List[int]
    kcuts: List[int]
    ranges: List[Tuple[int, int]]           # new-numbering row range of every rank (rank = bk * py + bj)

    @property
    def world(self) -> int:
        return self.py * self.pz

    def box(self, rank: int):
        bj, bk = rank % self.py, rank // self.py
        return self.jcuts[bj], self.jcuts[bj + 1], self.kcuts[bk], self.kcuts[bk + 1]

    def new_id(self, natural: np.ndarray) -> np.ndarray:
        """natural cell id (i + nx (j + ny k)) -> id in the rank-major numbering."""
        c = np.asarray(natural, dtype=np.int64)
        i = c % self.nx
        j = (c // self.nx) % self.ny
        k = c // (self.nx * self.ny)
        jc, kc = np.asarray(self.jcuts), np.asarray(self.kcuts)
        bj = np.searchsorted(jc, j, side="right") - 1
        bk = np.searchsorted(kc, k, side="right") - 1
        rank = bk * self.py + bj
        j0, k0 = jc[bj], kc[bk]
        nyr = jc[bj + 1] - j0
        base = np.asarray([r[0] for r in self.ranges], dtype=np.int64)[rank]
        return base + ((k - k0) * nyr + (j - j0)) * self.nx + i


def block_partition(nx: int, ny: int, nz: int, world: int) -> BlockPartition:
    """py x pz = world minimising the level count nx + ny/py + nz/pz of a rank's sub-grid (ties: fewer cuts in y)."""
    best = None
    for py in range(1, world + 1):
        if world % py:
            continue
        pz = world // py
        if py > ny or pz > nz:
            continue
        cost = -(-ny // py) + -(-nz // pz)
        if best is None or cost < best[0]:
            best = (cost, py, pz)
    if best is None:
        raise ValueError("more ranks than grid planes")
    _, py, pz = best
    jcuts = [(ny * b) // py for b in range(py + 1)]
    kcuts = [(nz * b) // pz for b in range(pz + 1)]
    ranges, o = [], 0
    for bk in range(pz):
        for bj in range(py):
            n = nx * (jcuts[bj + 1] - jcuts[bj]) * (kcuts[bk + 1] - kcuts[bk])
            ranges.append((o, o + n))
            o += n
    return BlockPartition(nx, ny, nz, py, pz, jcuts, kcuts, ranges)


def block_system(cfg, rank: int, world: int) -> LocalSystem:
    """This rank's y-z block of a synthetic configuration in the rank-major numbering, generated locally."""
    from . import synth
    import copy
    bp = block_partition(cfg.nx, cfg.ny, cfg.nz, world)
    j0, j1, k0, k1 = bp.box(rank)
    s = synth.generate(cfg, k0, k1)                         # planes [k0, k1), natural global column ids, wells of these planes
    plane = cfg.nx * cfg.ny
    loc = np.arange(plane * (k1 - k0), dtype=np.int64)
    jj = (loc // cfg.nx) % cfg.ny
    mine = (jj >= j0) & (jj < j1)
    rows_sel = np.nonzero(mine)[0]
    cnt = np.diff(s.rows)[rows_sel]
    rowptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    take = np.concatenate([np.arange(s.rows[r], s.rows[r + 1]) for r in rows_sel]) if len(rows_sel) else np.zeros(0, np.int64)
    cols_new = bp.new_id(np.asarray(s.cols)[take])
    e3 = (3 * rows_sel[:, None] + np.arange(3)[None, :]).reshape(-1)
    wells = None
    if s.wells is not None and s.wells.nwells > 0:
        keep = []
        for w in range(s.wells.nwells):
            a, e = int(s.wells.val_pointers[w]), int(s.wells.val_pointers[w + 1])
            cj = (np.asarray(s.wells.Bcols[a:e], dtype=np.int64) // cfg.nx) % cfg.ny
            inside = (cj >= j0) & (cj < j1)
            if inside.all():
                keep.append(w)
            elif inside.any():
                raise ValueError("standard well %d spans two ranks" % w)
        if keep:
            wells = copy.copy(s.wells)
            idx = np.concatenate([np.arange(int(s.wells.val_pointers[w]), int(s.wells.val_pointers[w + 1])) for w in keep])
            wells.val_pointers = np.concatenate([[0], np.cumsum([int(s.wells.val_pointers[w + 1] - s.wells.val_pointers[w]) for w in keep])]).astype(np.uint32)
            wells.Bcols = bp.new_id(np.asarray(s.wells.Bcols)[idx])
            wells.Ccols = bp.new_id(np.asarray(s.wells.Ccols)[idx])
            wells.B = np.asarray(s.wells.B)[idx]
            wells.C = np.asarray(s.wells.C)[idx]
            wells.Dinv = np.asarray(s.wells.Dinv)[keep]
    return localize(rank, bp.ranges, rowptr, cols_new, np.asarray(s.vals)[take], s.b[e3], s.x_true[e3], wells)


def permute_to_blocks(s, bp: BlockPartition):
    """A whole system (natural numbering) in the rank-major numbering of `bp`: (rows, cols, vals, b, x_true, wells) --
    the input of the partitioned oracle (contiguous parts) the multi-GPU run is compared with."""
    import copy
    Nb = len(s.rows) - 1
    new_of = bp.new_id(np.arange(Nb, dtype=np.int64))
    old_of = np.empty(Nb, dtype=np.int64)
    old_of[new_of] = np.arange(Nb)
    cnt = np.diff(s.rows)[old_of]
    rows = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    cols = np.empty(int(rows[-1]), dtype=np.int32)
    vals = np.empty((int(rows[-1]), 3, 3))
    sv = np.asarray(s.vals).reshape(-1, 3, 3)
    for q in range(Nb):
        r = old_of[q]
        a, e = int(s.rows[r]), int(s.rows[r + 1])
        c = new_of[np.asarray(s.cols[a:e], dtype=np.int64)]
        o = np.argsort(c, kind="stable")
        cols[rows[q]:rows[q + 1]] = c[o]
        vals[rows[q]:rows[q + 1]] = sv[a:e][o]
    e3 = (3 * old_of[:, None] + np.arange(3)[None, :]).reshape(-1)
    wells = None
    if s.wells is not None and s.wells.nwells > 0:
        wells = copy.copy(s.wells)
        wells.Bcols = new_of[np.asarray(s.wells.Bcols, dtype=np.int64)].astype(np.int32)
        wells.Ccols = new_of[np.asarray(s.wells.Ccols, dtype=np.int64)].astype(np.int32)
    return rows, cols, vals, np.asarray(s.b)[e3], np.asarray(s.x_true)[e3], wells, new_of


class DistSolver:
    """One rank of the multi-GPU ILU0-BiCGSTAB solve.  ``group`` is a torch.distributed process group
    (any backend) used only for the set-up exchange; the data path runs over peer memory and NCCL inside
    libb200bda.so."""

    def __init__(self, ls: LocalSystem, device: int, maxit: int = 200, tolerance: float = 1e-2, verbosity: int = 0,
                 group=None, options: Optional[dict] = None):
        import torch.distributed as td
        from . import bridge
        self.ls = ls
        self.bridge = bridge
        rank, world = ls.rank, ls.world
        if world > 1:
            assert td.is_initialized() and td.get_world_size(group) == world and td.get_rank(group) == rank
            reqs = _all_gather_object(requests_of(ls), group)
            plan_from_requests(ls, reqs)
        else:
            plan_from_requests(ls, [requests_of(ls)])
        self.be = bridge.B200SolverBackend(verbosity, maxit, tolerance, device)
        for k, v in (options or {}).items():
            self.be.set_option(k, v)
        uid = None
        if world > 1:
            uid = bridge.B200SolverBackend.dist_unique_id() if rank == 0 else None
            uid = _all_gather_object(uid, group)[0]
        self.be.dist_init(rank, world, uid)
        handle = self.be.dist_set_halo(ls.n_ghost, ls.neigh_rank, ls.send_ptr, ls.send_rows, ls.recv_ptr)
        if world > 1:
            infos = _all_gather_object({"handle": handle, "n_ghost": ls.n_ghost, "neigh": list(ls.neigh_rank),
                                        "recv_ptr": [int(v) for v in ls.recv_ptr]}, group)
            for r in range(world):                      # every rank: the dot products travel through peer-memory mailboxes
                self.be.dist_map_rank(r, None if r == rank else infos[r]["handle"])
            for n, peer in enumerate(ls.neigh_rank):
                pi = infos[peer]
                slot = pi["neigh"].index(rank)
                self.be.dist_connect_peer(n, pi["n_ghost"], pi["recv_ptr"][slot], slot)
            td.barrier(group)
        w = ls.wells
        self.wc = bridge.WellContributions("b200", False) if w is None else \
            bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv)
        self.N = 3 * ls.n_owned
        self.nnz = 9 * ls.nnzb
        self._registered = []

    def register_host_buffers(self, x: Optional[np.ndarray] = None) -> None:
        """Page-lock the slab's values and right-hand side (and the caller's solution vector): what the glue code of a Flow
        rank does once, since those buffers live as long as the simulator (b200_host_register)."""
        for a in (self.ls.vals, self.ls.b, x):
            if a is not None and a.size:
                self.be.host_register(a)
                self._registered.append(a)

    def solve_system(self, res=None):
        res = res or self.bridge.BdaResult()
        self.be.solve_system(self.N, self.nnz, 3, self.ls.vals, self.ls.rows, self.ls.cols, self.ls.b, self.wc, res)
        return res

    def upload(self):
        self.be.upload_system(self.N, self.nnz, 3, self.ls.vals, self.ls.rows, self.ls.cols, self.ls.b, self.wc)

    def solve_resident(self, res=None):
        res = res or self.bridge.BdaResult()
        self.be.solve_resident(res)
        return res

    def get_result(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Owned part of the solution; `out` lets a caller keep ONE solution vector over the solves, as Flow does (the
        library page-locks a vector it sees on consecutive calls)."""
        x = np.zeros(self.N) if out is None else out
        self.be.get_result(x)
        return x

    def spmv(self, x_owned) -> np.ndarray:
        return self.be.dist_spmv(x_owned)
