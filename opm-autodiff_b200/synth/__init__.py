"""Deterministic synthetic black-oil Jacobians for the BASELINE.json configs (harness input).

The arithmetic is in ``synth.c`` (hash-based, slab-independent); this module wraps it with
ctypes/numpy, adds the standard-well generator and the named configurations C2..C5 of
SURVEY.md 8(d).  Wells are horizontal (perforations along x at fixed (j, k)) so that a well
never crosses a k-slab boundary of the multi-GPU row partition.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libb200synth.so")
_GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"


class _Cfg(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("sigma", C.c_double),
                ("kvkh", C.c_double), ("acc_frac", C.c_double), ("offdiag_rand", C.c_double),
                ("seed", C.c_uint64), ("nfaults", C.c_int), ("fault_i", C.c_int * 4),
                ("fault_throw", C.c_int * 4), ("fault_mult", C.c_double)]


def build(force: bool = False) -> None:
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call([_GCC, "-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-shared",
                               "-o", _LIB, src, "-lm"])


_lib = None


def _L():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
        f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
        L.synth_count.argtypes = [C.POINTER(_Cfg), C.c_int, C.c_int, i64p]
        L.synth_count.restype = C.c_int64
        L.synth_fill.argtypes = [C.POINTER(_Cfg), C.c_int, C.c_int, i64p, i64p, f64p, f64p]
        L.synth_xtrue.argtypes = [C.POINTER(_Cfg), C.c_int64, C.c_int64, f64p]
        L.synth_halo_planes.argtypes = [C.POINTER(_Cfg)]
        L.synth_halo_planes.restype = C.c_int
        L.synth_u01_array.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int64, f64p]
        _lib = L
    return _lib


@dataclass
class GridConfig:
    name: str
    nx: int
    ny: int
    nz: int
    sigma: float = 1.0
    kvkh: float = 0.1
    acc_frac: float = 0.05
    offdiag_rand: float = 0.25
    seed: int = 1
    faults: tuple = ()          # ((i_plane, throw), ...)
    fault_mult: float = 0.5
    nwells: int = 0
    nperf: int = 20
    well_seed: int = 50
    well_rand: float = 0.05     # relative size of the random part of the B, C, D and perforation blocks

    @property
    def ncells(self) -> int:
        return self.nx * self.ny * self.nz

    def _c(self) -> _Cfg:
        c = _Cfg(self.nx, self.ny, self.nz, self.sigma, self.kvkh, self.acc_frac, self.offdiag_rand,
                 self.seed, len(self.faults))
        for q, (fi, thr) in enumerate(self.faults):
            c.fault_i[q] = fi
            c.fault_throw[q] = thr
        c.fault_mult = self.fault_mult
        return c


# BASELINE.json configs[1..4].  C2's fault planes give ~2% NNC blocks.
CONFIGS = {
    "c2": GridConfig("c2-norne-size-36x56x22", 36, 56, 22, sigma=1.0, seed=44431,
                     faults=((12, 1), (24, 2))),
    "c3": GridConfig("c3-100x100x100-50wells", 100, 100, 100, sigma=1.0, seed=1000003, nwells=50, nperf=20),
    "c4": GridConfig("c4-250x200x200", 250, 200, 200, sigma=1.0, seed=10000019),
    "c5": GridConfig("c5-500x400x250-heterogeneous", 500, 400, 250, sigma=3.0, kvkh=0.01, seed=50000017),
}


@dataclass
class WellData:
    """Standard wells, reference export layout (StandardWellEval.cpp:1202-1251)."""
    val_pointers: np.ndarray   # uint32 [nwells+1]
    Bcols: np.ndarray          # int32 [nblocks] (block-row indices, same numbering as the matrix columns)
    Ccols: np.ndarray          # int32 [nblocks]
    B: np.ndarray              # [nblocks,4,3]
    C: np.ndarray              # [nblocks,4,3]
    Dinv: np.ndarray           # [nwells,4,4]
    Tperf: np.ndarray = field(default_factory=lambda: np.zeros(0))  # perforation well indices

    @property
    def nwells(self) -> int:
        return len(self.val_pointers) - 1


@dataclass
class System:
    cfg: GridConfig
    k0: int
    k1: int
    rows: np.ndarray           # int32 [nrows+1]
    cols: np.ndarray           # int64 global block-column ids
    vals: np.ndarray           # [nnzb,3,3]
    b: np.ndarray              # [nrows*3]  (= A x_true, wells included)
    x_true: np.ndarray         # [nrows*3]  (owned rows)
    wells: Optional[WellData]
    row0: int                  # global id of the first owned row

    @property
    def Nb(self) -> int:
        return len(self.rows) - 1

    @property
    def nnzb(self) -> int:
        return int(self.rows[-1])


def _u01(seed, stream, idx0, n):
    out = np.empty(n)
    _L().synth_u01_array(seed, stream, idx0, n, out)
    return out


def make_wells(cfg: GridConfig, k0: int = 0, k1: Optional[int] = None):
    """Wells whose (j,k) line lies in planes [k0,k1).  Returns (WellData with GLOBAL cell ids,
    diag_add dict-like arrays (cells, 3x3 blocks) to be added to A's diagonal)."""
    k1 = cfg.nz if k1 is None else k1
    if cfg.nwells == 0:
        return None, None, None
    seed = cfg.well_seed
    wr = cfg.well_rand
    nperf = min(cfg.nperf, cfg.nx)
    cs = np.array([1e-7, 1.0, 1.0])
    E = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.3, 0.3, 0.3]], dtype=np.float64)
    wptr, Bc, Bv, Cv, Dv, cells_all, dblk_all, Tall = [0], [], [], [], [], [], [], []
    used = set()
    for w in range(cfg.nwells):
        u = _u01(seed, 11, w * 64, 64)
        t = 0
        while True:            # distinct (j,k) lines
            j = int(u[t % 64] * cfg.ny) % cfg.ny
            k = int(u[(t + 1) % 64] * cfg.nz) % cfg.nz
            t += 2
            if (j, k) not in used or t > 60:
                used.add((j, k))
                break
        i0 = int(u[62] * (cfg.nx - nperf + 1))
        if not (k0 <= k < k1):
            continue
        cells = np.arange(i0, i0 + nperf, dtype=np.int64) + cfg.nx * (j + cfg.ny * k)
        T = 0.5 + 1.5 * _u01(seed, 12, w * 1024, nperf)           # well indices O(1), like T_ij
        RB = 2 * _u01(seed, 13, w * 65536, nperf * 12).reshape(nperf, 4, 3) - 1
        RC = 2 * _u01(seed, 14, w * 65536, nperf * 12).reshape(nperf, 4, 3) - 1
        RD = 2 * _u01(seed, 15, w * 64, 16).reshape(4, 4) - 1
        RA = 2 * _u01(seed, 16, w * 65536, nperf * 9).reshape(nperf, 3, 3) - 1
        B = -T[:, None, None] * (E[None] + wr * RB) * cs[None, None, :]
        Cm = -T[:, None, None] * (E[None] + wr * RC)
        D = T.sum() * (np.eye(4) * np.array([1.0, 1.0, 1.0, 1.9]) + wr * RD)
        Dinv = np.linalg.inv(D)
        dblk = T[:, None, None] * (np.eye(3)[None] + wr * RA) * cs[None, None, :]
        Bc.append(cells); Bv.append(B); Cv.append(Cm); Dv.append(Dinv)
        cells_all.append(cells); dblk_all.append(dblk); Tall.append(T)
        wptr.append(wptr[-1] + nperf)
    if len(Bc) == 0:
        return None, None, None
    wd = WellData(np.array(wptr, np.uint32), np.concatenate(Bc), np.concatenate(Bc).copy(),
                  np.concatenate(Bv), np.concatenate(Cv), np.stack(Dv), np.concatenate(Tall))
    return wd, np.concatenate(cells_all), np.concatenate(dblk_all)


def generate(cfg: GridConfig, k0: int = 0, k1: Optional[int] = None) -> System:
    """Rows of planes [k0,k1) with GLOBAL column ids (int64); b = (A - C^T D^-1 B) x_true."""
    L = _L()
    k1 = cfg.nz if k1 is None else k1
    cc = cfg._c()
    plane = cfg.nx * cfg.ny
    nrows = plane * (k1 - k0)
    rowptr = np.zeros(nrows + 1, np.int64)
    nnzb = L.synth_count(C.byref(cc), k0, k1, rowptr)
    cols = np.empty(nnzb, np.int64)
    vals = np.empty(nnzb * 9)
    b = np.empty(nrows * 3)
    L.synth_fill(C.byref(cc), k0, k1, rowptr, cols, vals, b)
    xt = np.empty(nrows * 3)
    row0 = k0 * plane
    L.synth_xtrue(C.byref(cc), row0, nrows, xt)
    vals = vals.reshape(-1, 3, 3)
    wells, wcells, wdiag = make_wells(cfg, k0, k1)
    if wells is not None:
        # perforation conductance on A's diagonal, then b += A_add x_true - C^T D^-1 B x_true
        loc = wcells - row0
        dpos = rowptr[loc] + np.array([np.searchsorted(cols[rowptr[r]:rowptr[r + 1]], r + row0) for r in loc])
        assert np.all(cols[dpos] == wcells)
        np.add.at(vals, dpos, wdiag)
        xl = xt.reshape(-1, 3)
        np.add.at(b.reshape(-1, 3), loc, np.einsum("prc,pc->pr", wdiag, xl[loc]))
        for w in range(wells.nwells):
            s, e = int(wells.val_pointers[w]), int(wells.val_pointers[w + 1])
            cl = wells.Bcols[s:e] - row0
            z1 = np.einsum("prc,pc->r", wells.B[s:e], xl[cl])
            z2 = wells.Dinv[w] @ z1
            np.subtract.at(b.reshape(-1, 3), cl, np.einsum("prc,r->pc", wells.C[s:e], z2))
    return System(cfg, k0, k1, rowptr.astype(np.int32), cols, vals, b, xt, wells, row0)


def full_system(name_or_cfg) -> System:
    """Whole system with int32 columns (single-GPU use)."""
    cfg = CONFIGS[name_or_cfg] if isinstance(name_or_cfg, str) else name_or_cfg
    s = generate(cfg)
    s.cols = s.cols.astype(np.int32)
    if s.wells is not None:
        s.wells.Bcols = s.wells.Bcols.astype(np.int32)
        s.wells.Ccols = s.wells.Ccols.astype(np.int32)
    return s


def small(nx, ny, nz, seed=7, sigma=1.0, faults=(), nwells=0, nperf=4, kvkh=0.1, acc_frac=0.05) -> System:
    """Small test systems of the same family."""
    return full_system(GridConfig("small-%dx%dx%d" % (nx, ny, nz), nx, ny, nz, sigma=sigma, seed=seed,
                                  faults=tuple(faults), nwells=nwells, nperf=nperf, kvkh=kvkh,
                                  acc_frac=acc_frac))


@dataclass
class MSWellData:
    """One multisegment well in the layout of WellContributions::addMultisegmentWellContribution
    (bda/WellContributions.hpp:195-213): B, C blocked CSR with 4x3 blocks on one pattern, D scalar CSC."""
    Mb: int
    Bvalues: np.ndarray        # [nblocks,4,3]
    BcolIndices: np.ndarray    # uint32 [nblocks]
    BrowPointers: np.ndarray   # uint32 [Mb+1]
    DnumBlocks: int
    Dvalues: np.ndarray        # [16 DnumBlocks]
    DcolPointers: np.ndarray   # int32 [4 Mb + 1]
    DrowIndices: np.ndarray    # int32 [16 DnumBlocks]
    Cvalues: np.ndarray        # [nblocks,4,3]

    def dense_D(self) -> np.ndarray:
        M = 4 * self.Mb
        D = np.zeros((M, M))
        for c in range(M):
            for q in range(self.DcolPointers[c], self.DcolPointers[c + 1]):
                D[self.DrowIndices[q], c] += self.Dvalues[q]
        return D


def add_mswells(system: System, nwells: int, nseg: int, seed: int = 4242, share_cells: bool = True, cell_range=None):
    """Synthetic multisegment wells on a whole system (single-GPU use), in place: every well is a tree of `nseg` segments
    (segment s > 0 drains into a random earlier one; D couples a segment with its outlet, 4x4 blocks, as
    MultisegmentWell_impl.hpp:636-643 describes), segment 0 is the top segment without perforations (unless it is the only one), the others perforate
    0-2 cells (one segment per cell inside a well; with share_cells two wells may meet in a cell).  The perforation
    conductance goes on A's diagonal and b is updated so that x_true stays the solution:
    b += A_add x_true - sum_w C^T D^-1 B x_true.  cell_range = (lo, hi) perforates only block rows lo <= c < hi (the rows of one
    rank: a well does not span ranks, as with the reference's default AllowDistributedWells = false).  Returns the list of MSWellData."""
    rng = np.random.default_rng(seed)
    Nb = system.Nb
    c_lo, c_hi = cell_range if cell_range is not None else (0, Nb)
    cs = np.array([1e-7, 1.0, 1.0])
    E = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [0.3, 0.3, 0.3]], dtype=np.float64)
    wr = 0.25
    xt = system.x_true.reshape(-1, 3)
    bb = system.b.reshape(-1, 3)
    rows, cols = system.rows, system.cols
    vals = system.vals.reshape(-1, 3, 3)
    out = []
    prev_cells = np.zeros(0, np.int64)
    for w in range(nwells):
        nperf_seg = rng.integers(0, 3, size=nseg)
        nperf_seg[0] = 0 if nseg > 1 else 2
        if nseg > 1 and nperf_seg.sum() == 0:
            nperf_seg[-1] = 1
        nblk = int(nperf_seg.sum())
        cells = c_lo + rng.choice(c_hi - c_lo, size=nblk, replace=False).astype(np.int64)
        if share_cells and w > 0 and nblk > 0 and len(prev_cells) > 0:
            cells[0] = prev_cells[0]                    # two wells meet in one cell
            if len(set(cells.tolist())) != nblk:
                cells = c_lo + rng.choice(c_hi - c_lo, size=nblk, replace=False).astype(np.int64)
        prev_cells = cells
        Brow = np.concatenate([[0], np.cumsum(nperf_seg)]).astype(np.uint32)
        T = 0.5 + 1.5 * rng.random(nblk)
        B = -T[:, None, None] * (E[None] + wr * (2 * rng.random((nblk, 4, 3)) - 1)) * cs[None, None, :]
        Cm = -T[:, None, None] * (E[None] + wr * (2 * rng.random((nblk, 4, 3)) - 1))
        # D: block tree
        parent = np.array([-1] + [int(rng.integers(0, s)) for s in range(1, nseg)])
        flow = 0.5 + 1.5 * rng.random(nseg)
        blocks = {}
        segT = np.array([T[Brow[s]:Brow[s + 1]].sum() for s in range(nseg)])
        for s_ in range(nseg):
            blocks[(s_, s_)] = (segT[s_] + 1.0) * np.diag([1.0, 1.0, 1.0, 1.9]) + wr * (2 * rng.random((4, 4)) - 1)
        for s_ in range(1, nseg):
            q = parent[s_]
            blocks[(s_, q)] = -flow[s_] * (np.eye(4) + wr * (2 * rng.random((4, 4)) - 1))
            blocks[(q, s_)] = -flow[s_] * (np.eye(4) + wr * (2 * rng.random((4, 4)) - 1))
            blocks[(s_, s_)] += flow[s_] * np.diag([1.0, 1.0, 1.0, 1.9])
            blocks[(q, q)] += flow[s_] * np.diag([1.0, 1.0, 1.0, 1.9])
        M = 4 * nseg
        # scalar CSC with the block pattern (every entry of a block stored, zeros included, as UMFPACK gets it from Dune)
        colptr, rowidx, dvals = [0], [], []
        for c in range(M):
            bc, cc = divmod(c, 4)
            for br in sorted(r for (r, c2) in blocks if c2 == bc):
                for rr in range(4):
                    rowidx.append(4 * br + rr)
                    dvals.append(blocks[(br, bc)][rr, cc])
            colptr.append(len(rowidx))
        ms = MSWellData(nseg, B, cells.astype(np.uint32), Brow, len(blocks), np.array(dvals), np.array(colptr, np.int32),
                        np.array(rowidx, np.int32), Cm)
        out.append(ms)
        # A's diagonal and the right-hand side
        dblk = T[:, None, None] * (np.eye(3)[None] + wr * (2 * rng.random((nblk, 3, 3)) - 1)) * cs[None, None, :]
        for p_, cell in enumerate(cells):
            lo, hi = rows[cell], rows[cell + 1]
            d = lo + int(np.searchsorted(cols[lo:hi], cell))
            assert cols[d] == cell
            vals[d] += dblk[p_]
            bb[cell] += dblk[p_] @ xt[cell]
        z1 = np.zeros(M)
        for s_ in range(nseg):
            for blk in range(Brow[s_], Brow[s_ + 1]):
                z1[4 * s_:4 * s_ + 4] += B[blk] @ xt[cells[blk]]
        z2 = np.linalg.solve(ms.dense_D(), z1)
        for s_ in range(nseg):
            for blk in range(Brow[s_], Brow[s_ + 1]):
                bb[cells[blk]] -= Cm[blk].T @ z2[4 * s_:4 * s_ + 4]
    return out
