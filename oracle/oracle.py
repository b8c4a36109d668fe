"""ctypes/numpy front-end of the CPU ORACLE (test infrastructure, NOT product code).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The arithmetic lives in
``oracle_bda.c`` (each function there cites the reference file:line it restates);
``_ref/libref_reorder.so`` is the reference's own ``bda/Reorder.cpp`` compiled
unmodified and is used to pin the level sets.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_bda.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_reorder.so")
_REF_MSWELL_PATH = os.path.join(_HERE, "_ref", "libref_mswell.so")       # reference MultisegmentWellContribution (UMFPACK shimmed)
_REF_CUSPARSE_PATH = os.path.join(_HERE, "_ref", "libref_cusparse.so")   # incumbent GPU backend (tests/incumbent_cusparse.py)

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class _OrcResult(C.Structure):
    _fields_ = [("it", C.c_double), ("iterations", C.c_int), ("converged", C.c_int),
                ("breakdown", C.c_int), ("reduction", C.c_double), ("conv_rate", C.c_double),
                ("norm0", C.c_double), ("norm", C.c_double), ("t_decomp", C.c_double),
                ("t_solve", C.c_double)]


def build(force: bool = False) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "oracle_bda.c")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_bda.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/opm") and (force or not os.path.exists(_REF_PATH) or not os.path.exists(_REF_CUSPARSE_PATH)
                                                 or not os.path.exists(_REF_MSWELL_PATH)):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_inv3.argtypes = [_f64p, _f64p]
        L.orc_inv3.restype = C.c_int
        L.orc_check_zero_diagonal.argtypes = [C.c_int, _i32p, _i32p, _f64p]
        L.orc_check_zero_diagonal.restype = C.c_int
        L.orc_spmv.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p, _f64p]
        L.orc_well_apply.argtypes = [C.c_int, _u32p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.orc_ilu0_decompose_range.argtypes = [_i32p, _i32p, _f64p, _i32p, C.c_int, C.c_int]
        L.orc_ilu0_decompose_range.restype = C.c_int
        L.orc_ilu0_apply_range.argtypes = [_i32p, _i32p, _i32p, _f64p, _f64p, _f64p, C.c_double,
                                           C.c_int, C.c_int]
        L.orc_level_schedule.argtypes = [C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p]
        L.orc_level_schedule.restype = C.c_int
        L.orc_solve.argtypes = [C.c_int, _i32p, _i32p, _f64p, _f64p,
                                C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                _f64p, C.POINTER(_OrcResult), C.c_void_p, C.c_int]
        L.orc_solve.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        L.orc_ms_create.restype = C.c_void_p
        L.orc_ms_destroy.argtypes = [C.c_void_p]
        L.orc_ms_add.argtypes = [C.c_void_p, C.c_uint, C.c_uint, C.c_uint, _f64p, _u32p, _u32p, C.c_uint, _f64p,
                                 _i32p, _i32p, _f64p]
        L.orc_ms_add.restype = C.c_int
        L.orc_ms_apply.argtypes = [C.c_void_p, _f64p, _f64p]
        L.orc_attach_mswells.argtypes = [C.c_void_p]
        L.orc_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def ref_lib():
    """The compiled reference functions (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None:
        build()
        if not os.path.exists(_REF_PATH):
            return None
        R = C.CDLL(_REF_PATH)
        for name in ("ref_level_schedule", "ref_graph_coloring"):
            f = getattr(R, name)
            f.argtypes = [C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p]
            f.restype = C.c_int
        R.ref_reorder_matrix.argtypes = [C.c_int, C.c_int, _f64p, _i32p, _i32p, _i32p, _i32p,
                                         _f64p, _i32p, _i32p]
        R.ref_block_mult.argtypes = [_f64p, _f64p, _f64p]
        R.ref_block_mult_sub.argtypes = [_f64p, _f64p, _f64p]
        _ref = R
    return _ref


_ref_ms = None


def ref_mswell_apply(w: "MultisegmentWell", x, y):
    """y -= C^T D^-1 B x by the reference's OWN Opm::MultisegmentWellContribution (constructor + apply, compiled unmodified
    into oracle/_ref/libref_mswell.so; UMFPACK's entry points are a dense-LU stand-in, oracle/ref_mswell_glue.cpp).
    Returns None when that library was never built."""
    global _ref_ms
    if _ref_ms is None:
        build()
        if not os.path.exists(_REF_MSWELL_PATH):
            return None
        _ref_ms = C.CDLL(_REF_MSWELL_PATH)
        _ref_ms.ref_mswell_apply.argtypes = [C.c_uint, C.c_uint, C.c_uint, _f64p, _u32p, _u32p, C.c_uint, _f64p, _i32p, _i32p,
                                             _f64p, _f64p, _f64p]
    y = np.array(y, dtype=np.float64).reshape(-1)
    _ref_ms.ref_mswell_apply(3, 4, int(w.Mb), _c(w.Bvalues, np.float64).reshape(-1), _c(w.BcolIndices, np.uint32),
                             _c(w.BrowPointers, np.uint32), int(w.DnumBlocks), _c(w.Dvalues, np.float64).reshape(-1).copy(),
                             _c(w.DcolPointers, np.int32).copy(), _c(w.DrowIndices, np.int32).copy(),
                             _c(w.Cvalues, np.float64).reshape(-1), _c(x, np.float64).reshape(-1).copy(), y)
    return y


@dataclass
class Wells:
    """Standard wells in the reference's export layout (StandardWellEval.cpp:1202-1251)."""
    val_pointers: np.ndarray   # uint32 [nwells+1]
    Bcols: np.ndarray          # int32 [nblocks]
    Ccols: np.ndarray          # int32 [nblocks]
    B: np.ndarray              # float64 [nblocks, 4, 3]
    C: np.ndarray              # float64 [nblocks, 4, 3]
    Dinv: np.ndarray           # float64 [nwells, 4, 4]

    @property
    def nwells(self) -> int:
        return len(self.val_pointers) - 1


@dataclass
class MultisegmentWell:
    """One multisegment well in the layout MultisegmentWellContribution's constructor takes
    (bda/MultisegmentWellContribution.cpp:32-37): B, C blocked CSR with 4x3 blocks, D scalar CSC."""
    Mb: int
    Bvalues: np.ndarray        # float64 [nblocks, 4, 3]
    BcolIndices: np.ndarray    # uint32 [nblocks]
    BrowPointers: np.ndarray   # uint32 [Mb+1]
    DnumBlocks: int
    Dvalues: np.ndarray        # float64 [16 DnumBlocks]
    DcolPointers: np.ndarray   # int32 [4 Mb + 1]
    DrowIndices: np.ndarray    # int32 [16 DnumBlocks]
    Cvalues: np.ndarray        # float64 [nblocks, 4, 3]


class MSWells:
    """The `multisegments` vector of a WellContributions object (bda/WellContributions.hpp:92)."""

    def __init__(self, wells):
        self.wells = list(wells)
        self._h = lib().orc_ms_create()
        for w in self.wells:
            st = lib().orc_ms_add(self._h, 3, 4, int(w.Mb), _c(w.Bvalues, np.float64).reshape(-1),
                                  _c(w.BcolIndices, np.uint32), _c(w.BrowPointers, np.uint32), int(w.DnumBlocks),
                                  _c(w.Dvalues, np.float64).reshape(-1), _c(w.DcolPointers, np.int32),
                                  _c(w.DrowIndices, np.int32), _c(w.Cvalues, np.float64).reshape(-1))
            if st != 0:
                raise RuntimeError("multisegment well: singular D" if st == 2 else "multisegment well: bad block sizes")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.orc_ms_destroy(h)

    def apply(self, x, y):
        """y -= C^T D^-1 B x over all wells (on a copy, returned).  MultisegmentWellContribution.cpp:70-110."""
        y = np.array(y, dtype=np.float64).reshape(-1)
        lib().orc_ms_apply(self._h, _c(x, np.float64).reshape(-1), y)
        return y


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def check_zero_diagonal(rows, cols, vals) -> int:
    """In place.  BdaBridge.cpp:125-161."""
    assert vals.dtype == np.float64 and vals.flags.c_contiguous
    return lib().orc_check_zero_diagonal(len(rows) - 1, _c(rows, np.int32), _c(cols, np.int32),
                                         vals.reshape(-1))


def spmv(rows, cols, vals, x):
    Nb = len(rows) - 1
    y = np.empty(3 * Nb)
    lib().orc_spmv(Nb, _c(rows, np.int32), _c(cols, np.int32), _c(vals, np.float64).reshape(-1),
                   _c(x, np.float64).reshape(-1), y)
    return y


def well_apply(w: Wells, x, y):
    """y -= C^T D^-1 B x (in place on a copy, returned)."""
    y = np.array(y, dtype=np.float64).reshape(-1).copy()
    lib().orc_well_apply(w.nwells, _c(w.val_pointers, np.uint32), _c(w.Bcols, np.int32),
                         _c(w.Ccols, np.int32), _c(w.B, np.float64).reshape(-1),
                         _c(w.C, np.float64).reshape(-1), _c(w.Dinv, np.float64).reshape(-1),
                         _c(x, np.float64).reshape(-1), y)
    return y


def ilu0(rows, cols, vals, r0=0, r1=None):
    """Returns (LU, diag, status)."""
    Nb = len(rows) - 1
    r1 = Nb if r1 is None else r1
    LU = np.array(vals, dtype=np.float64).reshape(-1).copy()
    diag = np.zeros(Nb, dtype=np.int32)
    st = lib().orc_ilu0_decompose_range(_c(rows, np.int32), _c(cols, np.int32), LU, diag, r0, r1)
    return LU.reshape(-1, 3, 3), diag, st


def ilu0_apply(rows, cols, diag, LU, d, w=1.0, r0=0, r1=None):
    Nb = len(rows) - 1
    r1 = Nb if r1 is None else r1
    v = np.zeros(3 * Nb)
    lib().orc_ilu0_apply_range(_c(rows, np.int32), _c(cols, np.int32), _c(diag, np.int32),
                               _c(LU, np.float64).reshape(-1), _c(d, np.float64).reshape(-1), v,
                               float(w), r0, r1)
    return v


def level_schedule(rows, cols):
    """Returns (toOrder, fromOrder, levelPtr) -- restated Reorder.cpp:266-318."""
    Nb = len(rows) - 1
    to, fr, lp = (np.zeros(Nb, np.int32), np.zeros(Nb, np.int32), np.zeros(Nb + 1, np.int32))
    n = lib().orc_level_schedule(Nb, _c(rows, np.int32), _c(cols, np.int32), to, fr, lp)
    if n < 0:
        raise RuntimeError("level scheduling cannot cover all rows")
    return to, fr, lp[:n + 1].copy()


def ref_level_schedule(rows, cols):
    """The reference's own findLevelScheduling (compiled, unmodified)."""
    R = ref_lib()
    if R is None:
        return None
    Nb = len(rows) - 1
    to, fr, rpc = (np.zeros(Nb, np.int32), np.zeros(Nb, np.int32), np.zeros(max(Nb, 1), np.int32))
    n = R.ref_level_schedule(Nb, _c(rows, np.int32).copy(), _c(cols, np.int32).copy(), to, fr, rpc)
    return to, fr, np.concatenate([[0], np.cumsum(rpc[:n])]).astype(np.int32)


@dataclass
class OracleResult:
    x: np.ndarray
    it: float
    iterations: int
    converged: bool
    breakdown: bool
    reduction: float
    conv_rate: float
    norm0: float
    norm: float
    t_decomp: float
    t_solve: float
    history: np.ndarray


def solve(rows, cols, vals, b, wells: Optional[Wells] = None, tol=1e-10, maxit=200, relaxation=1.0,
          part_ptr=None, threads: Optional[int] = None, mswells: Optional["MSWells"] = None) -> OracleResult:
    """The whole reference CPU path: ILU0 (per partition) + Dune BiCGSTAB (+ wells)."""
    L = lib()
    if threads is not None:
        L.orc_set_threads(int(threads))
    rows = _c(rows, np.int32)
    cols = _c(cols, np.int32)
    vals = _c(vals, np.float64).reshape(-1)
    b = _c(b, np.float64).reshape(-1)
    Nb = len(rows) - 1
    x = np.zeros(3 * Nb)
    res = _OrcResult()
    hist = np.zeros(2 * maxit + 4)
    keep = []
    if wells is not None and wells.nwells > 0:
        keep = [_c(wells.val_pointers, np.uint32), _c(wells.Bcols, np.int32), _c(wells.Ccols, np.int32),
                _c(wells.B, np.float64).reshape(-1), _c(wells.C, np.float64).reshape(-1),
                _c(wells.Dinv, np.float64).reshape(-1)]
        wargs = [wells.nwells] + [_ptr(a) for a in keep]
    else:
        wargs = [0, None, None, None, None, None, None]
    pp = None if part_ptr is None else _c(part_ptr, np.int32)
    L.orc_attach_mswells(mswells._h if mswells is not None else None)
    try:
        st = L.orc_solve(Nb, rows, cols, vals, b, *wargs, float(tol), int(maxit), float(relaxation),
                         0 if pp is None else len(pp) - 1, _ptr(pp), x, C.byref(res),
                         _ptr(hist), len(hist))
    finally:
        L.orc_attach_mswells(None)
    if st != 0:
        raise RuntimeError({1: "diagonal entry missing", 2: "ILU failed to invert matrix block",
                            3: "bad partition"}.get(st, "oracle error %d" % st))
    nh = int(round(2 * res.it)) + 1
    return OracleResult(x, res.it, res.iterations, bool(res.converged), bool(res.breakdown),
                        res.reduction, res.conv_rate, res.norm0, res.norm, res.t_decomp, res.t_solve,
                        hist[:max(nh, 1)].copy())


def true_residual(rows, cols, vals, b, x, wells: Optional[Wells] = None, mswells: Optional["MSWells"] = None) -> float:
    y = spmv(rows, cols, vals, x)
    if mswells is not None:
        y = mswells.apply(x, y)
    if wells is not None and wells.nwells > 0:
        y = well_apply(wells, x, y)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(b - y) / np.linalg.norm(b))


def max_threads() -> int:
    return int(lib().orc_max_threads())
