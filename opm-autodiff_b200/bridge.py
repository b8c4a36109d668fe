"""Host-side mirror of the reference's accelerator-bridge interface for the B200 backend.

Same names, argument meaning and error behaviour as the reference classes, bound to the C ABI of
``libb200bda.so`` (``include/b200bda.h``) with ctypes:

  ``SolverStatus``            bda::SolverStatus            bda/BdaSolver.hpp:32-37
  ``BdaResult``               bda::BdaResult               bda/BdaResult.hpp:28-40
  ``WellContributions``       Opm::WellContributions       bda/WellContributions.hpp:60-214, .cpp:31-259
  ``B200SolverBackend``       bda::BdaSolver<3> subclass   bda/BdaSolver.hpp:86-90 (cf. cusparseSolverBackend)
  ``BdaBridge``               Opm::BdaBridge<M,V,3>        bda/BdaBridge.cpp:56-121,192-263

There is no CPU fallback: if the CUDA library is missing or no sm_100 device is visible, the
constructors raise.
"""
from __future__ import annotations

import ctypes as C
import enum
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200bda.so")
CSRC = os.path.join(_HERE, "csrc")


class SolverStatus(enum.IntEnum):
    BDA_SOLVER_SUCCESS = 0
    BDA_SOLVER_ANALYSIS_FAILED = 1
    BDA_SOLVER_CREATE_PRECONDITIONER_FAILED = 2
    BDA_SOLVER_UNKNOWN_ERROR = 3


class _CResult(C.Structure):
    _fields_ = [("iterations", C.c_int), ("reduction", C.c_double), ("converged", C.c_int),
                ("conv_rate", C.c_double), ("elapsed", C.c_double), ("it", C.c_double),
                ("norm0", C.c_double), ("norm", C.c_double), ("breakdown", C.c_int),
                ("num_levels", C.c_int), ("t_analysis", C.c_double), ("t_copy", C.c_double),
                ("t_factor", C.c_double), ("t_krylov", C.c_double)]


@dataclass
class BdaResult:
    """bda/BdaResult.hpp:32-36 (first five fields) + extras filled by this backend."""
    iterations: int = 0
    reduction: float = 0.0
    converged: bool = False
    conv_rate: float = 0.0
    elapsed: float = 0.0
    it: float = 0.0
    norm0: float = 0.0
    norm: float = 0.0
    breakdown: bool = False
    num_levels: int = 0
    t_analysis: float = 0.0
    t_copy: float = 0.0
    t_factor: float = 0.0
    t_krylov: float = 0.0

    def _fill(self, r: _CResult) -> None:
        for name, _ in _CResult._fields_:
            v = getattr(r, name)
            setattr(self, name, bool(v) if name in ("converged", "breakdown") else v)


def build(force: bool = False) -> None:
    """Compile libb200bda.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "b200bda.h"))
    newest = max(os.path.getmtime(f) for f in srcs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", CSRC], stdout=subprocess.DEVNULL)


_lib = None

_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def lib():
    """The loaded C-ABI library.  Raises if it was never built: no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libb200bda.so is not built (run __graft_entry__.build()); "
                               "this backend has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        vp, ip = C.c_void_p, C.c_int
        L.b200_create.argtypes = [ip, ip, C.c_double, C.c_uint]
        L.b200_create.restype = vp
        L.b200_destroy.argtypes = [vp]
        L.b200_set_option.argtypes = [vp, C.c_char_p, C.c_double]
        L.b200_solve_system.argtypes = [vp, ip, ip, ip, vp, vp, vp, vp, vp, C.POINTER(_CResult)]
        L.b200_upload_system.argtypes = [vp, ip, ip, ip, vp, vp, vp, vp, vp]
        L.b200_solve_resident.argtypes = [vp, C.POINTER(_CResult)]
        L.b200_get_result.argtypes = [vp, vp]
        L.b200_last_error.restype = C.c_char_p
        L.b200_version.restype = C.c_char_p
        L.b200_wells_create.argtypes = [C.c_char_p, ip]
        L.b200_wells_create.restype = vp
        L.b200_wells_destroy.argtypes = [vp]
        L.b200_wells_set_block_size.argtypes = [vp, C.c_uint, C.c_uint]
        L.b200_wells_add_num_blocks.argtypes = [vp, C.c_uint]
        L.b200_wells_alloc.argtypes = [vp]
        L.b200_wells_add_matrix.argtypes = [vp, ip, vp, vp, C.c_uint]
        L.b200_wells_get_num_wells.argtypes = [vp]
        L.b200_wells_get_num_wells.restype = C.c_uint
        L.b200_wells_add_multisegment.argtypes = [vp, C.c_uint, C.c_uint, C.c_uint, vp, vp, vp, C.c_uint, vp, vp, vp, vp]
        L.b200_wells_get_multisegment_inverse.argtypes = [vp, C.c_uint, vp]
        L.b200_spmv.argtypes = [vp, _f64p, _f64p]
        L.b200_well_apply.argtypes = [vp, _f64p, _f64p]
        L.b200_ilu0_factorize.argtypes = [vp]
        L.b200_ilu0_apply.argtypes = [vp, _f64p, _f64p]
        L.b200_get_ilu0.argtypes = [vp, _f64p]
        L.b200_get_level_schedule.argtypes = [vp, _i32p, _i32p, _i32p, C.POINTER(C.c_int)]
        L.b200_level_schedule_host.argtypes = [ip, _i32p, _i32p, _i32p, _i32p, _i32p, C.POINTER(C.c_int)]
        L.b200_sweep_schedule_check_host.argtypes = [ip, _i32p, _i32p, ip, ip, ip, C.c_uint, C.POINTER(C.c_double),
                                                     np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")]
        L.b200_sweep2_schedule_check_host.argtypes = [ip, _i32p, _i32p, ip, ip, ip, ip, ip, ip, ip, C.c_uint, C.c_double,
                                                      C.POINTER(C.c_double), np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")]
        L.b200_graph_coloring_host.argtypes = [ip, _i32p, _i32p, C.c_uint, _i32p, _i32p, _i32p, C.POINTER(C.c_int)]
        L.b200_get_reorder.argtypes = [vp, _i32p, _i32p, C.POINTER(C.c_int)]
        L.b200_host_register.argtypes = [vp, vp, C.c_size_t]
        L.b200_host_unregister.argtypes = [vp, vp]
        L.b200_time_kernel.argtypes = [vp, C.c_char_p, ip, ip, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.b200_kernel_stats.argtypes = [vp, C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_double),
                                        C.POINTER(C.c_double)]
        L.b200_reset_stats.argtypes = [vp]
        L.b200_launch_count.argtypes = [vp]
        L.b200_launch_count.restype = C.c_longlong
        L.b200_timer_start.argtypes = [vp]
        L.b200_timer_stop.argtypes = [vp, C.POINTER(C.c_double)]
        L.b200_device_available.restype = ip
        L.b200_factor_plan_check_host.argtypes = [ip, _i32p, _i32p, _f64p, _f64p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.b200_get_sweep_trace.argtypes = [vp, np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS"), C.c_longlong]
        L.b200_dist_unique_id.argtypes = [vp]
        L.b200_dist_init.argtypes = [vp, ip, ip, vp]
        L.b200_dist_set_halo.argtypes = [vp, ip, ip, vp, vp, vp, vp, vp]
        L.b200_dist_map_rank.argtypes = [vp, ip, vp]
        L.b200_dist_connect_peer.argtypes = [vp, ip, ip, ip, ip]
        L.b200_dist_spmv.argtypes = [vp, _f64p, _f64p]
        L.b200_dist_rank.argtypes = [vp]
        L.b200_dist_world.argtypes = [vp]
        _lib = L
    return _lib


EXPORTED_SYMBOLS = [
    "b200_create", "b200_destroy", "b200_set_option", "b200_solve_system", "b200_get_result",
    "b200_upload_system", "b200_solve_resident", "b200_last_error", "b200_wells_create",
    "b200_wells_destroy", "b200_wells_set_block_size", "b200_wells_add_num_blocks", "b200_wells_alloc",
    "b200_wells_add_matrix", "b200_wells_get_num_wells", "b200_wells_add_multisegment", "b200_wells_get_multisegment_inverse", "b200_spmv", "b200_well_apply",
    "b200_ilu0_factorize", "b200_ilu0_apply", "b200_get_ilu0", "b200_get_level_schedule",
    "b200_level_schedule_host", "b200_sweep_schedule_check_host", "b200_sweep2_schedule_check_host", "b200_graph_coloring_host", "b200_get_reorder", "b200_host_register", "b200_host_unregister", "b200_time_kernel", "b200_kernel_stats", "b200_reset_stats",
    "b200_launch_count", "b200_timer_start", "b200_timer_stop", "b200_device_available", "b200_version",
    "b200_dist_unique_id", "b200_dist_init", "b200_dist_set_halo", "b200_dist_map_rank", "b200_dist_connect_peer", "b200_dist_spmv",
    "b200_dist_rank", "b200_dist_world", "b200_get_sweep_trace", "b200_factor_plan_check_host",
]


def last_error() -> str:
    return lib().b200_last_error().decode()


def device_available() -> bool:
    return bool(lib().b200_device_available())


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class WellContributions:
    """Opm::WellContributions.  Standard wells in three phases, as the reference:
    setBlockSize + addNumBlocks per well, alloc, then per well addMatrix C, D, B
    (bda/WellContributions.cpp:152-259; fill order wells/StandardWellEval.cpp:1202-1251);
    multisegment wells one call each (addMultisegmentWellContribution, .cpp:261-271)."""

    class MatrixType(enum.IntEnum):
        C = 0
        D = 1
        B = 2

    def __init__(self, accelerator_mode: str, useWellConn: bool):
        self._h = lib().b200_wells_create(accelerator_mode.encode(), int(bool(useWellConn)))
        if not self._h:
            raise ValueError(last_error())       # std::logic_error("Invalid accelerator mode")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.b200_wells_destroy(h)

    def _check(self, st):
        if st != 0:
            raise ValueError(last_error())       # OPM_THROW(std::logic_error, ...)

    def setBlockSize(self, dim: int, dim_wells: int) -> None:
        self._check(lib().b200_wells_set_block_size(self._h, dim, dim_wells))

    def addNumBlocks(self, numBlocks: int) -> None:
        self._check(lib().b200_wells_add_num_blocks(self._h, numBlocks))

    def alloc(self) -> None:
        self._check(lib().b200_wells_alloc(self._h))

    def addMatrix(self, type, colIndices, values, val_size: int) -> None:
        ci = None if colIndices is None else np.ascontiguousarray(colIndices, dtype=np.int32)
        va = np.ascontiguousarray(values, dtype=np.float64)
        self._check(lib().b200_wells_add_matrix(self._h, int(type), _ptr(ci), _ptr(va), int(val_size)))

    def getNumWells(self) -> int:
        return int(lib().b200_wells_get_num_wells(self._h))

    def addMultisegmentWellContribution(self, dim, dim_wells, Mb, Bvalues, BcolIndices, BrowPointers, DnumBlocks, Dvalues,
                                        DcolPointers, DrowIndices, Cvalues) -> None:
        """bda/WellContributions.hpp:195-213: B, C blocked CSR (dim_wells x dim blocks), D scalar CSC (UMFPACK layout)."""
        Bv = np.ascontiguousarray(Bvalues, dtype=np.float64)
        Bc = np.ascontiguousarray(BcolIndices, dtype=np.uint32)
        Br = np.ascontiguousarray(BrowPointers, dtype=np.uint32)
        Dv = np.ascontiguousarray(Dvalues, dtype=np.float64)
        Dc = np.ascontiguousarray(DcolPointers, dtype=np.int32)
        Dr = np.ascontiguousarray(DrowIndices, dtype=np.int32)
        Cv = np.ascontiguousarray(Cvalues, dtype=np.float64)
        if (dim, dim_wells) == (3, 4) and (len(Br) != Mb + 1 or len(Bc) < Br[-1] or Bv.size < 12 * Br[-1] or Cv.size < 12 * Br[-1] or len(Dc) != dim_wells * Mb + 1 \
                or len(Dr) < Dc[-1] or Dv.size < Dc[-1]):
            raise ValueError("addMultisegmentWellContribution: array sizes do not match Mb / the row and column pointers")
        self._check(lib().b200_wells_add_multisegment(self._h, int(dim), int(dim_wells), int(Mb), _ptr(Bv), _ptr(Bc), _ptr(Br),
                                                      int(DnumBlocks), _ptr(Dv), _ptr(Dc), _ptr(Dr), _ptr(Cv)))

    def multisegment_inverse(self, index: int, M: int) -> np.ndarray:
        """Test hook: the dense M x M inverse of D the library holds for multisegment well `index`."""
        out = np.empty((M, M))
        self._check(lib().b200_wells_get_multisegment_inverse(self._h, int(index), _ptr(out)))
        return out

    @classmethod
    def from_arrays(cls, val_pointers, Bcols, Ccols, B, C, Dinv, accelerator_mode="b200"):
        """Convenience: fill from CSR-over-wells arrays exactly as BlackoilWellModel::getWellContributions
        (wells/BlackoilWellModel_impl.hpp:1061-1099) would."""
        w = cls(accelerator_mode, False)
        nw = len(val_pointers) - 1
        if nw == 0:
            return w
        w.setBlockSize(3, 4)
        for i in range(nw):
            w.addNumBlocks(int(val_pointers[i + 1] - val_pointers[i]))
        w.alloc()
        B = np.asarray(B, dtype=np.float64).reshape(-1, 12)
        Cm = np.asarray(C, dtype=np.float64).reshape(-1, 12)
        Dinv = np.asarray(Dinv, dtype=np.float64).reshape(-1, 16)
        for i in range(nw):
            s, e = int(val_pointers[i]), int(val_pointers[i + 1])
            w.addMatrix(cls.MatrixType.C, Ccols[s:e], Cm[s:e], e - s)
            w.addMatrix(cls.MatrixType.D, np.zeros(1, np.int32), Dinv[i], 1)
            w.addMatrix(cls.MatrixType.B, Bcols[s:e], B[s:e], e - s)
        return w


class B200SolverBackend:
    """bda::BdaSolver<3> implementation backed by libb200bda.so (cf. cusparseSolverBackend<3>)."""

    def __init__(self, linear_solver_verbosity: int, maxit: int, tolerance: float, deviceID: int = 0):
        self.verbosity, self.maxit, self.tolerance, self.deviceID = linear_solver_verbosity, maxit, tolerance, deviceID
        self._h = lib().b200_create(int(linear_solver_verbosity), int(maxit), float(tolerance), int(deviceID))
        if not self._h:
            raise RuntimeError(last_error())     # OPM_THROW(std::logic_error) in cuda_header.hpp:34-44
        self.N = 0
        self._keep = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib is not None:
            _lib.b200_destroy(h)

    def set_option(self, key: str, value: float) -> None:
        if lib().b200_set_option(self._h, key.encode(), float(value)) != 0:
            raise ValueError(last_error())

    @staticmethod
    def _arrs(vals, rows, cols, b):
        vals = np.ascontiguousarray(vals, dtype=np.float64)
        rows = None if rows is None else np.ascontiguousarray(rows, dtype=np.int32)
        cols = None if cols is None else np.ascontiguousarray(cols, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.float64)
        return vals, rows, cols, b

    def solve_system(self, N, nnz, dim, vals, rows, cols, b, wellContribs: Optional[WellContributions],
                     res: BdaResult) -> SolverStatus:
        vals, rows, cols, b = self._arrs(vals, rows, cols, b)
        self._keep = (vals, rows, cols, b)       # pinned by the library until the next call
        r = _CResult()
        st = lib().b200_solve_system(self._h, int(N), int(nnz), int(dim), _ptr(vals), _ptr(rows), _ptr(cols),
                                     _ptr(b), wellContribs._h if wellContribs is not None else None, C.byref(r))
        res._fill(r)
        self.N = int(N)
        if st == SolverStatus.BDA_SOLVER_UNKNOWN_ERROR:
            raise RuntimeError(last_error())     # device/runtime errors throw in the reference
        return SolverStatus(st)

    def get_result(self, x: np.ndarray) -> None:
        assert x.dtype == np.float64 and x.flags.c_contiguous and x.size >= self.N
        self._keep_x = x                         # page-locked by the library until the next call
        if lib().b200_get_result(self._h, _ptr(x)) != 0:
            raise RuntimeError(last_error())

    # -- device-resident leg and kernel-level entry points (benchmark / parity tests) --
    def upload_system(self, N, nnz, dim, vals, rows, cols, b, wellContribs=None) -> None:
        vals, rows, cols, b = self._arrs(vals, rows, cols, b)
        self._keep = (vals, rows, cols, b)
        st = lib().b200_upload_system(self._h, int(N), int(nnz), int(dim), _ptr(vals), _ptr(rows), _ptr(cols),
                                      _ptr(b), wellContribs._h if wellContribs is not None else None)
        self.N = int(N)
        if st != 0:
            raise RuntimeError(last_error())

    def solve_resident(self, res: BdaResult) -> SolverStatus:
        r = _CResult()
        st = lib().b200_solve_resident(self._h, C.byref(r))
        res._fill(r)
        if st == SolverStatus.BDA_SOLVER_UNKNOWN_ERROR:
            raise RuntimeError(last_error())
        return SolverStatus(st)

    def _chk(self, st):
        if st != 0:
            raise RuntimeError(last_error())

    def spmv(self, x):
        y = np.empty(self.N)
        self._chk(lib().b200_spmv(self._h, np.ascontiguousarray(x, dtype=np.float64).reshape(-1), y))
        return y

    def well_apply(self, x, y):
        y = np.array(y, dtype=np.float64).reshape(-1).copy()
        self._chk(lib().b200_well_apply(self._h, np.ascontiguousarray(x, dtype=np.float64).reshape(-1), y))
        return y

    def host_register(self, arr: np.ndarray) -> None:
        """Page-lock a caller-owned array (matrix values, right-hand side, solution vector) for the copies of solve_system /
        get_result; the caller keeps it alive until host_unregister or the solver is destroyed (b200_host_register)."""
        assert arr.flags["C_CONTIGUOUS"]
        self._chk(lib().b200_host_register(self._h, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def host_unregister(self, arr: np.ndarray) -> None:
        self._chk(lib().b200_host_unregister(self._h, arr.ctypes.data_as(C.c_void_p)))

    def ilu0_factorize(self) -> SolverStatus:
        st = lib().b200_ilu0_factorize(self._h)
        if st == SolverStatus.BDA_SOLVER_UNKNOWN_ERROR:
            raise RuntimeError(last_error())
        return SolverStatus(st)

    def ilu0_apply(self, d):
        v = np.empty(self.N)
        self._chk(lib().b200_ilu0_apply(self._h, np.ascontiguousarray(d, dtype=np.float64).reshape(-1), v))
        return v

    def get_ilu0(self, nnzb: int):
        lu = np.empty(nnzb * 9)
        self._chk(lib().b200_get_ilu0(self._h, lu))
        return lu.reshape(-1, 3, 3)

    def get_level_schedule(self):
        Nb = self.N // 3
        to, fr, rpl = np.zeros(Nb, np.int32), np.zeros(Nb, np.int32), np.zeros(Nb, np.int32)
        n = C.c_int(0)
        self._chk(lib().b200_get_level_schedule(self._h, to, fr, rpl, C.byref(n)))
        return to, fr, rpl[:n.value].copy()

    def get_reorder(self):
        """(toOrder, fromOrder, colours) of a solver created with option reorder = 1, after its first solve."""
        Nb = self.N // 3
        to, fr = np.zeros(Nb, np.int32), np.zeros(Nb, np.int32)
        n = C.c_int(0)
        self._chk(lib().b200_get_reorder(self._h, to, fr, C.byref(n)))
        return to, fr, n.value

    def time_kernel(self, which: str, reps: int = 20, flush_l2: bool = True):
        ms, by = C.c_double(0), C.c_double(0)
        self._chk(lib().b200_time_kernel(self._h, which.encode(), int(reps), int(flush_l2), C.byref(ms), C.byref(by)))
        return ms.value, by.value

    def kernel_stats(self, which: str):
        n, ms, by = C.c_longlong(0), C.c_double(0), C.c_double(0)
        self._chk(lib().b200_kernel_stats(self._h, which.encode(), C.byref(n), C.byref(ms), C.byref(by)))
        return n.value, ms.value, by.value

    def reset_stats(self) -> None:
        lib().b200_reset_stats(self._h)

    def launch_count(self) -> int:
        return int(lib().b200_launch_count(self._h))

    # -- multi-GPU (one backend per rank; see include/b200bda.h and dist.py) --
    @staticmethod
    def dist_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        if lib().b200_dist_unique_id(buf) != 0:
            raise RuntimeError(last_error())
        return buf.raw

    def dist_init(self, rank: int, world: int, unique_id: Optional[bytes]) -> None:
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._chk(lib().b200_dist_init(self._h, int(rank), int(world), buf))

    def dist_set_halo(self, n_ghost, neigh_rank, send_ptr, send_rows, recv_ptr) -> bytes:
        a = [np.ascontiguousarray(v, dtype=np.int32) for v in (neigh_rank, send_ptr, send_rows, recv_ptr)]
        out = C.create_string_buffer(64)
        self._chk(lib().b200_dist_set_halo(self._h, int(n_ghost), len(a[0]), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), out))
        return out.raw

    def dist_map_rank(self, rank: int, handle: Optional[bytes]) -> None:
        buf = C.create_string_buffer(handle, 64) if handle is not None else None
        self._chk(lib().b200_dist_map_rank(self._h, int(rank), buf))

    def dist_connect_peer(self, neigh_index, peer_n_ghost, peer_recv_offset, peer_slot) -> None:
        self._chk(lib().b200_dist_connect_peer(self._h, int(neigh_index), int(peer_n_ghost), int(peer_recv_offset), int(peer_slot)))

    def dist_spmv(self, x):
        y = np.empty(self.N)
        self._chk(lib().b200_dist_spmv(self._h, np.ascontiguousarray(x, dtype=np.float64).reshape(-1), y))
        return y

    def sweep_trace(self):
        """[148, 1024, 4] SM-clock stamps of the last traced sweep (option sweep_trace = 1)."""
        out = np.zeros(3 * 148 * 1024 * 4, np.int64)
        self._chk(lib().b200_get_sweep_trace(self._h, out, out.size))
        return out.reshape(3, 148, 1024, 4)

    def timer_start(self) -> None:
        self._chk(lib().b200_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double(0)
        self._chk(lib().b200_timer_stop(self._h, C.byref(ms)))
        return ms.value


def level_schedule_host(rows, cols):
    """Host-only level scheduling of a pattern (no device).  bda/Reorder.cpp:266-318."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    Nb = len(rows) - 1
    to, fr, rpl = np.zeros(Nb, np.int32), np.zeros(Nb, np.int32), np.zeros(Nb, np.int32)
    n = C.c_int(0)
    if lib().b200_level_schedule_host(Nb, rows, cols, to, fr, rpl, C.byref(n)) != 0:
        raise RuntimeError(last_error())
    return to, fr, rpl[:n.value].copy()


SWEEP_STAT_NAMES = ("parts", "lines", "strips", "stages_L", "stages_U", "chunks_L", "window_deps_L", "global_deps_L",
                    "max_meta_ints", "max_vals_doubles", "max_rhs_rows", "levels")


def sweep_schedule_check_host(rows, cols, parts=0, stage_bytes=0, window=0, seed=1):
    """Host-only emulation of the triangular-sweep schedule of a pattern (no device): returns the largest
    error against the sequential natural-order substitution (relative to max|x|) and the schedule statistics."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    err = C.c_double(0)
    stats = np.zeros(12, np.int64)
    if lib().b200_sweep_schedule_check_host(len(rows) - 1, rows, cols, parts, stage_bytes, window, seed, C.byref(err), stats) != 0:
        raise RuntimeError(last_error())
    return err.value, dict(zip(SWEEP_STAT_NAMES, (int(v) for v in stats)))


SWEEP2_STAT_NAMES = ("parts", "lines", "strips", "records_L", "records_U", "multi_record_warp_steps_L", "window_deps_L", "external_deps_L",
                     "multi_lane_records_L", "mutual_part_pairs_L", "max_chunks_per_step", "lanes_L")


def sweep2_schedule_check_host(rows, cols, parts=0, window=0, ext_window=0, consumer_warps=0, helpers=0, groups=0, wg=0, seed=1,
                               relax=1.0):
    """Host-only emulation of the round-2 sweep schedule (groups of consumer warps streaming their own records): returns the
    largest error against the sequential natural-order substitution (relative to max|x|) and the schedule statistics."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    err = C.c_double(0)
    stats = np.zeros(12, np.int64)
    if lib().b200_sweep2_schedule_check_host(len(rows) - 1, rows, cols, parts, window, ext_window, consumer_warps, helpers, groups, wg,
                                             seed, relax, C.byref(err), stats) != 0:
        raise RuntimeError(last_error())
    return err.value, dict(zip(SWEEP2_STAT_NAMES, (int(v) for v in stats)))


def graph_coloring_host(rows, cols, seed=1):
    """Host-only multi-colour ordering (coloring.hpp, restating Reorder.cpp:58-172,209-222): returns (toOrder, fromOrder,
    rowsPerColor)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    Nb = len(rows) - 1
    to, fr, rpc = np.zeros(Nb, np.int32), np.zeros(Nb, np.int32), np.zeros(256, np.int32)
    nc = C.c_int(0)
    if lib().b200_graph_coloring_host(Nb, rows, cols, seed, to, fr, rpc, C.byref(nc)) != 0:
        raise RuntimeError(last_error())
    return to, fr, rpc[:nc.value].copy()


def factor_plan_check_host(rows, cols, vals):
    """Host-only replay of the device's ILU0 elimination plan: returns (LU [nnzb,3,3], longest row, longest plan)."""
    rows = np.ascontiguousarray(rows, dtype=np.int32)
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64).reshape(-1)
    lu = np.zeros_like(vals)
    mr, mo = C.c_int(0), C.c_int(0)
    st = lib().b200_factor_plan_check_host(len(rows) - 1, rows, cols, vals, lu, C.byref(mr), C.byref(mo))
    if st != 0:
        raise RuntimeError(last_error())
    return lu.reshape(-1, 3, 3), mr.value, mo.value


@dataclass
class BsrMatrix:
    """What BdaBridge sees of a Dune::BCRSMatrix<MatrixBlock<double,3,3>>: N() block rows, the
    contiguous row-major block array and (once) the sparsity pattern (BdaBridge.cpp:167-189)."""
    rows: np.ndarray     # int32 [Nb+1]
    cols: np.ndarray     # int32 [nnzb]
    vals: np.ndarray     # float64 [nnzb,3,3]

    def __post_init__(self):
        self.rows = np.ascontiguousarray(self.rows, dtype=np.int32)
        self.cols = np.ascontiguousarray(self.cols, dtype=np.int32)
        self.vals = np.ascontiguousarray(self.vals, dtype=np.float64).reshape(-1, 3, 3)

    def N(self) -> int:
        return len(self.rows) - 1

    def nonzeroes(self) -> int:
        return int(self.rows[-1])

    @property
    def block_dim(self) -> int:
        return self.vals.shape[1]


@dataclass
class FlexibleSolverOptions:
    """The part of a FlexibleSolver property tree this backend honours (keys and defaults as
    FlexibleSolver_impl.hpp:147-150 and setupPropertyTree.cpp:175-188 `setupILU`): ``tol``, ``maxiter``, ``verbosity``,
    ``solver`` (must be ``bicgstab``), ``preconditioner.type`` (``ILU0`` / ``ParOverILU0``), ``preconditioner.relaxation``,
    ``preconditioner.ilulevel`` (must be 0).  JSON values may be strings, as in tests/options_flexiblesolver.json."""
    tol: float = 1e-2
    maxiter: int = 200
    verbosity: int = 0
    relaxation: float = 1.0

    @classmethod
    def from_tree(cls, prm: dict, strict: bool = True) -> "FlexibleSolverOptions":
        o = cls(float(prm.get("tol", 1e-2)), int(prm.get("maxiter", 200)), int(prm.get("verbosity", 0)))
        solver = str(prm.get("solver", "bicgstab"))
        pre = prm.get("preconditioner", {}) or {}
        ptype = str(pre.get("type", "ParOverILU0"))
        if strict:
            if solver != "bicgstab":
                raise ValueError("the b200 backend implements solver 'bicgstab' only, got '%s'" % solver)
            if ptype.lower() not in ("ilu0", "paroverilu0"):
                raise ValueError("the b200 backend implements preconditioner ILU0 / ParOverILU0 only, got '%s'" % ptype)
            if int(pre.get("ilulevel", 0)) != 0:
                raise ValueError("the b200 backend implements fill level 0 only")
        if ptype.lower() in ("ilu0", "paroverilu0"):
            o.relaxation = float(pre.get("relaxation", 1.0))
        return o

    @classmethod
    def from_json(cls, path: str, strict: bool = True) -> "FlexibleSolverOptions":
        import json
        with open(path) as f:
            return cls.from_tree(json.load(f), strict)


@dataclass
class InverseOperatorResult:
    """Dune::InverseOperatorResult fields BdaBridge fills (BdaBridge.cpp:247-251)."""
    iterations: int = 0
    reduction: float = 0.0
    converged: bool = False
    conv_rate: float = 0.0
    elapsed: float = 0.0


class BdaBridge:
    """Opm::BdaBridge<BridgeMatrix, BridgeVector, 3> restricted to the new accelerator mode.

    Constructor signature as BdaBridge.cpp:56-63; accepts accelerator_mode "b200" (the branch a
    maintainer adds next to "cusparse"/"opencl") or "none".  Any other mode
    raises like the reference's final else branch (:118-120)."""

    def __init__(self, accelerator_mode: str, fpga_bitstream: str = "", linear_solver_verbosity: int = 0,
                 maxit: int = 200, tolerance: float = 1e-2, platformID: int = 0, deviceID: int = 0,
                 opencl_ilu_reorder: str = "none"):
        self.verbosity = linear_solver_verbosity
        self.accelerator_mode = accelerator_mode
        self.use_gpu = False
        self.backend = None
        self._h_rows = None
        self._h_cols = None
        self._diag_indices = None
        self.last_result = BdaResult()
        if opencl_ilu_reorder not in ("none", "level_scheduling", "graph_coloring"):
            # BdaBridge.cpp:72-80
            raise ValueError("Error invalid argument for --opencl-ilu-reorder, usage: '--opencl-ilu-reorder=[level_scheduling|graph_coloring]'")
        if accelerator_mode == "b200":
            self.use_gpu = True
            self.backend = B200SolverBackend(linear_solver_verbosity, maxit, tolerance, deviceID)
            # natural-order ILU0 ("none" / "level_scheduling": the level sets only schedule the same factorisation) is the default;
            # "graph_coloring" is the opt-in colour ordering of the reference's OpenCL backend (another preconditioner)
            if opencl_ilu_reorder == "graph_coloring":
                self.backend.set_option("reorder", 1)
        elif accelerator_mode == "none":
            self.use_gpu = False
        else:
            raise ValueError("Error unknown value for parameter 'AcceleratorMode', should be passed like "
                             "'--accelerator-mode=[none|cusparse|opencl|fpga|amgcl|b200]")

    @classmethod
    def from_flexible_solver_options(cls, accelerator_mode: str, options: "FlexibleSolverOptions", deviceID: int = 0) -> "BdaBridge":
        """Bridge configured from a FlexibleSolver property tree, the way ISTLSolverEbos passes
        linear_solver_reduction / maxiter / verbosity to BdaBridge (ISTLSolverEbos.hpp:133-150); the ILU relaxation,
        which the reference's GPU backends hard-wire to 1, is honoured here."""
        br = cls(accelerator_mode, "", options.verbosity, options.maxiter, options.tol, 0, deviceID, "none")
        if br.backend is not None:
            br.backend.set_option("relaxation", options.relaxation)
        return br

    def getUseGpu(self) -> bool:
        return self.use_gpu

    def checkZeroDiagonal(self, mat: BsrMatrix) -> int:
        """BdaBridge.cpp:125-161: exact zeros on the diagonal of diagonal blocks become 1e-15 in the
        CALLER's matrix; diagonal offsets cached on the first call."""
        if self._diag_indices is None:
            Nb = mat.N()
            rowid = np.repeat(np.arange(Nb, dtype=np.int64), np.diff(mat.rows))
            idx = np.nonzero(mat.cols == rowid)[0]
            if len(idx) != Nb:
                raise AssertionError("diagonal block missing")
            self._diag_indices = idx
        d = mat.vals[self._diag_indices]
        ii = np.arange(3)
        dd = d[:, ii, ii]
        zeros = dd == 0.0
        n = int(zeros.sum())
        if n:
            dd[zeros] = 1e-15
            d[:, ii, ii] = dd
            mat.vals[self._diag_indices] = d
        return n

    def solve_system(self, mat: BsrMatrix, b: np.ndarray, wellContribs: Optional[WellContributions],
                     res: InverseOperatorResult) -> None:
        """BdaBridge.cpp:192-255."""
        if not self.use_gpu:
            res.converged = False
            return
        result = BdaResult()
        result.converged = False
        dim = mat.block_dim
        Nb = mat.N()
        N = Nb * dim
        nnzb = mat.nonzeroes() if self._h_rows is None else int(self._h_rows[-1])
        nnz = nnzb * dim * dim
        if dim != 3:
            import warnings
            warnings.warn("BdaSolver only accepts blocksize = 3 at this time, will use Dune for the remainder of the program")
            self.use_gpu = False
            return
        if self._h_rows is None:
            self._h_rows, self._h_cols = mat.rows.copy(), mat.cols.copy()
            if int(self._h_rows[Nb]) != mat.nonzeroes():
                raise ValueError("Error size of rows do not sum to number of nonzeroes in BdaBridge::getSparsityPattern()")
        self.checkZeroDiagonal(mat)
        status = self.backend.solve_system(N, nnz, dim, mat.vals, self._h_rows, self._h_cols, b, wellContribs, result)
        if status != SolverStatus.BDA_SOLVER_SUCCESS:
            import warnings
            warnings.warn({SolverStatus.BDA_SOLVER_ANALYSIS_FAILED:
                           "BdaSolver could not analyse level information of matrix, perhaps there is still a 0.0 on the diagonal of a block on the diagonal",
                           SolverStatus.BDA_SOLVER_CREATE_PRECONDITIONER_FAILED:
                           "BdaSolver could not create preconditioner, perhaps there is still a 0.0 on the diagonal of a block on the diagonal"}
                          .get(status, "BdaSolver returned unknown status code"))
        res.iterations = result.iterations
        res.reduction = result.reduction
        res.converged = result.converged
        res.conv_rate = result.conv_rate
        res.elapsed = result.elapsed
        self.last_result = result

    def get_result(self, x: np.ndarray) -> None:
        """BdaBridge.cpp:258-263."""
        if self.use_gpu:
            self.backend.get_result(x)
