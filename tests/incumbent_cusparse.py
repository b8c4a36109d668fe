#!/usr/bin/env python
"""Times the reference's own GPU backend (cusparseSolverBackend<3>: cuSPARSE bsrilu02 / bsrsv2 / bsrmv + cuBLAS, compiled
UNMODIFIED from /root/reference into oracle/_ref/libref_cusparse.so, see oracle/Makefile) beside the B200 backend on the
same box, the same system and the same call: solve_system + get_result with host buffers (SURVEY 8d "incumbent GPU
number").  Reported only; the product never loads that library.

  python tests/incumbent_cusparse.py [--workload c3] [--solves 4] [--out gpurun_out/incumbent.json]

Lives under tests/ because it loads oracle/_ref and the oracle (test infrastructure); tests/test_gpu_incumbent.py uses
run_incumbent() to pin small solves -- wells included -- against the reference's own GPU kernels.

Two legs: without wells (both backends solve the identical system; the solutions are cross-checked) and with the
workload's standard wells (the reference's well kernel is launched with 32 threads and only applies the first 10
perforations of a well in its C^T phase, WellContributions.cu:115-124,192, so on wells with more perforations it solves a
different system -- its time is still reported)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_lib():
    p = os.path.join(ROOT, "oracle", "_ref", "libref_cusparse.so")
    if not os.path.exists(p):
        raise SystemExit("oracle/_ref/libref_cusparse.so is missing: run `make -C oracle ref` where /root/reference exists")
    L = C.CDLL(p)
    L.ref_cusparse_last_error.restype = C.c_char_p
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def run_incumbent(L, s, wells, tol, maxit, nsolves):
    N, nnz = 3 * s.Nb, 9 * s.nnzb
    vals = np.ascontiguousarray(s.vals, np.float64).reshape(-1).copy()
    rows = np.ascontiguousarray(s.rows, np.int32)
    # the reference copies nnz (not nnzb) column indices to the device (cusparseSolverBackend.cu:295): it reads 8/9 of that
    # range past the end of the caller's array, which faults for a 1 M-cell system; the harness therefore hands it a
    # column array padded to nnz ints (the reference itself stays unmodified)
    cols = np.zeros(nnz, np.int32)
    cols[:s.nnzb] = s.cols
    b = np.ascontiguousarray(s.b, np.float64).copy()
    x = np.zeros(N)
    wall = np.zeros(nsolves)
    iters = np.zeros(nsolves, np.int32)
    red = np.zeros(nsolves)
    conv = np.zeros(nsolves, np.int32)
    if wells is not None:
        keep = [np.ascontiguousarray(wells.val_pointers, np.uint32), np.ascontiguousarray(wells.Bcols, np.int32),
                np.ascontiguousarray(wells.Ccols, np.int32), np.ascontiguousarray(wells.B, np.float64).reshape(-1),
                np.ascontiguousarray(wells.C, np.float64).reshape(-1), np.ascontiguousarray(wells.Dinv, np.float64).reshape(-1)]
        wargs = [C.c_int(len(keep[0]) - 1)] + [_ptr(a) for a in keep]
    else:
        wargs = [C.c_int(0)] + [None] * 6
    st = L.ref_cusparse_solve(C.c_int(N), C.c_int(nnz), _ptr(vals), _ptr(rows), _ptr(cols), _ptr(b), *wargs, C.c_double(tol),
                              C.c_int(maxit), C.c_int(nsolves), _ptr(x), _ptr(wall), _ptr(iters), _ptr(red), _ptr(conv))
    if st != 0:
        raise RuntimeError("reference cusparse backend: " + L.ref_cusparse_last_error().decode())
    return x, wall, iters, red, conv


def run_b200(s, wells, tol, maxit, nsolves):
    from opm_autodiff_b200 import bridge
    be = bridge.B200SolverBackend(0, maxit, tol, 0)
    res = bridge.BdaResult()
    N, nnz = 3 * s.Nb, 9 * s.nnzb
    x = np.zeros(N)
    wall, its = [], []
    for _ in range(nsolves):
        t0 = time.perf_counter()
        wc = bridge.WellContributions("b200", False) if wells is None else \
            bridge.WellContributions.from_arrays(wells.val_pointers, wells.Bcols, wells.Ccols, wells.B, wells.C, wells.Dinv)
        be.solve_system(N, nnz, 3, s.vals, s.rows, s.cols, s.b, wc, res)
        be.get_result(x)
        wall.append(time.perf_counter() - t0)
        its.append(res.it)
    return x, np.array(wall), its, res.converged


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--solves", type=int, default=4)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--maxit", type=int, default=2000)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import __graft_entry__ as ge
    ge.build()
    from opm_autodiff_b200 import synth
    from oracle import oracle
    L = ref_lib()
    s = synth.full_system(a.workload)

    def rel(u, v):
        return float(np.linalg.norm(u - v) / np.linalg.norm(v))

    out = {"workload": s.cfg.name, "cells": s.Nb, "tolerance": a.tol, "solves": a.solves,
           "what": "wall clock of solve_system + get_result with host buffers, per call; call 0 includes the pattern analysis",
           "legs": {}}
    for leg, wells in (("no_wells", None), ("with_wells", s.wells)):
        if leg == "with_wells" and wells is None:
            continue
        ow = None if wells is None else oracle.Wells(wells.val_pointers, wells.Bcols, wells.Ccols, wells.B, wells.C, wells.Dinv)
        xi, wi, iti, redi, convi = run_incumbent(L, s, wells, a.tol, a.maxit, a.solves)
        xb, wb, itb, convb = run_b200(s, wells, a.tol, a.maxit, a.solves)
        steady_i, steady_b = float(np.mean(wi[1:])), float(np.mean(wb[1:]))
        out["legs"][leg] = {
            "incumbent_cusparse": {"first_call_s": float(wi[0]), "steady_s_per_solve": steady_i, "solves_per_s": 1.0 / steady_i,
                                   "iterations": [int(v) for v in iti], "converged": [int(v) for v in convi],
                                   "reduction": [float(v) for v in redi],
                                   "true_residual": oracle.true_residual(s.rows, s.cols, s.vals, s.b, xi, ow)},
            "b200": {"first_call_s": float(wb[0]), "steady_s_per_solve": steady_b, "solves_per_s": 1.0 / steady_b,
                     "iterations": [float(v) for v in itb], "converged": bool(convb),
                     "true_residual": oracle.true_residual(s.rows, s.cols, s.vals, s.b, xb, ow)},
            "speedup_steady": steady_i / steady_b, "speedup_first_call": float(wi[0] / wb[0]),
            "x_b200_vs_x_incumbent": rel(xb, xi),
        }
    line = json.dumps(out)
    print(line)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(line + "\n")


if __name__ == "__main__":
    main()
