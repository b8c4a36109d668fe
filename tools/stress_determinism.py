#!/usr/bin/env python
"""Race hunting without a sanitizer (GPU box tool): many solves of systems of different shapes with every size-dependent
feature forced on and several sweep geometries; every repeat must reproduce the first solution bit for bit and match the
solution with the features off."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
bad = 0
for shape, nw, opts in (((40, 30, 20), 4, {}), ((33, 17, 29), 0, {"sweep_parts": 37}), ((60, 50, 40), 8, {"sweep_stage_bytes": 16384}),
                        ((25, 25, 25), 5, {"sweep_parts": 9, "fuse_unit_slices": 1}), ((100, 100, 12), 6, {})):
    s = synth.small(*shape, nwells=nw, nperf=6)
    w = s.wells
    wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv) if w is not None else None
    sols = {}
    for feat in (2, 0):
        be = bridge.B200SolverBackend(0, 300, 1e-10, 0)
        for k in ("spmv_sell", "fuse_spmv", "defer_x", "sweep_early"):
            be.set_option(k, feat)
        for k, v in opts.items():
            be.set_option(k, v)
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
        res = bridge.BdaResult()
        first = None
        for rep in range(25 if feat else 3):
            be.solve_resident(res)
            x = np.zeros(3 * s.Nb); be.get_result(x)
            if first is None:
                first = x.copy(); it0 = res.it
            elif not (np.array_equal(x, first) and res.it == it0):
                bad += 1
                print("NOT REPRODUCED: shape", shape, "features", feat, "repeat", rep, "max diff", np.max(np.abs(x - first)), flush=True)
        sols[feat] = (first, it0)
        del be
    d = np.linalg.norm(sols[2][0] - sols[0][0]) / np.linalg.norm(sols[0][0])
    e = np.linalg.norm(sols[2][0] - s.x_true) / np.linalg.norm(s.x_true)
    print("shape %s: iterations %.1f / %.1f, features on vs off %.2e, vs generator %.2e" % (shape, sols[2][1], sols[0][1], d, e), flush=True)
    if d > 1e-8 or sols[2][1] != sols[0][1]:
        bad += 1
print("FAILED" if bad else "all solves reproduced bit for bit")
sys.exit(1 if bad else 0)
