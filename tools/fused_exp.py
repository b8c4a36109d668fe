#!/usr/bin/env python
"""A/B of the SpMV fused into the upper sweep (GPU box tool): python tools/fused_exp.py c3 [key=value ...]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
s = synth.small(*[int(t) for t in sys.argv[1].split("x")]) if "x" in sys.argv[1] else synth.full_system(sys.argv[1])
w = s.wells
wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv) if w is not None else None
xs = {}
for fuse, defer in ((1, 1), (0, 1), (1, 0), (0, 0)):
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    be.set_option("fuse_spmv", fuse); be.set_option("defer_x", defer)
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        be.set_option(k, float(v))
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
    res = bridge.BdaResult()
    be.solve_resident(res); be.solve_resident(res)
    be.timer_start()
    for _ in range(3):
        be.solve_resident(res)
    t = be.timer_stop() / 3
    x = np.zeros(3 * s.Nb); be.get_result(x); xs[fuse] = x if defer else xs.get(fuse, x)
    if fuse and os.environ.get("FUSE_DEBUG"):
        be.set_option("use_graph", 0); be.set_option("fuse_debug", int(os.environ["FUSE_DEBUG"])); be.solve_resident(res)
    print("fuse_spmv %d defer_x %d: %.2f ms per solve, %.1f iterations, converged %s, reduction %.2e" % (fuse, defer, t, res.it, res.converged, res.reduction), flush=True)
    del be
print("relative difference of the two solutions: %.2e" % (np.linalg.norm(xs[1] - xs[0]) / np.linalg.norm(xs[0])))
