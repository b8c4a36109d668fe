#!/usr/bin/env python
"""Sweep ONE solver option over several values on one configuration (GPU box tool):
  python tools/ab_values.py c2 sweep_parts 0 32 48 64 96      (or a grid: 64,64,48[,nwells])
Prints device ms per resident solve (mean of 8 after 2 warm-up solves, second of two passes) and the iteration count."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth

name = sys.argv[1]
if "," in name:                                   # nx,ny,nz[,nwells]
    t = [int(v) for v in name.split(",")]
    s = synth.full_system(synth.GridConfig("custom-%dx%dx%d" % tuple(t[:3]), t[0], t[1], t[2], nwells=t[3] if len(t) > 3 else 0, nperf=8))
else:
    s = synth.full_system(name)
key = sys.argv[2]
values = [float(t) for t in sys.argv[3:]]
w = s.wells
wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv) if w is not None else None
for rep in range(2):
    for v in values:
        be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
        be.set_option(key, v)
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
        res = bridge.BdaResult()
        be.solve_resident(res); be.solve_resident(res)
        be.timer_start()
        for _ in range(8):
            be.solve_resident(res)
        t = be.timer_stop() / 8
        if rep:
            print("%s=%g : %.3f ms per solve, %.1f iterations, converged %d" % (key, v, t, res.it, res.converged), flush=True)
        del be
