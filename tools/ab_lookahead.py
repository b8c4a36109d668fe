#!/usr/bin/env python
"""A/B of the number of BiCGSTAB iterations enqueued per convergence read-back (option "lookahead") on one configuration
(GPU box tool): python tools/ab_lookahead.py c3 2 3 4 6"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth

s = synth.full_system(sys.argv[1])
w = s.wells
wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv) if w is not None else None
vals = [int(t) for t in sys.argv[2:]] or [2, 3, 4]
for rep in range(2):
    for la in vals:
        be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
        be.set_option("lookahead", la)
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
        res = bridge.BdaResult()
        be.solve_resident(res); be.solve_resident(res)
        be.timer_start()
        for _ in range(8):
            be.solve_resident(res)
        t = be.timer_stop() / 8
        if rep:
            print("lookahead=%d : %.3f ms per solve, %.1f iterations" % (la, t, res.it), flush=True)
        del be
