#!/usr/bin/env python
"""A/B of solver options on one configuration (GPU box tool): python tools/ab_options.py c2 spmv_sell sweep_early fuse_spmv defer_x"""
import sys, os, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
s = synth.small(*[int(t) for t in sys.argv[1].split("x")]) if "x" in sys.argv[1] else synth.full_system(sys.argv[1])
w = s.wells
wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv) if w is not None else None
keys = sys.argv[2:]
for rep in range(2):
    for combo in itertools.product((1, 0), repeat=len(keys)):
        be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
        for k, v in zip(keys, combo):
            be.set_option(k, v)
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
        res = bridge.BdaResult()
        be.solve_resident(res); be.solve_resident(res)
        be.timer_start()
        for _ in range(5):
            be.solve_resident(res)
        t = be.timer_stop() / 5
        if rep:
            print(" ".join("%s=%d" % kv for kv in zip(keys, combo)), ": %.3f ms per solve, %.1f iterations" % (t, res.it), flush=True)
        del be
