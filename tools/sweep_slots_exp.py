#!/usr/bin/env python
"""Sweep time against ring depth / stage size / helper count (GPU box tool)."""
import sys; sys.path.insert(0,'/root/repo')
from opm_autodiff_b200 import bridge, synth
s = synth.full_system("c3")
for slots, stage, helpers in ((2, 81920, 2), (3, 54000, 3), (4, 40960, 4), (4, 40960, 2), (6, 27000, 6), (6, 27000, 3), (8, 20000, 4)):
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    try:
        be.set_option("sweep_slots", slots); be.set_option("sweep_stage_bytes", stage); be.set_option("sweep_helpers", helpers)
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
        r = []
        for nw in (0, 1):
            be.set_option("sweep_nowait", nw)
            r.append((1e3*be.time_kernel("ilu_lower", 10, False)[0], 1e3*be.time_kernel("ilu_upper", 10, False)[0]))
        print("slots %d stage %6d helpers %d: lower %.1f upper %.1f us | unstarved: lower %.1f upper %.1f" % (slots, stage, helpers, r[0][0], r[0][1], r[1][0], r[1][1]), flush=True)
    except Exception as e:
        print("slots %d stage %d helpers %d: %s" % (slots, stage, helpers, e), flush=True)
    del be
