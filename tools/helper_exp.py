#!/usr/bin/env python
"""Sweep time against the helper warps' poll interval and count (GPU box tool)."""
import sys; sys.path.insert(0,'/root/repo')
from opm_autodiff_b200 import bridge, synth
s = synth.full_system("c3")
for helpers in [int(a) for a in sys.argv[1:]] or [2, 1]:
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    be.set_option("sweep_helpers", helpers)
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
    for sleep in (0, 60, 250):
        be.set_option("sweep_helper_sleep", sleep)
        print("helpers %d sleep %4d ns: lower %.1f us, upper %.1f us" % (helpers, sleep, 1e3*be.time_kernel("ilu_lower", 10, False)[0], 1e3*be.time_kernel("ilu_upper", 10, False)[0]), flush=True)
    del be
