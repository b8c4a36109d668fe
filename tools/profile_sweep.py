#!/usr/bin/env python
"""Tiny driver for ncu: factorise one grid and launch each triangular sweep a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
shape = tuple(int(t) for t in sys.argv[1].split("x"))
s = synth.small(*shape)
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    be.set_option(k, float(v))
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
be.ilu0_factorize()
lo, _ = be.time_kernel("ilu_lower", 2, False)
up, _ = be.time_kernel("ilu_upper", 2, False)
print("lower %.1f us upper %.1f us" % (lo * 1e3, up * 1e3))
