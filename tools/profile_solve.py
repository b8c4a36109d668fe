#!/usr/bin/env python
"""Tiny driver for ncu: a few BiCGSTAB iterations of a workload through the resident-solve path (the kernels of the solve with
their tails: lower sweep + deferred x update, upper sweep + SpMV).  usage: python tools/profile_solve.py c3 [maxit] [key=value ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
from tests.helpers import bridge_wells
wl = sys.argv[1]
maxit = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 3
s = synth.full_system(wl)
be = bridge.B200SolverBackend(0, maxit, 1e-30, 0)
for kv in sys.argv[2:]:
    if "=" in kv:
        k, v = kv.split("=")
        be.set_option(k, float(v))
be.set_option("use_graph", 0)
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells))
res = bridge.BdaResult()
for _ in range(2):
    be.solve_resident(res)
print("it %.1f reduction %.3e" % (res.it, res.reduction))
