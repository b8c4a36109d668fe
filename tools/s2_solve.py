"""Whole-solve timing for option sets (GPU box tool).  usage: python tools/s2_solve.py c3 "k=v,k=v" ..."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from opm_autodiff_b200 import bridge, synth
wl = sys.argv[1]
s = synth.full_system(wl)
from tests.helpers import bridge_wells
for spec in sys.argv[2:] or [""]:
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    late = {}
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        if k == "sweep_nowait": late[k] = float(v)
        else: be.set_option(k, float(v))
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells))
    for k, v in late.items(): be.set_option(k, v)
    res = bridge.BdaResult()
    for _ in range(3): be.solve_resident(res)
    t0 = time.time()
    n = 8
    for _ in range(n): be.solve_resident(res)
    dt = (time.time() - t0) / n
    lo = 1e3 * be.time_kernel("ilu_lower", 10, False)[0]
    up = 1e3 * be.time_kernel("ilu_upper", 10, False)[0]
    print("%-50s %7.2f ms per solve (%5.1f solves/s), it %.1f conv %d, lower %6.1f upper %6.1f us" % (spec or "(defaults)", dt * 1e3, 1 / dt, res.it, res.converged, lo, up), flush=True)
    del be
