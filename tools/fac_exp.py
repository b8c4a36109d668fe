"""Factorisation timing: one dataflow launch (fac_rows3=1) against one launch per level (GPU box tool).
usage: python tools/fac_exp.py c3 "fac_rows3=1,fac_flow_cap=2" ..."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from opm_autodiff_b200 import bridge, synth
from tests.helpers import bridge_wells
wl = sys.argv[1]
s = synth.full_system(wl)
x0 = None
for spec in sys.argv[2:] or ["fac_rows3=0", "fac_rows3=1"]:
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        be.set_option(k, float(v))
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, bridge_wells(s.wells))
    res = bridge.BdaResult()
    for _ in range(2): be.solve_resident(res)
    x = np.zeros(3 * s.Nb); be.get_result(x)
    if x0 is None: x0 = x
    fac = 1e3 * be.time_kernel("ilu_factor", 10, False)[0]
    print("%s %-40s factorisation %8.1f us, it %.1f conv %d, x identical %s" % (wl, spec, fac, res.it, res.converged, bool(np.array_equal(x, x0))), flush=True)
    del be
