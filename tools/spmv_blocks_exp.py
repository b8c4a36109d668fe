import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np
from opm_autodiff_b200 import bridge, synth
s = synth.full_system("c3")
w = s.wells
wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv)
for blocks in (4096, 2368, 1184, 592):
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    be.set_option("spmv_blocks", blocks)
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
    res = bridge.BdaResult()
    be.solve_resident(res); be.solve_resident(res)
    be.set_option("profile", 1); be.reset_stats(); be.solve_resident(res); be.set_option("profile", 0)
    n, ms, by = be.kernel_stats("spmv")
    be.timer_start(); be.solve_resident(res); be.solve_resident(res); t = be.timer_stop() / 2
    print("spmv blocks cap %5d: %.1f us per launch (%d launches), solve %.2f ms, it %.1f" % (blocks, 1e3 * ms / n, n, t, res.it), flush=True)
    del be
