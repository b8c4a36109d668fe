#!/usr/bin/env python
"""Stage timeline of one triangular sweep (GPU box tool): where do the consumer warps of a part wait?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
shape = tuple(int(t) for t in sys.argv[1].split("x"))
s = synth.small(*shape)
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
for kv in sys.argv[2:]:
    k, v = kv.split("=")
    be.set_option(k, float(v))
be.set_option("sweep_trace", 1)
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
be.ilu0_factorize()
lo, _ = be.time_kernel("ilu_lower", 2, False)
tr2 = be.sweep_trace()
tr, det, hlp = tr2[0], tr2[1], tr2[2]
print("lower %.1f us" % (lo * 1e3))
ghz = 1.965
for part in sorted(set([0, 1, 5, 20, 74, 147])):
    t = tr[part]
    n = int(np.count_nonzero(t[:, 2]))
    if n == 0:
        continue
    t = t[:n].astype(np.float64)
    t0 = t[0, 0]
    wait = (t[:, 1] - t[:, 0]) / ghz / 1e3
    work = (t[:, 2] - t[:, 1]) / ghz / 1e3
    issue_to_land = (t[:, 1] - t[:, 3]) / ghz / 1e3
    print("part %3d: %4d stages, total %7.1f us | waiting for data %7.1f us (mean %.2f, max %.2f) | working %7.1f us (mean %.2f) | issue->consumed mean %.2f us"
          % (part, n, (t[-1, 2] - t0) / ghz / 1e3, wait.sum(), wait.mean(), wait.max(), work.sum(), work.mean(), issue_to_land.mean()))
    d = det[part][:n].astype(np.float64)
    nrec = max(d[:, 0].sum(), 1)
    print("          warp 0: %d records; cycles per record: fetch+pre-barrier %.0f, barriers+ext wait %.0f, dependent part %.0f"
          % (nrec, d[:, 1].sum() / nrec, d[:, 2].sum() / nrec, d[:, 3].sum() / nrec))
    if part in (0, 74):
        hh = hlp[part][:n].astype(np.float64)
        for i in range(min(n, 40)):
            print("   stage %3d: wait %6.2f work %6.2f us  (issued %7.2f, wait-begin %7.2f, landed %7.2f, done %7.2f)  helper: %3d ext rows, start %7.2f done %7.2f"
                  % (i, wait[i], work[i], (t[i, 3] - t0) / ghz / 1e3, (t[i, 0] - t0) / ghz / 1e3, (t[i, 1] - t0) / ghz / 1e3, (t[i, 2] - t0) / ghz / 1e3,
                     hh[i, 2], (hh[i, 0] - t0) / ghz / 1e3, (hh[i, 1] - t0) / ghz / 1e3))
