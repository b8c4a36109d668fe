#!/usr/bin/env python
"""Wavefront of one lower sweep over the parts (GPU box tool): when does every part finish its first stage and its last?"""
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
tr = be.sweep_trace()[0]
ghz = 1.965
print("lower %.1f us" % (lo * 1e3))
first, last, nst = [], [], []
for p in range(148):
    t = tr[p]
    n = int(np.count_nonzero(t[:, 2]))
    if n == 0:
        continue
    first.append((t[0, 2] - t[0, 0]) / ghz / 1e3); last.append((t[n - 1, 2] - t[0, 0]) / ghz / 1e3); nst.append(n)
print("first stage done [us] per part:")
for r in range(0, len(first), 12):
    print(" ".join("%6.1f" % v for v in first[r:r + 12]))
print("last stage done [us] per part:")
for r in range(0, len(last), 12):
    print(" ".join("%6.1f" % v for v in last[r:r + 12]))
