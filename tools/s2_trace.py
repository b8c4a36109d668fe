#!/usr/bin/env python
"""Round-2 sweep wavefront (GPU box tool): per part, when its first step ran, when its last, how long it waited for external
rows; for a few parts the step timeline.  usage: python tools/s2_trace.py c3 [lower|upper] [key=value ...]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
wl = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] in ("lower", "upper") else "lower"
s = synth.full_system(wl) if wl in ("c2", "c3") else synth.small(*[int(t) for t in wl.split("x")])
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
for kv in sys.argv[2:]:
    if "=" in kv and not kv.startswith("late:"):
        k, v = kv.split("=")
        be.set_option(k, float(v))
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
be.ilu0_factorize()
for kv in sys.argv[2:]:
    if kv.startswith("late:"):
        k, v = kv[5:].split("=")
        be.set_option(k, float(v))
t_plain, _ = be.time_kernel("ilu_" + which, 5, False)
be.set_option("sweep_trace", 1)
t_tr, _ = be.time_kernel("ilu_" + which, 1, False)
out = np.zeros(3 * 148 * 1024 * 4, np.int64)
be._chk(bridge.lib().b200_get_sweep_trace(be._h, out, out.size))
summ = out[:8192].reshape(1024, 8)
np_ = int(np.count_nonzero(summ[:, 5]))
det = out[8192:8192 + 2 * 256 * np_].reshape(np_, 256, 2)
t0 = summ[:np_, 0].min()
nst = np.minimum(256, summ[:np_, 5]).astype(int)
first = np.array([det[p, 0, 0] for p in range(np_)]); last = np.array([det[p, nst[p] - 1, 0] for p in range(np_)])
print("%s sweep: %.1f us plain, %.1f us traced, %d parts" % (which, t_plain * 1e3, t_tr * 1e3, np_))
print("part: first step at / last step at [us], median step time [us] (steps)")
order = np.argsort(first)
for r in range(0, np_, 4):
    print("   ".join("%3d: %6.1f %6.1f %.3f (%3d)" % (p, (first[p] - t0) * 1e-3, (last[p] - t0) * 1e-3, np.median(np.diff(det[p, :nst[p], 0])) * 1e-3 if nst[p] > 1 else 0, summ[p, 5])
                     for p in order[r:r + 4]))
for p in (int(order[0]), int(order[np_ // 2]), int(np.argmax(last))):
    n = nst[p]
    ts = (det[p, :n, 0] - t0) * 1e-3
    print("part %d: step start times [us]: %s" % (p, " ".join("%.1f" % v for v in ts[:n:max(1, n // 40)])))
prof = out[8192 + 2 * 256 * 1024:8192 + 2 * 256 * 1024 + 512].reshape(2, 32, 8)
for pi, name in ((0, "part 0"), (1, "part %d" % (np_ // 2))):
    print("%s, cycles per record: [fetch issue + start value | wait for previous step | external rows | fma | stores + arrive | shared loads of the dependencies], records" % name)
    for w in range(16):
        n = prof[pi, w, 6]
        if n > 0: print("   warp %2d: %s   %d" % (w, " ".join("%7.1f" % (prof[pi, w, k] / n) for k in range(6)), n))
