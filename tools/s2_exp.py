"""Round-2 sweep experiments on one GPU: isolated lower / upper sweep time for option sets (before the first solve).
usage: python tools/s2_exp.py c3 "s2_groups=4,s2_wg=2" "sweep_nowait=1" ...   (each argument = one solver with those options)"""
import sys, time
sys.path.insert(0, '/root/repo')
from opm_autodiff_b200 import bridge, synth
wl = sys.argv[1]
s = synth.full_system(wl)
for spec in sys.argv[2:] or [""]:
    be = bridge.B200SolverBackend(1 if "verbose" in spec else 0, 2000, 1e-10, 0)
    late = {}
    for kv in filter(None, spec.split(",")):
        if kv == "verbose": continue
        k, v = kv.split("=")
        if k in ("sweep_nowait",): late[k] = float(v)
        else: be.set_option(k, float(v))
    t0 = time.time()
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
    for k, v in late.items(): be.set_option(k, v)
    lo = 1e3 * be.time_kernel("ilu_lower", 10, False)[0]
    up = 1e3 * be.time_kernel("ilu_upper", 10, False)[0]
    print("%-60s lower %7.1f us  upper %7.1f us   (setup %.1f s)" % (spec or "(defaults)", lo, up, time.time() - t0), flush=True)
    del be
