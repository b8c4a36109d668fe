"""Sweep time against the number of parts for given grid shapes (GPU box tool).
usage: python tools/s2_parts_exp.py 100x100x13 100x100x25 ..."""
import sys, time
sys.path.insert(0, '/root/repo')
from opm_autodiff_b200 import bridge, synth
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(48, 48, 48), (64, 64, 64), (30, 30, 30)]
for shape in shapes:
    s = synth.small(*shape)
    import os
    for parts in [int(v) for v in os.environ.get('PARTS', '0,16,24,37,48,74,100,148').split(',')]:
        be = bridge.B200SolverBackend(int(os.environ.get("VERB", "0")), 2000, 1e-10, 0)
        be.set_option("sweep_parts", parts)
        for kv in filter(None, os.environ.get("OPTS", "").split(",")):
            k, v = kv.split("=")
            be.set_option(k, float(v))
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
        res = bridge.BdaResult()
        for _ in range(2): be.solve_resident(res)
        t0 = time.time()
        for _ in range(5): be.solve_resident(res)
        dt = (time.time() - t0) / 5
        lo = 1e3 * be.time_kernel("ilu_lower", 10, False)[0]
        up = 1e3 * be.time_kernel("ilu_upper", 10, False)[0]
        print("%s Nb %d parts %3d: %.2f ms per solve, it %.1f, lower %.1f upper %.1f us" % (shape, s.Nb, parts, dt*1e3, res.it, lo, up), flush=True)
        del be
