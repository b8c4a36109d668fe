"""One pencil on one SM (GPU box tool): per-step time of the round-2 sweep kernel on an nx x ny x nz grid swept as ONE part.
usage: python tools/s2_pencil.py 120x8x8 [key=value ...]"""
import sys
sys.path.insert(0, '/root/repo')
from opm_autodiff_b200 import bridge, synth
for shape in sys.argv[1].split(","):
    nx, ny, nz = [int(t) for t in shape.split("x")]
    s = synth.small(nx, ny, nz)
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    be.set_option("sweep_parts", 1)
    for kv in sys.argv[2:]:
        k, v = kv.split("=")
        be.set_option(k, float(v))
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
    be.ilu0_factorize()
    lo = 1e3 * be.time_kernel("ilu_lower", 20, False)[0]
    up = 1e3 * be.time_kernel("ilu_upper", 20, False)[0]
    nlev = nx + ny + nz - 2
    print("%-12s %s: lower %7.1f us = %6.0f cycles per level, upper %7.1f us = %6.0f cycles per level (%d levels, %d rows per level)" %
          (shape, " ".join(sys.argv[2:]), lo, lo * 1965 / nlev, up, up * 1965 / nlev, nlev, ny * nz), flush=True)
