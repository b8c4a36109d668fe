import sys; sys.path.insert(0,'/root/repo')
from opm_autodiff_b200 import bridge, synth
s = synth.full_system("c3")
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
for nw in (0, 1):
    be.set_option("sweep_nowait", nw)
    print("nowait %d: lower %.1f us, upper %.1f us" % (nw, 1e3*be.time_kernel("ilu_lower", 10, False)[0], 1e3*be.time_kernel("ilu_upper", 10, False)[0]), flush=True)
