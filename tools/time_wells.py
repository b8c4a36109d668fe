import sys; sys.path.insert(0,'/root/repo')
from opm_autodiff_b200 import bridge, synth
s = synth.full_system("c3"); w = s.wells
wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv)
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
print("well_apply %.2f us (warm) %.2f us (L2 flushed)" % (1e3*be.time_kernel("well_apply", 20, False)[0], 1e3*be.time_kernel("well_apply", 20, True)[0]))
