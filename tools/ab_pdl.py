import sys; sys.path.insert(0,'/root/repo')
from opm_autodiff_b200 import bridge, synth
for name in ("c3", "c2"):
    s = synth.full_system(name); w = s.wells
    wc = bridge.WellContributions.from_arrays(w.val_pointers, w.Bcols, w.Ccols, w.B, w.C, w.Dinv) if w is not None else None
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, wc)
    res = bridge.BdaResult()
    for rep in range(2):
        for pdl in (1, 0):
            be.set_option("iter_pdl", pdl)
            be.solve_resident(res); be.solve_resident(res)
            be.timer_start()
            for _ in range(5):
                be.solve_resident(res)
            print("%s iter_pdl %d: %.3f ms per solve, %.1f iterations" % (name, pdl, be.timer_stop() / 5, res.it), flush=True)
    del be
