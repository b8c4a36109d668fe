#!/usr/bin/env python
"""ILU0 factorisation alone (GPU box tool): python tools/profile_factor.py c3"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
s = synth.full_system(sys.argv[1])
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
for graph in (1, 0):
    for pdl in (1, 0):
        be.set_option("use_graph", graph); be.set_option("fac_pdl", pdl)
        ms, by = be.time_kernel("ilu_factor", 5, False)
        print("graph %d pdl %d: %.3f ms per factorisation (+ stream fills)" % (graph, pdl, ms), flush=True)
