#!/usr/bin/env python
"""Isolates the latency terms of the triangular sweeps on special grids (GPU box tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth

def run(shape, label, **opts):
    s = synth.small(*shape)
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    for k, v in opts.items():
        be.set_option("sweep_" + k, v)
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
    be.ilu0_factorize()
    lo, _ = be.time_kernel("ilu_lower", 10, False)
    up, _ = be.time_kernel("ilu_upper", 10, False)
    nlev = sum(shape) - 2
    print("%-28s grid %-12s %-60s lower %8.1f us upper %8.1f us  (%d levels -> %.3f us/level lower)"
          % (label, "x".join(map(str, shape)), opts, lo * 1e3, up * 1e3, nlev, lo * 1e3 / nlev), flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "micro"):
    run((4000, 1, 1), "1-D chain, one part", parts=1)
    run((200, 8, 8), "one pencil 8x8", parts=1)
    run((200, 8, 8), "one pencil 8x8, 32K", parts=1, stage_bytes=32768, slots=4)
    run((200, 8, 8), "pencil 8x8 cut in 4 parts", parts=4)
    run((200, 32, 32), "4x4 pencils", parts=16)
if which in ("all", "c3"):
    run((100, 100, 100), "c3 default", parts=148)
    run((100, 100, 100), "c3 64K/3 slots", parts=148, stage_bytes=65536, slots=3)
    run((100, 100, 100), "c3 12 warps", parts=148, warps=12)
