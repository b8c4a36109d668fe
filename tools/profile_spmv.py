#!/usr/bin/env python
"""SpMV alone on one configuration (GPU box tool): both layouts, several grid caps; also the ncu driver for k_spmv*.
   python tools/profile_spmv.py c3 [sell=1] [blocks=2368] ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth
s = synth.full_system(sys.argv[1])
opts = dict(kv.split("=") for kv in sys.argv[2:])
sells = [int(opts["sell"])] if "sell" in opts else [1, 0]
blocks = [int(opts["blocks"])] if "blocks" in opts else [4096, 2368, 1184, 592]
for sell in sells:
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    be.set_option("spmv_sell", sell)
    be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
    for b in blocks:
        be.set_option("spmv_blocks", b)
        warm, by = be.time_kernel("spmv", 10, False)
        cold, _ = be.time_kernel("spmv", 10, True)
        print("sell %d blocks cap %5d: %.1f us (L2 as left by the previous launch) %.1f us (L2 flushed): %.0f / %.0f GB/s" %
              (sell, b, 1e3 * warm, 1e3 * cold, by / warm * 1e-6, by / cold * 1e-6), flush=True)
    del be
