#!/usr/bin/env python
"""Sweeps the launch knobs of the triangular-solve kernels on one workload (GPU box tool)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opm_autodiff_b200 import bridge, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
cfg = synth.CONFIGS[name]
s = synth.full_system(cfg)
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
be.ilu0_factorize()
out = []
for bps in (1, 2, 3, 4, 8):
    for sl in (0, 32, 100, 300):
        be.set_option("trsv_blocks_per_sm", bps)
        be.set_option("trsv_sleep_ns", sl)
        lo, byl = be.time_kernel("ilu_lower", 10, True)
        up, byu = be.time_kernel("ilu_upper", 10, True)
        out.append({"blocks_per_sm": bps, "sleep_ns": sl, "lower_us": round(lo * 1e3, 1), "upper_us": round(up * 1e3, 1),
                    "apply_gbs": round((byl + byu) / (lo + up) * 1e-6, 1)})
        print(out[-1], flush=True)
for k in ("spmv", "vec_p", "vec_xr1", "vec_xr2", "permute", "ilu_factor"):
    ms, by = be.time_kernel(k, 10, True)
    print(k, round(ms * 1e3, 1), "us", round(by / ms * 1e-6, 1), "GB/s", flush=True)
