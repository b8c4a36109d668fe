#!/usr/bin/env python
"""Sweeps the schedule knobs of the triangular-sweep kernels on one workload (GPU box tool).

  python tools/sweep_trsv.py c3 "parts,warps,slots,stage_bytes[,window];..."
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_autodiff_b200 import bridge, synth

name = sys.argv[1] if len(sys.argv) > 1 else "c3"
combos = sys.argv[2] if len(sys.argv) > 2 else "148,8,6,16384;148,4,6,16384;148,8,8,8192;148,12,4,32768;74,8,6,16384"
cfg = synth.CONFIGS[name] if name in synth.CONFIGS else synth.GridConfig("custom", *[int(t) for t in name.split("x")])
s = synth.full_system(cfg)
for combo in combos.split(";"):
    t = [int(v) for v in combo.split(",")]
    parts, warps, slots, sb = t[:4]
    window = t[4] if len(t) > 4 else 0
    be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
    for k, v in (("sweep_parts", parts), ("sweep_warps", warps), ("sweep_slots", slots), ("sweep_stage_bytes", sb), ("sweep_window", window)):
        be.set_option(k, v)
    t0 = time.time()
    try:
        be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
        t_an = time.time() - t0
        be.ilu0_factorize()
        lo, byl = be.time_kernel("ilu_lower", 10, True)
        up, byu = be.time_kernel("ilu_upper", 10, True)
        print({"parts": parts, "warps": warps, "slots": slots, "stage_bytes": sb, "window": window, "lower_us": round(lo * 1e3, 1),
               "upper_us": round(up * 1e3, 1), "apply_gbs": round((byl + byu) / (lo + up) * 1e-6, 1), "upload+analysis_s": round(t_an, 2)}, flush=True)
    except RuntimeError as e:
        print({"combo": combo, "error": str(e)[:120]}, flush=True)
    del be
be = bridge.B200SolverBackend(0, 2000, 1e-10, 0)
be.upload_system(3 * s.Nb, 9 * s.nnzb, 3, s.vals, s.rows, s.cols, s.b, None)
for k in ("spmv", "vec_p", "vec_xr1", "vec_xr2", "permute", "ilu_factor"):
    ms, by = be.time_kernel(k, 10, True)
    print(k, round(ms * 1e3, 1), "us", round(by / ms * 1e-6, 1), "GB/s", flush=True)
