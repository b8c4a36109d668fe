"""bench.py's N > 1 leg: one rank per GPU under torchrun, row slabs of the same synthetic system.

Every rank generates only its own slab (synth.generate is slab-independent), the set-up exchange (halo
requests, NCCL id, CUDA-IPC handles) goes through torch.distributed, the data path (halo over peer memory,
all-reduce of the dot products through peer-memory mailboxes) runs inside libb200bda.so.  Timing: barrier +
device synchronise on both sides, CUDA events on each rank's solver stream, MAX over ranks.

The main line is the workload asked for (default C3, the same system as the one-GPU line: strong scaling).
BASELINE.json's own multi-GPU configurations ride along under "also": C4 (10 M cells) at every N >= 2 and
C5 (50 M cells, heterogeneous) at N = 8, each with its own steps, clocks and per-kernel shares."""
from __future__ import annotations

import json
import os
import time

import numpy as np

KERNELS = ("permute", "ilu_factor", "ilu_stream", "ilu_lower", "ilu_upper", "ilu_upper_spmv", "spmv", "spmv_ghost", "halo_push", "allreduce", "finish",
           "well_apply", "vec_p", "vec_xr1", "vec_xr2", "init", "unpermute")


def measure(cfg, args, steps, warmup, tol, maxit, ClockSampler, measured_peak, td, torch, rank, world, local):
    """One workload on `world` row slabs: e2e (host buffers), device-resident value, per-kernel profile.  Returns the
    line's keys on rank 0 (None elsewhere)."""
    from . import bridge, dist

    def sync():
        torch.cuda.synchronize()
        td.barrier()

    def maxf(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t[0])

    def sumf(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        td.all_reduce(t, op=td.ReduceOp.SUM)
        return float(t[0])

    t0 = time.perf_counter()
    blocks = getattr(args, "partition", "slabs") == "blocks"
    bp = dist.block_partition(cfg.nx, cfg.ny, cfg.nz, world)
    ls = dist.block_system(cfg, rank, world) if blocks else dist.slab_system(cfg, rank, world)
    t_gen = time.perf_counter() - t0
    ds = dist.DistSolver(ls, local, maxit=maxit, tolerance=tol)
    for kv in filter(None, os.environ.get("B200_OPTIONS", "").split(",")):      # experiment knob: "option=value,..."
        k, v = kv.split("=")
        ds.be.set_option(k, float(v))
    res = bridge.BdaResult()

    # ---- e2e: host buffers, H2D of values + rhs and D2H of x inside the timed region --------------------
    x = np.zeros(ds.N)
    ds.register_host_buffers(x)               # page-locked once, explicitly, as a Flow rank's glue code would
    t_analysis = 0.0
    for _ in range(max(warmup, 2)):
        ds.solve_system(res)
        ds.get_result(x)
        t_analysis = max(t_analysis, res.t_analysis)      # the first call analyses the pattern (once per simulation)
    assert res.converged, "solve did not converge"
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        ds.solve_system(res)
        ds.get_result(x)
    sync()
    e2e_s = maxf((time.perf_counter() - t0) / steps)
    err2 = sumf(float(np.sum((x - ls.x_true) ** 2)))
    ref2 = sumf(float(np.sum(ls.x_true ** 2)))
    w = ls.wells
    h2d = sumf(ls.vals.nbytes + ls.b.nbytes + (0 if w is None else w.B.nbytes + w.C.nbytes + w.Dinv.nbytes + 8 * len(w.Bcols)))
    d2h = sumf(x.nbytes)

    # ---- value: system resident in HBM, device time, max over ranks --------------------------------------
    ds.upload()
    for _ in range(warmup):
        ds.solve_resident(res)
    ds.be.reset_stats()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    sync()
    ds.be.timer_start()
    for _ in range(steps):
        ds.solve_resident(res)
    dev_ms = ds.be.timer_stop()
    sync()
    clk = clocks.stop() if rank == 0 else None
    ms_per_step = maxf(dev_ms / steps)
    launches = sumf(ds.be.launch_count())
    assert res.converged

    # ---- per-kernel profile of one solve (CUDA events around every launch), rank-local, max over ranks ----
    ds.be.set_option("profile", 1)
    ds.be.reset_stats()
    sync()
    ds.solve_resident(res)
    ds.be.set_option("profile", 0)
    peak, peak_src = measured_peak()
    kernels, total_ms = {}, 0.0
    for k in KERNELS:
        n, ms, by = ds.be.kernel_stats(k)
        n, ms = int(maxf(n)), maxf(ms)
        if n:
            by = sumf(by)
            kernels[k] = {"launches": n, "ms_total": round(ms, 4), "us_per_launch": round(1e3 * ms / n, 3),
                          "alg_bytes_per_launch_all_ranks": by,
                          "gbs_all_ranks": round(by / (ms / n) * 1e-6, 1) if ms > 0 and by > 0 else None}
            total_ms += ms
    for k in kernels:
        kernels[k]["share"] = round(kernels[k]["ms_total"] / total_ms, 4)
    it = res.it
    big = ls.vals.nbytes > 1.3e8
    del ds
    sync()
    if rank != 0:
        return None
    # the upper sweep launch also runs the owned x owned SpMV (its CTAs take SpMV units as their parts finish): the unit is
    # "ILU apply + operator apply" with the algorithmic bytes of all three operations (as bench.py on one GPU)
    fused = "ilu_upper_spmv" in kernels
    up_key = "ilu_upper_spmv" if fused else "ilu_upper"
    ilu_ms = kernels["ilu_lower"]["ms_total"] + kernels[up_key]["ms_total"]
    ilu_n = kernels["ilu_lower"]["launches"]
    ilu_by = kernels["ilu_lower"]["alg_bytes_per_launch_all_ranks"] + kernels[up_key]["alg_bytes_per_launch_all_ranks"]
    cand = {"ilu_apply_spmv" if fused else "ilu_apply": (ilu_ms, ilu_n, ilu_by)}
    if not fused:
        cand["spmv"] = (kernels["spmv"]["ms_total"], kernels["spmv"]["launches"], kernels["spmv"]["alg_bytes_per_launch_all_ranks"])
    dom = max(cand, key=lambda k: cand[k][0])
    dms, dn, dby = cand[dom]
    achieved = dby / world / (dms / dn) * 1e-6          # per GPU, against one GPU's peak
    return {
        "value": 1e3 / ms_per_step, "unit": "solves/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
        "config": {"workload": cfg.name, "cells": cfg.ncells, "wells": cfg.nwells, "tolerance": tol, "relaxation": 1.0,
                   "partition": ("%d x %d blocks in (y, z), x not cut, rank-major numbering" % (bp.py, bp.pz) if blocks
                                 else "%d row slabs along k" % world) + ", ghosts last, block-Jacobi ILU0 per GPU",
                   "iterations": it, "x_error_vs_generator": float(np.sqrt(err2 / ref2)),
                   "l2": "per-GPU slab (matrix + factor) larger than L2, no flush" if big else "per-GPU slab fits L2 at this rank count",
                   "analysis_s_excluded": t_analysis, "generate_s": t_gen},
        "clocks": clk,
        "e2e": {"value": 1.0 / e2e_s, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": None, "peak_source": peak_src,
                     "note": "per GPU: algorithmic bytes of all ranks / ranks / slowest rank's mean launch time",
                     "share_of_step": round(dms / total_ms, 4)},
        "kernels": kernels,
    }


def run(args, metric, tol, maxit, get_cfg, ClockSampler, measured_peak, cpu_sample=None, load_system=None):
    import torch
    import torch.distributed as td
    from . import bridge
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if not bridge.device_available():
        raise SystemExit("bench.py needs B200s (sm_100); the backend has no CPU fallback")
    torch.cuda.set_device(local)
    if not td.is_initialized():
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = get_cfg(args.workload)
    main = measure(cfg, args, args.steps, args.warmup, tol, maxit, ClockSampler, measured_peak, td, torch, rank, world, local)

    # BASELINE.json's own multi-GPU configurations beside the main line (only next to the default workload)
    also = {}
    if cfg.name.startswith("c3") and not getattr(args, "no_also", False):
        extra = [("c4", max(5, min(args.steps, 6)))] + ([("c5", 5)] if world >= 8 else [])
        for name, steps in extra:
            try:
                r = measure(get_cfg(name), args, steps, 3, tol, maxit, ClockSampler, measured_peak, td, torch, rank, world, local)
                if rank == 0:
                    also[name] = r
            except Exception as e:                          # never lose the main line over a side measurement
                if rank == 0:
                    also[name] = {"error": str(e)}
                break                                        # ranks may be out of step after a failure: stop the side runs

    if rank == 0:
        # CPU baseline: the oracle port on this box's host cores, ONE thread = one MPI rank of the reference, bounded sample of
        # the whole (unpartitioned) system, as on the one-GPU line
        cpu = None
        if cpu_sample is not None and load_system is not None and not getattr(args, "no_cpu_baseline", False) and cfg.ncells <= 2_000_000:
            system = load_system(cfg)
            r, per_it, per_solve = cpu_sample(system, args.cpu_sample_iters, 1, 1, main["config"]["iterations"])
            cpu = {"value": 1.0 / per_solve, "unit": "solves/s", "cores": 1, "kind": "port",
                   "sample": "oracle (CPU port of the reference ISTL path, 1 thread = 1 MPI rank, the whole system): ILU0 "
                             "factorisation %.2f s + %d BiCGSTAB iterations at %.3f s/iteration, scaled to the %.1f iterations "
                             "of the converged %d-GPU solve" % (r.t_decomp, args.cpu_sample_iters, per_it, main["config"]["iterations"], world),
                   "host_cpus": os.cpu_count()}
        out = {"metric": metric, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
        out.update(main)
        out["cpu_baseline"] = cpu
        out["also"] = also or None
        print(json.dumps(out), flush=True)
    td.barrier()
    td.destroy_process_group()
