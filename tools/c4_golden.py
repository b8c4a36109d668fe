"""Golden fingerprint of BASELINE.json's C4 (10 M cells) solved by the CPU oracle with 8 block-Jacobi partitions (= mpirun -np 8
semantics, PreconditionerFactory.hpp:237-252): iteration count and the solution at 4000 sampled block rows.  Run on a CPU box
(~12 GB, minutes); the 8-GPU parity test compares the GPU solve of the same slabs with it
(tests/test_gpu_dist.py::test_c4_eight_gpu_solve_matches_the_partitioned_oracle_fingerprint).

  python tools/c4_golden.py            -> tests/golden/c4_n8_oracle.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from opm_autodiff_b200 import dist, synth
from oracle import oracle
from tests.helpers import oracle_wells

world = 8
cfg = synth.CONFIGS["c4"]
t0 = time.time()
s = synth.full_system(cfg)
print("generated %d cells in %.1f s" % (s.Nb, time.time() - t0), flush=True)
ranges = dist.slab_ranges(cfg.nz, cfg.nx * cfg.ny, world)
part_ptr = np.array([a for a, _ in ranges] + [s.Nb], np.int32)
t0 = time.time()
ref = oracle.solve(s.rows, s.cols, s.vals, s.b, oracle_wells(s.wells), tol=1e-10, maxit=200, part_ptr=part_ptr,
                   threads=min(8, oracle.max_threads()))
print("oracle: converged %s, %.1f iterations, %.1f s" % (ref.converged, ref.it, time.time() - t0), flush=True)
rng = np.random.default_rng(2026)
idx = np.sort(rng.choice(s.Nb, size=4000, replace=False))
x = ref.x.reshape(-1, 3)
out = {"workload": cfg.name, "cells": int(s.Nb), "world": world, "tolerance": 1e-10, "iterations": float(ref.it),
       "converged": bool(ref.converged), "x_norm": float(np.linalg.norm(ref.x)),
       "x_error_vs_generator": float(np.linalg.norm(ref.x - s.x_true) / np.linalg.norm(s.x_true)),
       "sample_rows": idx.tolist(), "sample_x": x[idx].tolist()}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c4_n8_oracle.json")
json.dump(out, open(path, "w"))
print("wrote", path, os.path.getsize(path), "bytes")
