#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.

  python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_<workload>.txt

The per-launch times of such a pass are cold-cache and serialised (B200_PROFILING.md): what is
compared with bench.py's CUDA-event numbers is each kernel's SHARE of the step, not the absolute."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("b200::", "")
        ns = float(r["Metric Value"]) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r["Metric Unit"], 1.0)
        rows.append((name, ns, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, {"n": 0, "ns": 0.0, "min": 1e30, "max": 0.0, "grid": grid, "block": block})
        a["n"] += 1; a["ns"] += ns; a["min"] = min(a["min"], ns); a["max"] = max(a["max"], ns)
    total = sum(a["ns"] for a in agg.values())
    print("# %d launches, %.3f ms of kernel time (serialised, cold cache, under ncu)" % (len(rows), total * 1e-6))
    print("%-28s %8s %12s %10s %10s %10s %7s  %s" % ("kernel", "launches", "total_us", "mean_us", "min_us", "max_us", "share", "grid x block (first)"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
        print("%-28s %8d %12.1f %10.2f %10.2f %10.2f %6.1f%%  %s x %s" % (name, a["n"], a["ns"] * 1e-3, a["ns"] / a["n"] * 1e-3,
                                                                        a["min"] * 1e-3, a["max"] * 1e-3, 100.0 * a["ns"] / total,
                                                                        a["grid"], a["block"]))


if __name__ == "__main__":
    main(sys.argv[1])
