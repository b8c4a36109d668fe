#!/usr/bin/env python
"""Key metrics of every kernel in an .ncu-rep (raw page) -> text summary for profiles/."""
import csv, io, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.max.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max"]
for n, r in enumerate(rows[2:]):
    print("## launch %d" % n)
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("%-72s %s %s" % (w, r[i], units[i]))
    try:
        rd, wr = float(r[hdr.index("dram__bytes_read.sum")]), float(r[hdr.index("dram__bytes_write.sum")])
        ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
        sc = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        t = float(r[hdr.index("gpu__time_duration.sum")]) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[units[hdr.index("gpu__time_duration.sum")]]
        tot = rd * sc[ur] + wr * sc[uw]
        print("%-72s %.1f MB -> %.1f GB/s over the kernel duration" % ("DRAM traffic (read + write)", tot * 1e-6, tot / t * 1e-9))
    except Exception as e:
        print("traffic n/a", e)
