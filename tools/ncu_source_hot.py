#!/usr/bin/env python
"""Hot instructions of one kernel from an .ncu-rep (source page, SASS view): samples per instruction with the
dominant stall reasons.  usage: ncu_source_hot.py report.ncu-rep [kernel-index] [top-N] [--no-barrier]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 1
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, out, k = None, [], 0
for r in rows:
    if r and r[0] == "Kernel Name":
        k += 1
        if k == kidx: print("kernel:", r[1])
        continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and k == kidx and len(r) == len(hdr): out.append(r)
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in out)
agg = {}
for r in out:
    for i in stalls: agg[hdr[i]] = agg.get(hdr[i], 0) + int(r[i])
print("total samples", tot, "instructions", len(out))
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
top = sorted(range(len(out)), key=lambda i: -int(out[i][si]))[:topn]
for i in sorted(top):
    r = out[i]
    st = sorted([(int(r[j]), hdr[j][6:]) for j in stalls], reverse=True)[:2]
    print("%5d %-64s smp %6s exe %8s %s" % (i, r[1][:64], r[si], r[ii], st))
