#!/usr/bin/env python
"""Hot SASS lines of the first kernel in an ncu report (needs --import-source on / --set full):
    python tools/ncu_hot.py REPORT.ncu-rep [N]
prints the N lines with most stall samples, with executed counts and the two top stall reasons."""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
blk = []
for r in rows[2:]:
    if len(r) < len(hdr) - 2 or r[0] == "Address":
        break
    blk.append(r)
tot = sum(int(r[i_s] or 0) for r in blk)
print("kernel", rows[0][1][:100], "lines", len(blk), "samples", tot)
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in blk:
    for c in stall_cols:
        agg[hdr[c][6:]] = agg.get(hdr[c][6:], 0) + int(r[c] or 0)
print("stalls", sorted(agg.items(), key=lambda kv: -kv[1])[:8])
top = sorted(range(len(blk)), key=lambda i: -int(blk[i][i_s] or 0))[:N]
for i in sorted(top):
    r = blk[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    print(i, r[i_s], r[i_ex], r[i_src].strip()[:90], st)
