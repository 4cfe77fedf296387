#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu --page source --csv` (needs -lineinfo, --import-source on).
usage: python scripts/ncu_hot.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
# first line is the kernel name row
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rows)
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
print("total samples", tot, "instructions", len(rows))
idx = sorted(range(len(rows)), key=lambda i: -int(rows[i]["# Samples"] or 0))[:N]
for i in sorted(idx):
    r = rows[i]
    n = int(r["# Samples"] or 0)
    top = sorted(((int(r[c] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%% %-9s %s   | %s" % (i, 100.0 * n / tot, r["Address"][-6:], r["Source"][:70], " ".join("%s=%d" % (c[6:], v) for v, c in top if v)))
