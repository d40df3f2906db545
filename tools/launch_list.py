"""Compact launch list of ONE step from an `ncu --metrics gpu__time_duration.sum --csv` log:
keeps the launches between the last two optimizer kernels (one whole step) and writes
id,kernel,gpu__time_duration_us,grid,block plus a TOTAL row.
usage: python tools/launch_list.py <ncu.csv> <out.csv>"""
import csv
import sys

src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]
ki, vi, gi, bi, ii = (h.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Block Size", "ID"))
body = rows[1:]
opt = [i for i, r in enumerate(body) if "adam" in r[ki]]
a, b = (opt[-2] + 1, opt[-1] + 1) if len(opt) >= 2 else (0, len(body))
tot = 0.0
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel", "gpu__time_duration_us", "grid", "block"])
    for r in body[a:b]:
        us = float(r[vi].replace(",", "")) / 1e3
        tot += us
        w.writerow([r[ii], r[ki][:90], "{:.3f}".format(us), r[gi], r[bi]])
    w.writerow(["", "TOTAL one step ({} launches, serialised under ncu)".format(b - a), "{:.3f}".format(tot), "", ""])
print("{} launches, {:.1f} us".format(b - a, tot))
