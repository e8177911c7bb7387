#!/usr/bin/env python
"""Per-SASS-instruction warp-stall samples of the first kernel in an ncu report.
    python scripts/sass_hot.py report.ncu-rep [top N] [lo hi]   (lo..hi = print that index range with counts)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
b = starts[0]
header = rows[b + 1]
end = starts[1] if len(starts) > 1 else len(rows)
body = [r for r in rows[b + 2:end] if len(r) == len(header) and r != header]
si = header.index("Warp Stall Sampling (All Samples)")
srci = header.index("Source")
ie = header.index("Instructions Executed")
tot = sum(int(r[si]) for r in body)
print(rows[b][1], "samples", tot, "instructions", len(body))
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for i in range(lo, hi):
        print(f"{int(body[i][si]):7d} {body[i][ie]:>10s} #{i:4d} {body[i][srci][:110]}")
else:
    for lo in range(0, len(body), 50):
        s = sum(int(r[si]) for r in body[lo:lo + 50])
        if s > tot * 0.005:
            print(f"range {lo:4d}: {100 * s / tot:5.1f}%  exec {body[lo][ie]}")
    for i in sorted(range(len(body)), key=lambda i: -int(body[i][si]))[:top]:
        print(f"{100 * int(body[i][si]) / tot:5.1f}% {body[i][ie]:>10s} #{i:4d} {body[i][srci][:100]}")
