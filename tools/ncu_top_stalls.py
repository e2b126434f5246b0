#!/usr/bin/env python
"""Summarise an ncu report's SASS source page: top instructions by warp-stall samples.
usage: ncu_top_stalls.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernel = None
hdr = None
data = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break  # first kernel only
        kernel = r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) >= 7:
        try:
            data.append((float(r[2]), float(r[5]), r[1].strip()))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
print(kernel)
print(f"instructions in kernel: {len(data)}, total samples {tot:.0f}")
for i, (s, ex, src) in enumerate(data):
    pass
for s, ex, src in sorted(data, reverse=True)[:top]:
    print(f"{100 * s / tot:6.2f}%  exec={ex:>10.0f}  {src[:100]}")
