#!/usr/bin/env python
"""Dynamic (executed) instruction mix of the first kernel in an ncu report, from its SASS source page.
usage: ncu_dyn_opcodes.py report.ncu-rep"""
import collections
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
mix = collections.Counter()
seen_hdr = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        seen_hdr += 1
        if seen_hdr > 1:
            break
        print(r[1][:120])
        continue
    if len(r) < 7 or r[0] == "Address":
        continue
    try:
        ex = float(r[5])
    except ValueError:
        continue
    src = r[1].strip()
    toks = src.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    mix[op.split(".")[0]] += ex
tot = sum(mix.values())
print(f"warp-instructions executed: {tot:.0f}")
for op, c in mix.most_common(25):
    print(f"{op:10s} {c:14.0f} {100 * c / tot:6.2f}%")
