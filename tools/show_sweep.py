#!/usr/bin/env python
"""Print a tools/sweep_wk.py result (or the shared_sweep table of a bench.py line) as a w x K table (fraction of the measured HBM peak / G codes/s);
with two files, the second is shown next to the first."""
import json
import sys

def load(path):
    """tools/sweep_wk.py output, or a bench.py line (its `shared_sweep` table)"""
    d = json.loads(open(path).read())
    if "shared_sweep" in d:
        cols = d["shared_sweep"]["columns"]
        return {"rows": [dict(zip(cols, r)) for r in d["shared_sweep"]["rows"]]}
    return d


files = [load(f) for f in sys.argv[1:]]
ops = sorted({r["op"] for d in files for r in d["rows"]})
for op in ops:
    ks = sorted({r["K"] for d in files for r in d["rows"] if r["op"] == op})
    print(op)
    print("w   " + "  ".join(f"K={k:<{14 * len(files) - 2}d}" for k in ks))
    for w in range(1, 33):
        cells = []
        for k in ks:
            c = []
            for d in files:
                r = [x for x in d["rows"] if x["op"] == op and x["w"] == w and x["K"] == k]
                c.append(f"{r[0]['frac_of_measured_peak']:.2f}/{r[0]['gcodes_per_s']:7.1f}" if r else " " * 12)
            cells.append(" | ".join(c))
        print(f"{w:2d}  " + "   ".join(cells))
