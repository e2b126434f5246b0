#!/usr/bin/env python
"""ncu CSV (dram bytes per scan launch) -> profiles/r01_scan_traffic.json"""
import csv
import json
import re
import sys

src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
per = {}
for r in rows:
    name, metric, unit, val = r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
    m = re.search(r"scan_(?:range|direct)_kernel<\(?(?:int\))?(\d+)", name)
    if not m:
        continue
    w = int(m.group(1))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "ms": 1e6}.get(unit, 1)
    per.setdefault(w, {})[metric] = val * scale
out = {"how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, "
              "one launch per width at 2^30 codes (tools/scan_traffic.py); writes still resident in the 126 MB L2 "
              "at kernel end are not counted by dram__bytes_write",
       "per_width": [{"w": w, "dram_read": per[w].get("dram__bytes_read.sum"), "dram_write": per[w].get("dram__bytes_write.sum"),
                      "ns": per[w].get("gpu__time_duration.sum"), "algorithmic": (1 << 30) * (w / 8 + 1 / 8)} for w in sorted(per)]}
out["traffic_per_step"] = sum((p["dram_read"] or 0) + (p["dram_write"] or 0) for p in out["per_width"])
out["algorithmic_per_step"] = sum(p["algorithmic"] for p in out["per_width"])
json.dump(out, open(dst, "w"), indent=1)
print(out["traffic_per_step"], out["algorithmic_per_step"], len(out["per_width"]))
