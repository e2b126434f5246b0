#!/usr/bin/env python
"""Developer micro-benchmark: the fused scan at chosen widths, timed like bench.py's per_width table
(10 back-to-back launches per event pair, best of 5 pairs).  Prints `w:frac` per width."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

PEAK = 6550.7
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass

ap = argparse.ArgumentParser()
ap.add_argument("--widths", default="1,2,3,6")
ap.add_argument("--logn", type=int, default=30)
ap.add_argument("--sel", default="quarter", choices=["quarter", "eq"])
a = ap.parse_args()
n = 1 << a.logn
out = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
cnt = torch.empty(1, dtype=torch.int64, device="cuda")
res = []
for w in [int(x) for x in a.widths.split(",")]:
    m = (1 << w) - 1
    lo, hi = (m // 4, m // 2) if a.sel == "quarter" else (3 & m, 3 & m)
    packed = api.generate_packed(n, w, seed=42 + w)
    fn = lambda: api.scan(packed, n, w, lo, hi, out=out, count=cnt)  # noqa: E731
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    nb = n * (w / 8 + 1 / 8)
    res.append(f"{w}:{nb / best / 1e6 / PEAK:.3f}({best * 1e3:.1f}us)")
    del packed
print(" ".join(res))
