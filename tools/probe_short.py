#!/usr/bin/env python
"""Developer probe: short columns (the ring-less scan kernel) -- 100 scans of 2^LOGN codes in one CUDA graph,
us per scan at several widths.   python tools/probe_short.py [logn ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
for logn in [int(x) for x in sys.argv[1:]] or [24]:
    n = 1 << logn
    out = []
    for w in [int(x) for x in os.environ.get("WIDTHS", "3,5,7,9,11,13,16,17,23,32").split(",")]:
        col = api.generate_packed(n, w, seed=42 + w)
        bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
        cnt = torch.empty(1, dtype=torch.int64, device="cuda")
        m = (1 << w) - 1
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                api.scan(col, n, w, m // 4, m // 2, out=bv, count=cnt)
            s.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(100):
                    api.scan(col, n, w, m // 4, m // 2, out=bv, count=cnt)
            g.replay()
            s.synchronize()
            best = 1e9
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(s)
                g.replay()
                b.record(s)
                b.synchronize()
                best = min(best, a.elapsed_time(b) / 100)
        ideal = n * (w / 8 + 1 / 8) / (peak * 1e9) * 1e3
        out.append(f"{w}:{best * 1e3:.2f}us({100 * ideal / best:.0f}%)")
        del g, col
    print(f"2^{logn}:", " ".join(out))
