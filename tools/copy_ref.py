#!/usr/bin/env python
"""Reference points for the roofline denominator: torch device-to-device copy bandwidth (read+write
bytes) at several sizes, timed like tools/quick_bench.py (best of 5, CUDA events)."""
import json
import torch

for mib in (32, 64, 128, 256, 512, 1024, 2048, 4096):
    n = mib << 20
    a = torch.empty(n, dtype=torch.uint8, device="cuda")
    b = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        b.copy_(a)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.copy_(a); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    print(json.dumps({"copy_mib_each_way": mib, "us": round(best * 1e6, 1), "gbs": round(2 * n / best / 1e9, 1)}))
