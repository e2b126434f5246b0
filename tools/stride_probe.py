#!/usr/bin/env python
"""Developer probe: does the caller-chosen bitvector stride of sss_shared_scan matter?

    python tools/stride_probe.py [w=17] [K=64]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

w = int(sys.argv[1]) if len(sys.argv) > 1 else 17
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n = (1 << 30) if K <= 64 else (1 << 36) // K // 8
m = (1 << w) - 1
packed = api.generate_packed(n, w, seed=42 + w)
g = torch.Generator().manual_seed(42 + K)
lo = torch.randint(0, m + 1, (K,), generator=g)
hi = torch.clamp(lo + torch.randint(0, 2 * (m + 1) // K, (K,), generator=g), max=m)
lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
cnt = torch.empty(K, dtype=torch.int64, device="cuda")
base = n // 8
arena = torch.empty(K * (base + (1 << 21)) + 4096, dtype=torch.uint8, device="cuda")
for pad in (0, 32, 128, 256, 512, 1024, 4096, 4096 + 128, 65536, 65536 + 4096 + 128, (1 << 20) + 128, 2 * (1 << 20) - 128):
    stride = base + pad
    out = arena[: K * stride].view(K, stride)
    f = lambda: api.shared_scan(packed, n, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
    for _ in range(2):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        f()
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    print(f"stride = n/8 + {pad:8d}: {best:.3f} ms  {n / best / 1e6:7.1f} G codes/s", flush=True)
