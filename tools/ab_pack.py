#!/usr/bin/env python
"""Developer A/B: compressor throughput per width for the library named by SSS_LIB_PATH."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

n = 1 << 28
out = []
for w in (1, 5, 9, 13, 17, 23, 32):
    codes = torch.randint(0, 1 << min(w, 31), (n,), dtype=torch.int32, device="cuda")
    packed = torch.empty(api.packed_bytes(n, w), dtype=torch.uint8, device="cuda")
    for _ in range(2):
        api.compress(codes, w, out=packed)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        api.compress(codes, w, out=packed)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / 5
    out.append("%d:%.3f" % (w, n * (4 + w / 8) / ms / 1e6 / 6550.7))
    del codes, packed
print(" ".join(out))
