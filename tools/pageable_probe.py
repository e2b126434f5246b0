"""Developer probe: sss_host_scan_range / sss_host_unpack from PAGEABLE host memory (the drop-in's real
input) at 17 bit, for the SSS_COPY_THREADS of the environment.  python tools/pageable_probe.py [log2_n]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n, w = 1 << logn, 17
m = (1 << w) - 1
col = api.generate_packed(n, w, seed=1)
nb = api.packed_bytes(n, w)
src = np.empty(nb, dtype=np.uint8)
src[:] = col[:nb].cpu().numpy()
out = np.zeros(api.bitvec_bytes(n), dtype=np.uint8)
api.host_scan(src, n, w, m // 4, m // 2, tail_granule=1, out=out)
t0 = time.perf_counter()
for _ in range(3):
    api.host_scan(src, n, w, m // 4, m // 2, tail_granule=1, out=out)
sec = (time.perf_counter() - t0) / 3
print(f"threads {os.environ.get('SSS_COPY_THREADS', 'default')}: scan from pageable {n * w / 8 / sec / 1e9:.1f} GB/s H2D", end="")
n2 = 1 << min(logn, 28)
dec = np.empty(n2, dtype=np.int32)
api.host_decompress(src, n2, w, out=dec)
t0 = time.perf_counter()
for _ in range(3):
    api.host_decompress(src, n2, w, out=dec)
sec = (time.perf_counter() - t0) / 3
print(f", decompress to pageable {n2 * 4 / sec / 1e9:.1f} GB/s D2H")
