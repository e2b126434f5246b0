#!/usr/bin/env python
"""One fused-scan launch per bit width at the bench size, for an ncu DRAM-traffic capture:
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
      -k 'regex:scan_(range|direct)_kernel' --csv --log-file gpurun_out/scan_traffic.csv python tools/scan_traffic.py
The result is summarised into profiles/ by tools/summarise_traffic.py and read by bench.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n = 1 << logn
out = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
cnt = torch.empty(1, dtype=torch.int64, device="cuda")
for w in range(1, 33):
    m = (1 << w) - 1
    col = api.generate_packed(n, w, seed=42 + w)
    api.scan(col, n, w, m // 4, m // 2, out=out, count=cnt)
    torch.cuda.synchronize()
    del col
print("done")
