#!/usr/bin/env python
"""Developer probe: BASELINE configs[0] (2^24 x 11 bit, 100 scans in a CUDA graph / in a stream) with the
build's launch settings; run once with SSS_PDL=0 and once with SSS_PDL=1 (read at context creation)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from shared_simd_scan_b200 import api  # noqa: E402

peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
r = bench.measure_cfg1(torch, api, peak)
print("SSS_PDL=%s" % os.environ.get("SSS_PDL", "(default)"), json.dumps({k: r[k] for k in ("us_per_scan_graph", "us_per_scan_stream", "pct_of_ideal_graph")}))
