#!/usr/bin/env python
"""Driver for per-family ncu captures: runs ONE op a few times so that `ncu -k regex:<kernel> -s 2 -c 1`
profiles a warm launch.

    python tools/ncu_family_driver.py unpack|pack|linear1024|scan|scanpos|shared:K|linear:K|positions [w] [log2_n]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shared_simd_scan_b200 import api  # noqa: E402

op = sys.argv[1]
w = int(sys.argv[2]) if len(sys.argv) > 2 else 17
n = 1 << (int(sys.argv[3]) if len(sys.argv) > 3 else 28)
m = (1 << w) - 1
packed = api.generate_packed(n, w, seed=42 + w)
if op == "unpack":
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    f = lambda: api.decompress(packed, n, w, out=out)  # noqa: E731
elif op == "pack":
    codes = api.decompress(packed, n, w).view(torch.int32)
    out = torch.empty_like(packed)
    f = lambda: api.compress(codes, w, out=out)  # noqa: E731
elif op == "scan":
    out = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    cnt = torch.empty(1, dtype=torch.int64, device="cuda")
    f = lambda: api.scan(packed, n, w, m // 4, m // 2, out=out, count=cnt)  # noqa: E731
elif op == "linear1024":
    K = 1024
    g = torch.Generator().manual_seed(42 + K)
    lo = torch.randint(0, m + 1, (K,), generator=g)
    hi = torch.clamp(lo + torch.randint(0, max(2, 2 * (m + 1) // K), (K,), generator=g), max=m)
    lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
    out = torch.empty(K * (n // 8), dtype=torch.uint8, device="cuda")
    cnt = torch.empty(K, dtype=torch.int64, device="cuda")
    f = lambda: api.shared_scan_linear(packed, n, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
elif op.startswith("shared:") or op.startswith("linear:"):
    K = int(op.split(":")[1])
    g = torch.Generator().manual_seed(42 + K)
    lo = torch.randint(0, m + 1, (K,), generator=g)
    hi = torch.clamp(lo + torch.randint(0, max(2, 2 * (m + 1) // K), (K,), generator=g), max=m)
    lo_d, hi_d = lo.to(torch.int32).cuda(), hi.to(torch.int32).cuda()
    cnt = torch.empty(K, dtype=torch.int64, device="cuda")
    if op.startswith("shared:"):
        out = torch.empty((K, (api.bitvec_bytes(n) + 127) // 128 * 128), dtype=torch.uint8, device="cuda")
        f = lambda: api.shared_scan(packed, n, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
    else:
        out = torch.empty((n // 8, K), dtype=torch.uint8, device="cuda")
        f = lambda: api.shared_scan_linear(packed, n, w, lo_d, hi_d, out=out, counts=cnt)  # noqa: E731
elif op == "positions":
    bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    c1 = torch.empty(1, dtype=torch.int64, device="cuda")
    sel = float(os.environ.get("SEL", "0.01"))
    api.scan(packed, n, w, 0, max(0, int(sel * (m + 1)) - 1), out=bv, count=c1)
    hits = int(c1.item())
    pos = torch.empty(max(hits, 1), dtype=torch.int64, device="cuda")
    num = torch.empty(1, dtype=torch.int64, device="cuda")
    f = lambda: api.positions(bv, n, capacity=hits, out=pos, num=num)  # noqa: E731
elif op == "scanpos":  # the fused call: scan (fills the list's segment counts) + group offsets + expansion
    bv = torch.empty(api.bitvec_bytes(n), dtype=torch.uint8, device="cuda")
    c1 = torch.empty(1, dtype=torch.int64, device="cuda")
    sel = float(os.environ.get("SEL", "0.01"))
    hi_v = max(0, int(sel * (m + 1)) - 1)
    api.scan(packed, n, w, 0, hi_v, out=bv, count=c1)
    hits = int(c1.item())
    pos = torch.empty(max(hits, 1), dtype=torch.int64, device="cuda")
    f = lambda: api.scan_positions(packed, n, w, 0, hi_v, capacity=hits, out=bv, pos_out=pos, count=c1)  # noqa: E731
else:
    raise SystemExit("unknown op " + op)
for _ in range(4):
    f()
torch.cuda.synchronize()
print("ok", op, w, n)
