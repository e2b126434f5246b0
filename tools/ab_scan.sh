#!/usr/bin/env bash
# A/B the scan kernel build variants under shared_simd_scan_b200/csrc/build/variants/ (developer tool)
set -e
for lib in "" shared_simd_scan_b200/csrc/build/variants/*.so; do
  name=${lib:-default}
  echo "== $name"
  SSS_LIB_PATH=${lib:+$PWD/$lib} python tools/quick_bench.py --ops scan --widths ${WIDTHS:-1-16} --logn 30 2>&1 | python -c "
import sys,json
out=[]
for l in sys.stdin:
    d=json.loads(l); out.append('%d:%.3f'%(d['w'],d['frac']))
print(' '.join(out))"
done
