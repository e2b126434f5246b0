#!/usr/bin/env bash
# A/B the unpack build variants under shared_simd_scan_b200/csrc/build/variants/ (developer tool)
for lib in "" shared_simd_scan_b200/csrc/build/variants/*.so; do
  name=${lib:-default}
  echo "== $name"
  SSS_LIB_PATH=${lib:+$PWD/$lib} python tools/quick_bench.py --ops unpack --widths ${WIDTHS:-1-32} --logn 28 2>&1 | python -c "
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    out.append('%d:%.3f'%(d['w'],d['frac']))
print(' '.join(out))"
done
