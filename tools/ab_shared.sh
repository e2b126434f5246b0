#!/usr/bin/env bash
# A/B the shared-scan build variants under shared_simd_scan_b200/csrc/build/variants/ (developer tool)
for lib in "" shared_simd_scan_b200/csrc/build/variants/*.so; do
  name=${lib:-default}
  echo "== $name"
  SSS_LIB_PATH=${lib:+$PWD/$lib} python tools/sweep_wk.py --ops ${OPS:-shared,linear} --ks ${KS:-8,64} --widths ${WIDTHS:-9-9} 2>&1 | python -c "
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'w' in d: out.append('%s w%d K%d: %.3f' % (d['op'][:3], d['w'], d['K'], d['frac_of_measured_peak']))
print('  '.join(out))"
done
