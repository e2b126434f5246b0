#!/usr/bin/env bash
# A/B the build variants under shared_simd_scan_b200/csrc/build/variants/ (developer tool):
#   OPS=shared,linear KS=64,1024 WIDTHS=16-24 tools/ab_variants.sh [variant names...]
# prints one line per variant: fraction of the measured HBM peak per (op, w, K); CHECK=1 also runs the
# benchmarked-grid parity test against each variant.
cd "$(dirname "$0")/.."
names=("$@")
if [ ${#names[@]} -eq 0 ]; then names=(default $(ls shared_simd_scan_b200/csrc/build/variants/*.so | xargs -n1 basename | sed 's/\.so$//')); fi
for name in "${names[@]}"; do
  lib=""
  [ "$name" != default ] && lib="$PWD/shared_simd_scan_b200/csrc/build/variants/$name.so"
  echo "== $name"
  if [ -n "$CHECK" ]; then
    SSS_LIB_PATH=$lib python -m pytest tests/test_gpu_parity.py -q -x -k "${CHECK_K:-benchmarked_grid}" 2>&1 | tail -1
  fi
  SSS_LIB_PATH=$lib python tools/sweep_wk.py --ops ${OPS:-shared} --ks ${KS:-64} --widths ${WIDTHS:-17-17} --logn ${LOGN:-30} 2>&1 | python -c "
import sys,json
out=[]
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'w' in d: out.append('%s w%d K%d: %.3f' % (d['op'][:3], d['w'], d['K'], d['frac_of_measured_peak']))
print('  '.join(out))"
done
