# developer A/B: forced predicate-group sizes for K = 1024 (SSS_PRED_GROUP), one line per width
for w in ${WIDTHS:-9 17 23 32}; do
  for g in ${GROUPS_:-0 128 256 512}; do
    echo -n "w$w G=$g: "; SSS_PRED_GROUP=$g python tools/sweep_wk.py --ops shared --ks 1024 --widths $w-$w --logn ${LOGN:-29} 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'w' in d: print('%.3f' % d['frac_of_measured_peak'])"
  done
done
