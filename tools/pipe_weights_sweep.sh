for w in "3,3,3,2,2,1,1,1" "4,4,3,2,2,1" "5,4,3,2,1,1" "6,5,3,1,1" "4,4,4,3,1" "8,4,2,1,1" "2,2,2,2,2,2,2,2" "6,6,3,1"; do
  echo "== $w"; PDC_PIPE_WEIGHTS=$w python bench.py --only-slots 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k in ('config4_slot_16cells_early_stop','config4_slot_16cells_fixed6'):
    l=d[k]['latency_us_c_abi']; print('  ',k,'p50',l['p50'],'min',l['min'],'p99',l['p99'])
"
done
