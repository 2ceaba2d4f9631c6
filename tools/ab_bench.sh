#!/bin/bash
# A/B measurement on one GPU box: alternates bench.py between a reference build of the library (_ab/lib_old.so) and the
# in-tree build. Usage: tools/ab_bench.sh [rounds] [extra bench.py args]
rounds=${1:-2}; shift
fmt='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "step %.3f ms  decode %.3f ms  dematch %.3f ms  e2e %.2f" % (d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline_hbm"]["ms_per_launch"], d["e2e"]["value"]))'
for i in $(seq $rounds); do
  PDC_LIBRARY=_ab/lib_old.so timeout 300 python bench.py --no-extras "$@" 2>&1 | tail -1 | python -c "$fmt" old
  timeout 300 python bench.py --no-extras "$@" 2>&1 | tail -1 | python -c "$fmt" new
done
