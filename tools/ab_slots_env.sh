#!/bin/bash
# A/B of the single-slot legs through environment switches of the library: each argument is one "VAR=value[,VAR=value]"
# set ("-" = defaults); bench.py --only-slots runs once per set and per round and the resident chain time and the C-ABI
# latency of every slot leg are printed. Usage: tools/ab_slots_env.sh rounds set [set ...]
rounds=$1; shift
fmt='
import json, sys
d = json.loads(sys.stdin.read())
def walk(o, path):
    if isinstance(o, dict):
        if "latency_us_c_abi" in o:
            l = o["latency_us_c_abi"]
            print("  %-44s chain %7.1f us   c-abi p50 %7.1f  min %7.1f  p99 %7.1f" % (path, o.get("us_per_slot", float("nan")), l.get("p50", -1), l.get("min", -1), l.get("p99", -1)))
        for k, v in o.items():
            walk(v, k)
walk(d, "")
'
for i in $(seq $rounds); do
  for set in "$@"; do
    envs=$(echo "$set" | tr ',' ' ')
    [ "$set" = "-" ] && envs=""
    echo "== $set"
    env $envs timeout 600 python bench.py --only-slots 2>&1 | tail -1 | python -c "$fmt"
  done
done
