#!/bin/bash
# A/B on one GPU box through environment switches of the library: each argument is one "VAR=value[,VAR=value]" set
# ("-" = defaults); bench.py --no-extras runs once per set and per round. Usage: tools/ab_env.sh rounds set [set ...]
rounds=$1; shift
fmt='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "step %.3f ms  decode %.3f ms  dematch %.3f ms  e2e %.2f  parity %s" % (d["ms_per_step"] / d["config"]["launches_per_step"], d["roofline"]["ms_per_launch"], d["roofline_hbm"]["ms_per_launch"], d["e2e"]["value"], d["run"]["parity_vs_oracle_all_distinct_codeblocks"]))'
for i in $(seq $rounds); do
  for set in "$@"; do
    envs=$(echo "$set" | tr ',' ' ')
    [ "$set" = "-" ] && envs=""
    env $envs timeout 300 python bench.py --no-extras --steps 20 --launches 1 2>&1 | tail -1 | python -c "$fmt" "$set"
  done
done
