#!/bin/bash
# A/B on one GPU box between prebuilt libraries: tools/ab_libs.sh rounds lib [lib ...] ("-" = the in-tree build).
# Per library: the headline step (bench.py --no-extras) and the resident config-3 / config-4 slots.
rounds=$1; shift
fmt='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "step %.3f ms  decode %.3f ms  dematch %.3f ms  e2e %.2f  parity %s" % (d["ms_per_step"] / d["config"]["launches_per_step"], d["roofline"]["ms_per_launch"], d["roofline_hbm"]["ms_per_launch"], d["e2e"]["value"], d["run"]["parity_vs_oracle_all_distinct_codeblocks"]))'
fmt2='import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], "  ".join("%s %.1f us%s" % (k.replace("_slot","").replace("cells","c").replace("cell","c"), v["us_per_slot"], "" if v.get("tb_crc_ok", True) else " TBFAIL") for k, v in d.items() if isinstance(v, dict) and "us_per_slot" in v and k.startswith("config")))'
for i in $(seq $rounds); do
  for lib in "$@"; do
    if [ "$lib" = "-" ]; then unset PDC_LIBRARY; else export PDC_LIBRARY=$lib; fi
    timeout 300 python bench.py --no-extras --steps 10 --launches 1 2>&1 | tail -1 | python -c "$fmt" "$lib"
    [ -n "$AB_SLOTS" ] && timeout 600 python bench.py --only-slots 2>&1 | tail -1 | python -c "$fmt2" "$lib"
  done
done
