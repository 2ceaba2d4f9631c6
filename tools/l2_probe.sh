#!/bin/bash
# DRAM traffic of the decoder / dematcher per launch for several L2 set-aside sizes (ncu metrics pass), then timings.
for mb in 0 32 64 96; do
  echo "== PDC_L2_PERSIST_MB=$mb"
  PDC_L2_PERSIST_MB=$mb ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'ldpc_decode_h2|rate_dematch' -s 6 -c 2 --csv python bench.py --no-extras --steps 3 --warmup 3 2>/dev/null | grep -E '^"[0-9]' | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' 
done
for mb in 0 32 64 96; do
  PDC_L2_PERSIST_MB=$mb tools/ab_libs.sh 1 - | sed "s/^-/persist_mb=$mb/"
done
