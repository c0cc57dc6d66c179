#!/bin/bash
# Update-kernel variants (two-launch path) at several churn levels (GPU box).  usage: compare_update_churn.sh variant...
for v in "$@"; do
  echo -n "$v c3 (no churn): "
  T3D_FUSED=0 T3D_UPDATE_VARIANT=$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=.*//'
  for cfg in 262144:1000,4000:260 262144:500,2000:140 1048576:50,400:30; do
    IFS=: read sp en wu <<< "$cfg"
    echo -n "$v spawn=$sp energy=$en: "
    T3D_FUSED=0 T3D_UPDATE_VARIANT=$v python bench.py --workload c5 --churn-spawn $sp --churn-energy $en --steps 40 --warmup $wu --no-cpu-baseline --no-also 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=[^ ]* launches=[^ ]*//; s/ clocks=.*//'
  done
done
