#!/bin/bash
# Compares update-kernel variants across the workloads (GPU box).  usage: compare_update.sh variant...
for w in c3 c2 c4 c5; do
  for v in "$@"; do
    echo -n "$w $v: "
    T3D_UPDATE_VARIANT=$v python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=.*//'
  done
done
