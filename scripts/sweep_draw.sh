#!/bin/bash
# Sweeps sub-tiles per draw CTA on two workloads (GPU box).
for w in c3 c4 c2; do
  for d in 1 2 4 8; do
    echo -n "$w draw_sub=$d: "
    T3D_DRAW_SUB=$d python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=.*//'
  done
done
