#!/bin/bash
# Fused against separate launches as a function of churn (GPU box): SPAWN particles spawned per frame, lifetimes scaled.
# usage: churn_crossover.sh "spawn:emin,emax:warmup" ...
for cfg in "$@"; do
  IFS=: read sp en wu <<< "$cfg"
  for f in 0 1; do
    echo -n "spawn=$sp energy=$en fused=$f: "
    T3D_FUSED=$f python bench.py --workload c5 --churn-spawn $sp --churn-energy $en --steps 40 --warmup $wu --no-cpu-baseline --no-also 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=[^ ]* launches=[^ ]*//; s/ clocks=.*//'
  done
done
