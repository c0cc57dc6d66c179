#!/bin/bash
# Fused against separate launches as a function of churn (GPU box): 256 Ki particles spawned per frame, lifetimes scaled.
for en in "125,500" "250,1000" "500,2000" "1000,4000"; do
  for f in 0 1; do
    echo -n "energy=$en fused=$f: "
    T3D_FUSED=$f python bench.py --workload c5 --churn-spawn 262144 --churn-energy $en --steps 40 --warmup 260 --no-cpu-baseline --no-also 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=[^ ]* launches=[^ ]*//; s/ clocks=.*//'
  done
done
