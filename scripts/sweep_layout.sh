#!/bin/bash
# Compares particle-storage layouts (plain SoA vs blocked SoA of various block sizes) on a workload (GPU box).
W=${1:-c3}
for b in 0 128 256 512 1024 2048 4096; do
  echo -n "block=$b  "
  T3D_LAYOUT_BLOCK=$b python bench.py --workload $W --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | scripts/bench_summary.py | head -1
done
