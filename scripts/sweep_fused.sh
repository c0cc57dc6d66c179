#!/bin/bash
# Fused update+draw variants on a workload (GPU box).  usage: sweep_fused.sh [workload]
w=${1:-c3}
echo -n "$w separate: "; python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=.*//'
for v in ${FUSED_VARIANTS:-g2t192s12b2 g2t192s8b2 g2t384s6b1 g4t128s8b2 f4t64s10b2}; do
  echo -n "$w $v: "
  T3D_FUSED=1 T3D_FUSED_VARIANT=$v timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --no-also 2>/dev/null | scripts/bench_summary.py | head -1 | sed 's/ e2e=.*//'
done
