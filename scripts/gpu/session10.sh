#!/bin/bash
# Round-2 GPU session 10: which shape should the update kernel WITHOUT the draw use (T3D_FUSED=0 / update-only frames)?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s10; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; grep -h "resident CTA" $O/$name.err | head -2; }
for v in u2t192s12b2 u2t128s12b4 u2t256s8b2 u4t128s8b2; do
  echo "== $v: c3 / c5 / c2 ==" | tee -a $O/summary.txt
  run c3_$v T3D_FUSED=0 T3D_UPDATE_VARIANT=$v $B --steps 20 --warmup 5
  run c5_$v T3D_FUSED=0 T3D_UPDATE_VARIANT=$v $B --workload c5 --steps 40
  run c2_$v T3D_FUSED=0 T3D_UPDATE_VARIANT=$v $B --workload c2 --steps 50
done
