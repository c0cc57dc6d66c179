#!/bin/bash
# Round-2 GPU session 14: the launch descriptor fetched in one round trip (guess + check); parity, then the bench lines it can move.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s14; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_fuzz.py -q -x -m gpu > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt; tail -5 $O/gpu.log | cut -c1-200
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
for w in "c2 100" "c4 12" "c3 20" "c5 100" "c3d 20"; do set -- $w
  echo -n "$1: " | tee -a $O/summary.txt
  timeout 300 $B --workload $1 --steps $2 --warmup 5 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
done
for w in "c3 20" "c2 100"; do set -- $w
  echo -n "separate $1: " | tee -a $O/summary.txt
  timeout 300 env T3D_FUSED=0 $B --workload $1 --steps $2 --warmup 5 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
done
echo -n "f2t128s8b3 c2: " | tee -a $O/summary.txt
timeout 300 env T3D_FUSED_VARIANT=f2t128s8b3 $B --workload c2 --steps 100 --warmup 5 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
