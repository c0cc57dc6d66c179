#!/bin/bash
# Round-2 GPU session 17: side-stage slots by rank among the warp's movers (second movers through the stage); parity, then churn.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s17; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_variants.py tests/test_gpu_fuzz.py tests/test_gpu_handoff.py -q -x -m gpu > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt; tail -5 $O/gpu.log | cut -c1-200
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
for w in "c5 100" "c3d 30" "c3 20" "c5 100" "c2 100"; do set -- $w
  echo -n "$1: " | tee -a $O/summary.txt
  timeout 300 $B --workload $1 --steps $2 --warmup 5 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
done
echo -n "c5 heavy (energies 30..120 ms, ~20 % replaced per frame): " | tee -a $O/summary.txt
timeout 300 $B --workload c5 --churn-energy 30,120 --steps 100 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
echo -n "separate c5: " | tee -a $O/summary.txt
timeout 300 env T3D_FUSED=0 $B --workload c5 --steps 100 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
