#!/bin/bash
# Fused launch against two launches as a function of churn (profiles/r2_churn.txt), and the spawn kernel alone.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/churn; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
if [ -n "$T3D_SPAWN_TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_facade.py -q -x -m gpu > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt; tail -4 $O/gpu.log | cut -c1-200
fi
for f in 1 0; do
  for w in "c5" "c3d" "c5 --churn-spawn 262144 --churn-energy 500,2000" "c5 --churn-spawn 131072 --churn-energy 2000,8000"; do
    echo "== fused=$f $w ==" | tee -a $O/summary.txt
    timeout 300 env T3D_FUSED=$f $B --workload $w --steps 30 2>/dev/null | scripts/bench_summary.py | head -1 | tee -a $O/summary.txt
  done
done
