#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/spawn; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "golden or device_rng or churn or libc or stream or capacity" > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt; tail -3 $O/gpu.log | cut -c1-200
for w in c5 c3d c5; do
  timeout 300 $B --workload $w --steps 60 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
done
