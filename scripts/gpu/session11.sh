#!/bin/bash
# Round-2 GPU session 11: new tests (batch_frame, storage reuse), variant experiment for the update kernel without the draw, host cost.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s11; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_facade.py tests/test_gpu_parity.py -q -m gpu -k "short_lived or batch_of_mixed" > $O/new.log 2>&1; echo "new rc=$?" | tee -a $O/rc.txt; tail -15 $O/new.log | cut -c1-200
bash scripts/gpu/session10.sh
cp -r gpurun_out/s10 $O/variants 2>/dev/null
for w in "c2 200 1" "c4 30 8"; do timeout 300 python scripts/host_cost.py $w 2>&1 | tee -a $O/host_cost.txt; done
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
timeout 300 env T3D_BENCH_EMULATE_WORLD=8 $B --workload c4 --steps 30 2>/dev/null | scripts/bench_summary.py | tee $O/c4w8.txt
timeout 300 $B --workload c2 --steps 50 2>/dev/null | scripts/bench_summary.py | tee $O/c2.txt
