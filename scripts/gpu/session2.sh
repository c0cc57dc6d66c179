#!/bin/bash
# Round-2 GPU session 2: look-back windows, mover prefetch, tile-shape heuristic; new hand-off tests; variants.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s2; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_handoff.py tests/test_gpu_fuzz.py tests/test_gpu_facade.py -q --maxfail=8 -k "not headline_size" > $O/parity.log 2>&1; echo "parity rc=$?" | tee -a $O/rc.txt
tail -30 $O/parity.log
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; }
echo "== c3 ==" | tee -a $O/summary.txt;      run c3 $B --steps 30 --warmup 5
echo "== c5 ==" | tee -a $O/summary.txt;      run c5 $B --workload c5 --steps 40
echo "== c3d ==" | tee -a $O/summary.txt;     run c3d $B --workload c3d --steps 30
echo "== c5 1.3% ==" | tee -a $O/summary.txt; run c5_13 $B --workload c5 --churn-spawn 262144 --churn-energy 500,2000 --steps 40
echo "== c5 0.34% ==" | tee -a $O/summary.txt; run c5_034 $B --workload c5 --churn-spawn 131072 --churn-energy 2000,8000 --steps 30
echo "== c2 / c4 ==" | tee -a $O/summary.txt
run c2 $B --workload c2 --steps 50
run c4 $B --workload c4 --steps 12
echo "== separate ==" | tee -a $O/summary.txt
run c3_sep T3D_FUSED=0 $B --steps 20 --warmup 5
run c5_sep T3D_FUSED=0 $B --workload c5 --steps 40
run c5_13_sep T3D_FUSED=0 $B --workload c5 --churn-spawn 262144 --churn-energy 500,2000 --steps 40
timeout 1200 python -m pytest tests/test_gpu_variants.py -q --maxfail=4 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/rc.txt
tail -15 $O/variants.log
