#!/bin/bash
# Round-2 GPU session 3: movers' records and extras through the async stages (one group ahead); ncu of c3 / c5 / spawn.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s3; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_handoff.py tests/test_gpu_fuzz.py tests/test_gpu_facade.py -q --maxfail=8 -k "not headline_size" > $O/parity.log 2>&1; echo "parity rc=$?" | tee -a $O/rc.txt
tail -30 $O/parity.log
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; grep -h "resident CTA" $O/$name.err | head -2; }
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
timeout 1200 python -m pytest tests/test_gpu_variants.py -q --maxfail=4 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/rc.txt
tail -5 $O/variants.log
# ncu: fused kernel on c3 and c5, spawn kernel on c5
N="ncu --set full --import-source on --clock-control none"
timeout 600 $N -k regex:k_update --launch-skip 4 -c 1 -o $O/prof_c3 $B --steps 2 --warmup 3 > $O/ncu_c3.log 2>&1
timeout 600 $N -k regex:k_update --launch-skip 34 -c 1 -o $O/prof_c5 $B --workload c5 --steps 2 > $O/ncu_c5.log 2>&1
timeout 600 $N -k regex:k_spawn --launch-skip 34 -c 1 -o $O/prof_spawn $B --workload c5 --steps 2 > $O/ncu_spawn.log 2>&1
timeout 600 $N -k regex:k_update --launch-skip 8 -c 1 -o $O/prof_c2 $B --workload c2 --steps 2 > $O/ncu_c2.log 2>&1
ls -la $O
