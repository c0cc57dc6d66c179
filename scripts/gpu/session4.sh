#!/bin/bash
# Round-2 GPU session 4: facade / properties / engine-header tests, the 64 Mi headline test, host-time breakdown, spawn specialisation.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s4; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_handoff.py tests/test_gpu_fuzz.py tests/test_gpu_facade.py tests/test_engine_headers.py -q --maxfail=8 -m gpu -k "not headline_size" > $O/parity.log 2>&1; echo "parity rc=$?" | tee -a $O/rc.txt
tail -30 $O/parity.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "headline_size" > $O/headline.log 2>&1; echo "headline rc=$?" | tee -a $O/rc.txt
tail -5 $O/headline.log
for w in "c2 200 1" "c4 30 8" "c5 60 1" "c3 30 1"; do timeout 300 python scripts/host_cost.py $w 2>&1 | tee -a $O/host_cost.txt; done
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; grep -h "resident CTA" $O/$name.err | head -2; }
echo "== c5 ==" | tee -a $O/summary.txt;      run c5 $B --workload c5 --steps 40
echo "== c2 ==" | tee -a $O/summary.txt;      run c2 $B --workload c2 --steps 50
echo "== c4 as rank 0 of 8 ==" | tee -a $O/summary.txt; run c4w8 T3D_BENCH_EMULATE_WORLD=8 $B --workload c4 --steps 30
N="ncu --set full --import-source on --clock-control none"
timeout 600 $N -k regex:k_spawn --launch-skip 34 -c 1 -o $O/prof_spawn $B --workload c5 --steps 2 > $O/ncu_spawn.log 2>&1
ls -la $O
