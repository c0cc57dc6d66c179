#!/bin/bash
# Round-2 GPU session 5: sprites / tile sets, multi-context driver, lean host path.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s5; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sprites.py tests/test_gpu_multi_context.py -q --maxfail=6 -m gpu > $O/new.log 2>&1; echo "new rc=$?" | tee -a $O/rc.txt
tail -40 $O/new.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_handoff.py tests/test_gpu_fuzz.py tests/test_gpu_facade.py tests/test_engine_headers.py -q --maxfail=8 -m gpu -k "not headline_size" > $O/parity.log 2>&1; echo "parity rc=$?" | tee -a $O/rc.txt
tail -12 $O/parity.log
timeout 300 python scripts/bench_sprites.py 2>&1 | tee $O/sprites.txt
for w in "c2 200 1" "c4 30 8"; do timeout 300 python scripts/host_cost.py $w 2>&1 | tee -a $O/host_cost.txt; done
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; }
echo "== c2 ==" | tee -a $O/summary.txt;      run c2 $B --workload c2 --steps 50
echo "== c4 as rank 0 of 8 ==" | tee -a $O/summary.txt; run c4w8 T3D_BENCH_EMULATE_WORLD=8 $B --workload c4 --steps 30
ls -la $O
