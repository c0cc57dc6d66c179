#!/bin/bash
# Round-2 GPU session 1: first contact of the new storage layout + rewritten update kernel with the hardware.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; SRC=$?; echo "smoke rc=$SRC" | tee -a $O/rc.txt
tail -15 $O/smoke.log
if [ $SRC -ne 0 ]; then
  # first contact failed: collect what a debugging session needs and stop (no point benchmarking a broken kernel)
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "golden or multi_tile_update_and or device_rng or removal_patterns" --maxfail=6 > $O/parity.log 2>&1
  tail -60 $O/parity.log
  exit 1
fi
timeout 1500 python -m pytest tests/test_gpu_parity.py -q --maxfail=5 -k "not headline_size" > $O/parity.log 2>&1; PRC=$?; echo "parity rc=$PRC" | tee -a $O/rc.txt
tail -40 $O/parity.log
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; }
echo "== c3 fused ==" | tee -a $O/summary.txt;      run c3_fused T3D_FUSED=1 $B --steps 30 --warmup 5
echo "== c3 separate ==" | tee -a $O/summary.txt;   run c3_sep T3D_FUSED=0 $B --steps 30 --warmup 5
for f in 1 0; do
  echo "== c5 fused=$f ==" | tee -a $O/summary.txt;   run c5_f$f T3D_FUSED=$f $B --workload c5 --steps 40
  echo "== c3d fused=$f ==" | tee -a $O/summary.txt;  run c3d_f$f T3D_FUSED=$f $B --workload c3d --steps 30
  echo "== c5 1.3% fused=$f ==" | tee -a $O/summary.txt; run c5_13_f$f T3D_FUSED=$f $B --workload c5 --churn-spawn 262144 --churn-energy 500,2000 --steps 40
  echo "== c5 0.34% fused=$f ==" | tee -a $O/summary.txt; run c5_034_f$f T3D_FUSED=$f $B --workload c5 --churn-spawn 131072 --churn-energy 2000,8000 --steps 30
done
echo "== c2 / c4 ==" | tee -a $O/summary.txt
run c2 T3D_FUSED=1 $B --workload c2 --steps 50
run c4 T3D_FUSED=1 $B --workload c4 --steps 12
# fused variants on c3
for v in f2t192s6b2 f2t128s12b3; do echo "== c3 $v ==" | tee -a $O/summary.txt; run c3_$v T3D_FUSED=1 T3D_FUSED_VARIANT=$v $B --steps 20 --warmup 5; done
for v in u2t256s8b2 u2t128s12b4 u4t128s8b2; do echo "== c3 sep $v ==" | tee -a $O/summary.txt; run c3_$v T3D_FUSED=0 T3D_UPDATE_VARIANT=$v $B --steps 20 --warmup 5; done
# ncu: fused kernel on c3 and on c5
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_update --launch-skip 4 -c 1 -o $O/prof_c3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-also --no-vertices-to-host > $O/ncu_c3.log 2>&1
timeout 600 env T3D_FUSED=1 ncu --set full --import-source on --clock-control none -k regex:k_update --launch-skip 34 -c 1 -o $O/prof_c5 python bench.py --workload c5 --steps 2 --no-cpu-baseline --no-also --no-vertices-to-host > $O/ncu_c5.log 2>&1
timeout 600 env T3D_FUSED=0 ncu --set full --import-source on --clock-control none -k regex:k_update --launch-skip 34 -c 1 -o $O/prof_c5_sep python bench.py --workload c5 --steps 2 --no-cpu-baseline --no-also --no-vertices-to-host > $O/ncu_c5_sep.log 2>&1
ls -la $O
