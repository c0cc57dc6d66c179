#!/bin/bash
# Round-2 final single-GPU measurements: the driver's bench command, the reference arm, ncu launch list and full captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt 2>&1
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/rc.txt
scripts/bench_summary.py < $O/bench_n1.json | tee $O/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" | tee -a $O/rc.txt
scripts/bench_summary.py < $O/bench_reference.json | tee -a $O/summary.txt
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
# launch list of the bench command (short: 2 timed steps), every kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-vertices-to-host > $O/ncu_launches.log 2>&1; echo "launches rc=$?" | tee -a $O/rc.txt
N="ncu --set full --import-source on --clock-control none"
timeout 600 $N -k regex:k_update --launch-skip 4 -c 1 -o $O/prof_c3 $B --steps 2 --warmup 3 > $O/ncu_c3.log 2>&1
timeout 600 $N -k regex:k_update --launch-skip 34 -c 1 -o $O/prof_c5 $B --workload c5 --steps 2 > $O/ncu_c5.log 2>&1
timeout 600 $N -k regex:k_spawn --launch-skip 34 -c 1 -o $O/prof_spawn $B --workload c5 --steps 2 > $O/ncu_spawn.log 2>&1
timeout 600 $N -k regex:k_update --launch-skip 104 -c 1 -o $O/prof_c3d $B --workload c3d --steps 2 > $O/ncu_c3d.log 2>&1
timeout 600 $N -k regex:k_update --launch-skip 8 -c 1 -o $O/prof_c2 $B --workload c2 --steps 2 > $O/ncu_c2.log 2>&1
timeout 600 $N -k regex:k_update --launch-skip 4 -c 1 -o $O/prof_c4 $B --workload c4 --steps 2 > $O/ncu_c4.log 2>&1
timeout 600 env T3D_FUSED=0 $N -k regex:"k_update|k_draw" --launch-skip 8 -c 2 -o $O/prof_c3_sep $B --steps 2 --warmup 3 > $O/ncu_c3_sep.log 2>&1
timeout 300 python scripts/host_cost.py c2 200 1 2>&1 | tee $O/host_cost.txt
timeout 300 python scripts/host_cost.py c4 30 8 2>&1 | tee -a $O/host_cost.txt
timeout 300 python scripts/bench_sprites.py 2>&1 | tee $O/sprites.txt
# the reports are too big to travel (64 MiB limit): keep their raw metrics and the source-level stall summaries
for r in $O/prof_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null
  python scripts/ncu_top_stalls.py $r 40 > ${r%.ncu-rep}_stalls.txt 2>/dev/null
  rm -f $r
done
ls -la $O
