#!/bin/bash
# experiment: how much do the look-back waits of the uncertain tiles cost a population in which nobody dies?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s7; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; }
for d in 0 1 0 1; do
run c3_d$d T3D_DEBUG_ASSUME_NO_DEATHS=$d $B --steps 30 --warmup 5
run c4_d$d T3D_DEBUG_ASSUME_NO_DEATHS=$d $B --workload c4 --steps 12
run c2_d$d T3D_DEBUG_ASSUME_NO_DEATHS=$d $B --workload c2 --steps 50
done
