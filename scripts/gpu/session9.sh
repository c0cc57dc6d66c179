#!/bin/bash
# Round-2 GPU session 9: spawn inside the update launch -- parity first, then what it buys.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s9; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/rc.txt; tail -3 $O/smoke.log
timeout 2400 python -m pytest tests -q --maxfail=10 -m gpu -k "not alternative_paths and not headline_size" > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt
tail -40 $O/gpu.log | cut -c1-250
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; grep -h "resident CTA" $O/$name.err | head -2; }
for f in 1 0; do
echo "== fuse spawn $f: c5 / c3d / c5 1.3% ==" | tee -a $O/summary.txt
run c5_s$f T3D_FUSE_SPAWN=$f $B --workload c5 --steps 40
run c3d_s$f T3D_FUSE_SPAWN=$f $B --workload c3d --steps 30
run c5_13_s$f T3D_FUSE_SPAWN=$f $B --workload c5 --churn-spawn 262144 --churn-energy 500,2000 --steps 40
done
run c3 $B --steps 20 --warmup 5
ls -la $O
