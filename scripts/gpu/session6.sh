#!/bin/bash
# Round-2 GPU session 6: half tiles at the tail of a launch; mapped sprite staging; multi-context driver; full GPU suite.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s6; mkdir -p $O
timeout 1800 python -m pytest tests -q --maxfail=8 -m gpu -k "not headline_size and not alternative_paths" > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt
tail -25 $O/gpu.log
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { name=$1; shift; timeout 300 env "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$?" >> $O/rc.txt; scripts/bench_summary.py < $O/$name.json | tee -a $O/summary.txt; grep -h "resident CTA" $O/$name.err | head -2; }
for tt in 1 0; do
echo "== tail tiles $tt: c3 / c5 / c3d / c5 1.3% / c2 / c4 ==" | tee -a $O/summary.txt
run c3_t$tt T3D_TAIL_TILES=$tt $B --steps 30 --warmup 5
run c5_t$tt T3D_TAIL_TILES=$tt $B --workload c5 --steps 40
run c3d_t$tt T3D_TAIL_TILES=$tt $B --workload c3d --steps 30
run c5_13_t$tt T3D_TAIL_TILES=$tt $B --workload c5 --churn-spawn 262144 --churn-energy 500,2000 --steps 40
run c2_t$tt T3D_TAIL_TILES=$tt $B --workload c2 --steps 50
run c4_t$tt T3D_TAIL_TILES=$tt $B --workload c4 --steps 12
done
run c3_sep T3D_FUSED=0 $B --steps 20 --warmup 5
timeout 300 python scripts/bench_sprites.py 2>&1 | tee $O/sprites.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "headline_size" > $O/headline.log 2>&1; echo "headline rc=$?" | tee -a $O/rc.txt
timeout 1500 python -m pytest tests/test_gpu_variants.py -q --maxfail=4 > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/rc.txt
tail -5 $O/variants.log
ls -la $O
