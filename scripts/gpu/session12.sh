#!/bin/bash
# Round-2 GPU session 12: the whole GPU suite at HEAD, the driver's bench command, the reference arm.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s12; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/rc.txt; tail -3 $O/smoke.log
timeout 2400 python -m pytest tests -q --maxfail=8 -m gpu > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt
tail -12 $O/gpu.log | cut -c1-200
timeout 900 python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/rc.txt
scripts/bench_summary.py < $O/bench_n1.json | tee $O/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference rc=$?" | tee -a $O/rc.txt
scripts/bench_summary.py < $O/bench_reference.json | tee -a $O/summary.txt
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
timeout 300 env T3D_FUSED=0 $B --steps 20 --warmup 5 2>/dev/null | scripts/bench_summary.py | tee -a $O/summary.txt
ls -la $O
