#!/bin/bash
# Round-2 GPU session 8: the whole GPU suite at HEAD (depth sort included), the bench line with traffic, smoke.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s8; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/rc.txt; tail -3 $O/smoke.log
timeout 2400 python -m pytest tests -q --maxfail=8 -m gpu > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt
tail -15 $O/gpu.log
timeout 900 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/rc.txt
scripts/bench_summary.py < $O/bench.json | tee $O/summary.txt
ls -la $O
