#!/bin/bash
# Round-2 GPU session 16: the whole GPU suite at HEAD, then the final single-GPU measurements (scripts/gpu/final_n1.sh).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s16; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/rc.txt; tail -3 $O/smoke.log
timeout 2400 python -m pytest tests -q --maxfail=8 -m gpu > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt
tail -12 $O/gpu.log | cut -c1-200
bash scripts/gpu/final_n1.sh
