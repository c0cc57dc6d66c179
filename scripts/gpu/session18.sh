#!/bin/bash
# Round-2 GPU session 18: the whole GPU suite at HEAD (after the spawn / draw descriptor heads and the mover slots), sprites + text bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s18; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/rc.txt; tail -3 $O/smoke.log
timeout 2400 python -m pytest tests -q --maxfail=8 -m gpu > $O/gpu.log 2>&1; echo "gpu rc=$?" | tee -a $O/rc.txt
tail -6 $O/gpu.log | cut -c1-200
timeout 600 python scripts/bench_sprites.py > $O/sprites_2d.txt 2> $O/sprites_2d.err; echo "bench_sprites rc=$?" | tee -a $O/rc.txt; tail -9 $O/sprites_2d.txt; tail -5 $O/sprites_2d.err
