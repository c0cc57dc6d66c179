#!/bin/bash
# Round-2 GPU session 15: Font::drawText on the device (parity, measurement), the pipelined frame with two launches per frame.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s15; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_sprites.py -q -x -m gpu > $O/sprites.log 2>&1; echo "sprites rc=$?" | tee -a $O/rc.txt; tail -25 $O/sprites.log | cut -c1-220
timeout 900 python -m pytest tests/test_gpu_variants.py -q -x -m gpu -k "update_ticket" > $O/variants.log 2>&1; echo "variants rc=$?" | tee -a $O/rc.txt; tail -5 $O/variants.log | cut -c1-220
timeout 600 python scripts/bench_sprites.py > $O/sprites_2d.txt 2> $O/sprites_2d.err; echo "bench_sprites rc=$?" | tee -a $O/rc.txt; cat $O/sprites_2d.txt; tail -5 $O/sprites_2d.err
