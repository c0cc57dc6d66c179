#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/text_tile; mkdir -p $O
V=$PWD/tractor3d_b200/csrc/_build/libt3d_particles_text256.so   # python scripts/build_variants.py text256
for lib in "" $V; do
  echo "== library: ${lib:-product} ==" | tee -a $O/summary.txt
  timeout 600 env T3D_LIBRARY=$lib python -m pytest tests/test_gpu_sprites.py -q -x -m gpu -k "text or fixtures" 2>&1 | tail -2 | tee -a $O/summary.txt
  timeout 600 env T3D_LIBRARY=$lib python scripts/bench_sprites.py 2>&1 | grep -A4 "^  text" | grep -E "device per text|resident:" | tee -a $O/summary.txt
  timeout 600 env T3D_LIBRARY=$lib python scripts/bench_sprites.py 8388608 256 256 2>&1 | grep -A4 "^  text" | grep -E "device per text|resident:" | tee -a $O/summary.txt
done
