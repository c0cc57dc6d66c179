#!/bin/bash
# Compares the draw-kernel output paths on the c3 workload (run on the GPU box).
P=${1:-67108864}
for v in lsu tma; do
  T3D_DRAW_VARIANT=$v python bench.py --steps 20 --warmup 3 --particles $P --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; r=d['rooflines']; print('$v', 'update ms %.3f (%.0f GB/s, %.3f)' % (k['update']['ms_per_launch'], r['update']['achieved'], r['update']['frac']), 'draw ms %.3f (%.0f GB/s, %.3f)' % (k['draw']['ms_per_launch'], r['draw']['achieved'], r['draw']['frac']), 'value %.3g' % d['value'])"
done
