#!/bin/bash
# Sweeps the update-kernel variants on the c3 workload (run on the GPU box): prints ms per launch of each.
P=${1:-67108864}
for v in a4t96s11b2 a4t96s8b2 a4t96s10b2 a4t64s10b3 a4t64s8b3 v4t128s4b3 v4t128s1b4 v4t256s4b2 v2t256s4b4; do
  T3D_UPDATE_VARIANT=$v python bench.py --steps 20 --warmup 3 --particles $P --no-cpu-baseline 2>/dev/null | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['kernels']; r=d['rooflines']; print('$v', 'update ms %.3f (%.0f GB/s, %.3f)' % (k['update']['ms_per_launch'], r['update']['achieved'], r['update']['frac']), 'draw ms %.3f (%.0f GB/s)' % (k['draw']['ms_per_launch'], r['draw']['achieved']), 'value %.3g' % d['value'])"
done
