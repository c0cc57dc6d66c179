#!/bin/bash
# Cache-hint / look-back-width experiment: the same bench lines with other builds of the library (python scripts/build_variants.py
# first; T3D_LIBRARY selects one).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/knobs; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
run() { # name, library path ("" = the product)
  for w in "c3 20" "c5 100" "c3d 20" "c2 100"; do set -- "$1" "$2" $w
    echo -n "$1 $3: " | tee -a $O/summary.txt
    timeout 300 env T3D_LIBRARY="$2" $B --workload $3 --steps $4 --warmup 5 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/summary.txt
  done
}
run base ""
for v in ${KNOBS:-ldcs ld256 ld128 ldna ldna256 ldlu cpa256 cpa128 bulkef bulkun}; do run $v $PWD/tractor3d_b200/csrc/_build/libt3d_particles_$v.so; done
run base ""
