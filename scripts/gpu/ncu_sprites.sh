#!/bin/bash
# ncu --set full of the 2-D sprite expansion and of the text expansion at 8 Mi items.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/ncu_sprites; mkdir -p $O
N="ncu --set full --import-source on --clock-control none"
timeout 600 $N -k regex:k_sprite_expand --launch-skip 5 -c 1 -o $O/prof_sprites python scripts/bench_sprites.py 8388608 256 256 > $O/ncu_sprites.log 2>&1
timeout 600 $N -k regex:k_text_expand --launch-skip 5 -c 1 -o $O/prof_text python scripts/bench_sprites.py 8388608 256 256 > $O/ncu_text.log 2>&1
for r in $O/prof_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null
  python scripts/ncu_top_stalls.py $r 30 > ${r%.ncu-rep}_stalls.txt 2>/dev/null
  rm -f $r
done
ls -la $O
