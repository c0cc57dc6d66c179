#!/bin/bash
# Round-2 GPU session 13: pipelined batch frame (test + e2e against the blocking form), fused tile shapes under the plain-load default.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/s13; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batch_of_mixed" > $O/new.log 2>&1; echo "new rc=$?" | tee -a $O/rc.txt; tail -15 $O/new.log | cut -c1-200
B="python bench.py --no-cpu-baseline --no-also --no-vertices-to-host"
e2e() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('$1', 'value=%.4g e2e=%.4g (%.3f of value) blocking=%s ms/step=%.3f' % (d['value'], e['value'], e['value']/d['value'], ('%.4g (%.3f)' % (e['blocking_counts'], e['blocking_counts']/d['value'])) if e.get('blocking_counts') else None, d['ms_per_step']))
"; }
timeout 300 env T3D_BENCH_EMULATE_WORLD=8 $B --workload c4 --steps 30 2>/dev/null | e2e c4w8 | tee -a $O/e2e.txt
timeout 300 env T3D_BENCH_EMULATE_WORLD=8 $B --workload c4 --steps 30 2>/dev/null | e2e c4w8 | tee -a $O/e2e.txt
timeout 300 $B --workload c2 --steps 100 2>/dev/null | e2e c2 | tee -a $O/e2e.txt
timeout 300 $B --workload c5 --steps 100 2>/dev/null | e2e c5 | tee -a $O/e2e.txt
for v in f2t192s12b2 f2t192s6b2 f2t128s12b3 f2t128s8b3 f2t128s6b3; do
  for w in "c2 100" "c3 20" "c5 100"; do set -- $w
    echo -n "$v $1: " | tee -a $O/shapes.txt
    timeout 300 env T3D_FUSED_VARIANT=$v $B --workload $1 --steps $2 --warmup 5 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/shapes.txt
  done
done
echo -n "default c2: " | tee -a $O/shapes.txt
timeout 300 $B --workload c2 --steps 100 --warmup 5 2>/dev/null | scripts/bench_summary.py | head -1 | sed -e 's/ e2e=.*//' | tee -a $O/shapes.txt
