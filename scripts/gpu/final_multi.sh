#!/bin/bash
# Round-2 multi-GPU measurements (gpurun --gpus 8): C4 over 2 / 4 / 8 GPUs as the driver launches it, also.c5 with 1 Mi
# spawned per frame PER GPU, and one process driving one context per GPU through the C ABI.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/multi; mkdir -p $O
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv > $O/smi.txt 2>&1
for n in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 30 --warmup 5 > $O/bench_n$n.json 2> $O/bench_n$n.err; echo "n$n rc=$?" | tee -a $O/rc.txt
  scripts/bench_summary.py < $O/bench_n$n.json | tee -a $O/summary.txt
done
# the N = 1 point of the same workloads, on this box
timeout 600 python bench.py --workload c4 --steps 12 --no-cpu-baseline --no-also --no-vertices-to-host > $O/bench_c4_n1.json 2> $O/bench_c4_n1.err; echo "c4n1 rc=$?" | tee -a $O/rc.txt
scripts/bench_summary.py < $O/bench_c4_n1.json | tee -a $O/summary.txt
timeout 600 python -m pytest tests/test_gpu_multi_context.py tests/test_sharding_gloo.py -q > $O/multi_ctx.log 2>&1; echo "multi_ctx rc=$?" | tee -a $O/rc.txt
tail -4 $O/multi_ctx.log
ls -la $O
