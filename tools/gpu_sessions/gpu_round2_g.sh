#!/bin/bash
# 2-GPU box: NCCL logits-equality test, weak- and strong-scaling bench lines
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/g_summary.txt; : > $S
nvidia-smi -L > gpurun_out/g_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_distributed_gpu.py -q -rs > gpurun_out/g_dist.log 2>&1; echo "dist rc=$?" >> $S
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g_bench_2gpu_weak.json 2> gpurun_out/g_bench_2gpu_weak.err; echo "weak2 rc=$?" >> $S
timeout 600 $TR bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong --global-batch 2048 > gpurun_out/g_bench_2gpu_strong.json 2> gpurun_out/g_bench_2gpu_strong.err; echo "strong2 rc=$?" >> $S
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --scaling strong --global-batch 2048 --no-cpu-baseline --no-library-bar > gpurun_out/g_bench_1gpu_strong.json 2> gpurun_out/g_bench_1gpu_strong.err; echo "strong1 rc=$?" >> $S
cat $S; tail -5 gpurun_out/g_dist.log
for f in g_bench_2gpu_weak g_bench_2gpu_strong g_bench_1gpu_strong; do python - <<PY
import json
try:
    d=json.load(open('gpurun_out/$f.json')); print('$f', d['value'], d['ms_per_step'], d['n_gpus'], d['scaling'], d['config']['global_batch'], d['e2e']['value'], d['clocks'])
except Exception as e: print('$f', 'ERR', e)
PY
done
