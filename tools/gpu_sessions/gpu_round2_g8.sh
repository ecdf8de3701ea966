#!/bin/bash
# 8-GPU box, final kernels: weak-scaling bench line (256 images per GPU)
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g8_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout -s KILL 600 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/g8_bench_8gpu_weak.json 2> gpurun_out/g8_bench_8gpu_weak.err; echo "weak8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/g8_bench_8gpu_weak.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['scaling'], d['e2e']['value'], d['clocks'], d['gpu_launches'])
PY
