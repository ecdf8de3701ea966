#!/bin/bash
# 2-GPU box with the own trunk / three-CTA attention: NCCL logits-equality tests, weak-scaling bench line, reference arm under torchrun
cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g2_gpus.txt 2>&1
timeout -s KILL 900 python -m pytest tests/test_distributed_gpu.py -q -rs > gpurun_out/g2_dist.log 2>&1; echo "dist rc=$?"; tail -3 gpurun_out/g2_dist.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout -s KILL 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g2_bench_2gpu_weak.json 2> gpurun_out/g2_bench_2gpu_weak.err; echo "weak2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/g2_bench_2gpu_weak.json')); print(d['value'], d['ms_per_step'], d['n_gpus'], d['scaling'], d['e2e']['value'], d['clocks'], d['gpu_launches'])
PY
