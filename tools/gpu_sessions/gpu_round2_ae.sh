#!/bin/bash
# fp32 FMA attention (N = 145 patch attention of config 4; S = 86 of the fp32 mode): warps per CTA 4 / 8 / 16
cd /root/repo
mkdir -p gpurun_out
for l in "" "--lib duoformer_tcga_b200/libduoformer_sm100_fwd_qb4.so" "--lib duoformer_tcga_b200/libduoformer_sm100_fwd_qb16.so"; do
  timeout -s KILL 300 python tools/qb_attention_time.py $l 2>&1 | tail -1
done
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention" --timeout 120 2>&1 | tail -2
