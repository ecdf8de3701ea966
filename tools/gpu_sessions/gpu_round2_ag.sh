#!/bin/bash
# patch attention with two CTAs per SM (single operand buffer): tests, kernel time against the one-CTA build, 2-scale forward
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q -k "attention" --timeout 120 2>&1 | tail -2
for l in "" "--lib duoformer_tcga_b200/libduoformer_sm100_fwd_oldpatch.so" "" "--lib duoformer_tcga_b200/libduoformer_sm100_fwd_oldpatch.so"; do
  timeout -s KILL 300 python tools/patch_attention_time.py $l 2>&1 | tail -1
done
timeout -s KILL 900 python -m pytest tests/test_parity_gpu.py -q -x -k "wo2 or wo4_d2 or config3 or mm2" 2>&1 | tail -2
timeout -s KILL 600 python tools/bench_configs.py cfg3_2scale_b128_bf16 mymodel 2>/dev/null | cut -c1-120
