#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
DUO_TEST_LIB=duoformer_tcga_b200/libduoformer_sm100_fwd_tanh.so timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "gemm" > gpurun_out/r_kernels_tanh.log 2>&1; echo "kernels_tanh rc=$?"
DUO_TEST_LIB=duoformer_tcga_b200/libduoformer_sm100_fwd_tanh.so timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "logits_match and bf16" > gpurun_out/r_parity_tanh.log 2>&1; echo "parity_tanh rc=$?"
tail -2 gpurun_out/r_kernels_tanh.log; tail -2 gpurun_out/r_parity_tanh.log
: > gpurun_out/r_ab_libs.log
for rep in 1 2 3 4; do for v in "" _fwd_tanh; do
  timeout 300 python tools/fwd_time.py --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag rep$rep >> gpurun_out/r_ab_libs.log 2>/dev/null
done; done
cat gpurun_out/r_ab_libs.log | cut -c1-120
for v in "" _fwd_tanh; do timeout 300 python tools/bench_kernels.py --images 64 --only fc1 --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag _r$v 2>/dev/null | grep "gemm_fc1" | cut -c1-110; done
