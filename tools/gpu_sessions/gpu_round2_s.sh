#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q > gpurun_out/s_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1800 python -m pytest tests/test_parity_gpu.py -q > gpurun_out/s_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/s_kernels.log; tail -3 gpurun_out/s_parity.log
: > gpurun_out/s_ab_libs.log
for rep in 1 2 3 4; do for v in "" _fwd_sigmoid; do
  timeout 300 python tools/fwd_time.py --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag rep$rep >> gpurun_out/s_ab_libs.log 2>/dev/null
done; done
cat gpurun_out/s_ab_libs.log | cut -c1-120
