#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/l_summary.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q -k "attention" > gpurun_out/l_kernels.log 2>&1; echo "kernels rc=$?" >> $S
timeout 300 python tools/bench_kernels.py --images 128 --S 6 --only attn --tag _l > gpurun_out/l_bench_kernels.log 2>&1; echo "bench_kernels rc=$?" >> $S
timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "wo2 or mm2 or config3" > gpurun_out/l_parity.log 2>&1; echo "parity rc=$?" >> $S
timeout 600 python tools/bench_configs.py cfg3_2scale_b128_bf16 mymodel > gpurun_out/l_configs.log 2>&1; echo "configs rc=$?" >> $S
cat $S; tail -4 gpurun_out/l_kernels.log; tail -3 gpurun_out/l_parity.log; grep attention gpurun_out/l_bench_kernels.log | cut -c1-160; cut -c1-420 gpurun_out/l_configs.log
