#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/p_summary.txt; : > $S
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q > gpurun_out/p_kernels.log 2>&1; echo "kernels rc=$?" >> $S
timeout 2400 python -m pytest tests/test_parity_gpu.py tests/test_fp16_range_gpu.py tests/test_library_bar_gpu.py -q > gpurun_out/p_parity.log 2>&1; echo "parity rc=$?" >> $S
timeout 300 python tools/bench_kernels.py --images 64 --only fwd,ln --tag _p > gpurun_out/p_bench_kernels.log 2>&1; echo "bench_kernels rc=$?" >> $S
timeout 600 python tools/ab_forwarding.py 256 > gpurun_out/p_ab.log 2>&1; echo "ab rc=$?" >> $S
cat $S; tail -6 gpurun_out/p_kernels.log; tail -6 gpurun_out/p_parity.log
grep -E "residual|layernorm" gpurun_out/p_bench_kernels.log | cut -c1-150; tail -1 gpurun_out/p_ab.log
