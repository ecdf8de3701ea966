#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/k_summary.txt; : > $S
B="bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar"
timeout 900 ncu --nvtx --nvtx-include "duo.bench_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python $B > gpurun_out/k_ncu_launches.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 900 python tools/bench_configs.py cfg3_2scale_b128_bf16 cfg4 mymodel > gpurun_out/k_configs.log 2>&1; echo "configs rc=$?" >> $S
timeout 300 python tools/bench_kernels.py --images 64 --only fwd,attn --tag _k > gpurun_out/k_bench_kernels.log 2>&1; echo "bench_kernels rc=$?" >> $S
cat $S; wc -l gpurun_out/r02_launches.csv; cat gpurun_out/k_configs.log | cut -c1-900
grep -E "residual|attention" gpurun_out/k_bench_kernels.log | cut -c1-140
