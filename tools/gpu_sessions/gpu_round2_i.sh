#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/i_summary.txt; : > $S
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/i_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/i_smoke.log 2>&1; echo "smoke rc=$?" >> $S
B="bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar"
timeout 600 python $B > gpurun_out/i_bench_short.json 2> gpurun_out/i_bench_short.err; echo "bench_short rc=$?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python $B > gpurun_out/i_ncu_launches.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 900 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05_pair_kernel -s 8 -c 4 -o gpurun_out/r02_ncu_scale_block_gemms -f python $B > gpurun_out/i_ncu_gemms.log 2>&1; echo "ncu_gemms rc=$?" >> $S
timeout 900 ncu --set full --import-source on --clock-control none -k regex:scale_attention_tc -s 1 -c 1 -o gpurun_out/r02_ncu_scale_attention -f python $B > gpurun_out/i_ncu_attn.log 2>&1; echo "ncu_attn rc=$?" >> $S
cat $S; tail -5 gpurun_out/i_pytest_gpu.log; tail -2 gpurun_out/i_smoke.log
