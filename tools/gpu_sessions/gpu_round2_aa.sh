#!/bin/bash
# evidence for the final kernels: ncu --set full of the conv kernel (v3) and of the three-CTA scale attention; other BASELINE configs
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05 -c 12 -o gpurun_out/aa_conv_full -f python tools/trunk_once.py 256 > gpurun_out/aa_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
B="bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar"
timeout -s KILL 900 ncu --set full --import-source on --clock-control none -k regex:scale_attention_tc -s 1 -c 1 -o gpurun_out/aa_attn_full -f python $B > gpurun_out/aa_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/aa_*.ncu-rep
timeout -s KILL 1200 python tools/bench_configs.py > gpurun_out/aa_configs.log 2>&1; echo "configs rc=$?"; cut -c1-200 gpurun_out/aa_configs.log
