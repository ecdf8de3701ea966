#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/c_summary.txt; : > $S
timeout 1800 python -m pytest tests/test_parity_gpu.py -q -k "forwarding or dead_work or scale_block_module or config3 or config4 or graph" > gpurun_out/c_parity.log 2>&1; echo "parity rc=$?" >> $S
for w in fc1_gelu_ln fc1_gelu; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_tcgen05 --launch-skip 3 --launch-count 1 -o gpurun_out/c_ncu_$w -f python tools/one_gemm.py $w 64 > gpurun_out/c_ncu_$w.log 2>&1; echo "ncu $w rc=$?" >> $S
done
cat $S; tail -5 gpurun_out/c_parity.log
