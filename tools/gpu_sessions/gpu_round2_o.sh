#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/o_ab_libs.log
for rep in 1 2 3; do for v in "" _fwd_stef _fwd_gelu1buf; do
  timeout 300 python tools/fwd_time.py --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag rep$rep >> gpurun_out/o_ab_libs.log 2>/dev/null
done; done
cat gpurun_out/o_ab_libs.log
