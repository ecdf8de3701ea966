#!/bin/bash
# in-step A/B of the 16-warp GELU epilogue build (only measured alone so far)
cd /root/repo
mkdir -p gpurun_out
: > gpurun_out/ad_fwd.log
for rnd in 1 2 3 4; do
  timeout -s KILL 300 python tools/fwd_time.py --tag gelu8 >> gpurun_out/ad_fwd.log 2>/dev/null
  timeout -s KILL 300 python tools/fwd_time.py --lib duoformer_tcga_b200/libduoformer_sm100_fwd_gelu16.so --tag gelu16 >> gpurun_out/ad_fwd.log 2>/dev/null
done
cut -c1-150 gpurun_out/ad_fwd.log
