#!/bin/bash
# evidence for the last kernel changes: ncu --set full of the 128 x 256 convolution tiles and of the two-CTA patch attention
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05 -s 24 -c 10 -o gpurun_out/aj_conv_full -f python tools/trunk_once.py 256 > gpurun_out/aj_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:patch_attention_tc -c 1 -o gpurun_out/aj_patch_full -f python tools/patch_attention_time.py > gpurun_out/aj_ncu_patch.log 2>&1; echo "ncu patch rc=$?"
ls -la gpurun_out/aj_*.ncu-rep
