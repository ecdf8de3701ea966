#!/bin/bash
# ncu --set full of the first convolution launches of the own trunk (stem, layer1 block 0: 1x1, 3x3, downsample, conv3 + residual)
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 300 python tools/trunk_once.py 256 > gpurun_out/w_once.log 2>&1; echo "once rc=$?"; tail -2 gpurun_out/w_once.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05 -c 9 -o gpurun_out/w_conv_full -f python tools/trunk_once.py 256 > gpurun_out/w_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/w_ncu.log
ls -la gpurun_out/w_conv_full.ncu-rep
timeout -s KILL 600 python -m pytest tests/test_fp16_range_gpu.py -q > gpurun_out/w_range.log 2>&1; echo "range rc=$?"; tail -2 gpurun_out/w_range.log
