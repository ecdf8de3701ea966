#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_conv_gpu.py -q --timeout 180 -k "non_square" > gpurun_out/ai_conv.log 2>&1
echo "conv rc=$?"; grep -E "passed|failed|^E  |Timeout|Error" gpurun_out/ai_conv.log | tail -8
