#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/d_summary.txt; : > $S
timeout 600 python -m pytest tests/test_bounds_gpu.py -q > gpurun_out/d_bounds.log 2>&1; echo "bounds rc=$?" >> $S
timeout 900 python tools/ab_forwarding.py 256 > gpurun_out/d_ab.log 2>&1; echo "ab rc=$?" >> $S
cat $S; tail -15 gpurun_out/d_bounds.log; tail -2 gpurun_out/d_ab.log
