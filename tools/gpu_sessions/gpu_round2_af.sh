#!/bin/bash
# composed patch linears again: A/B on the 2-scale model (where the patch stage matters) and the bench model; parity
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python tools/ab_patch_fusion.py > gpurun_out/af_ab.json 2> gpurun_out/af_ab.err; echo "ab rc=$?"; cat gpurun_out/af_ab.json
timeout -s KILL 1800 python -m pytest tests/test_parity_gpu.py -q > gpurun_out/af_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/af_parity.log
