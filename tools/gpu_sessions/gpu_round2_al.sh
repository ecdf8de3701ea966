#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests/test_parity_gpu.py -q -k "forwarding_on_off or dead_work or stagewise" > gpurun_out/al_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/al_parity.log
