#!/bin/bash
# first GPU pass of round 2: new forwarding kernels, then everything else
cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "forward or statistics" > gpurun_out/a_kernels_fwd.log 2>&1; echo "kernels_fwd rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q > gpurun_out/a_kernels.log 2>&1; echo "kernels rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q > gpurun_out/a_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/a_summary.txt
timeout 600 python tools/bench_kernels.py --images 64 > gpurun_out/a_bench_kernels.log 2>&1; echo "bench_kernels rc=$?" | tee -a gpurun_out/a_summary.txt
cp gpurun_out/bench_kernels.json gpurun_out/a_bench_kernels.json
timeout 600 python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" | tee -a gpurun_out/a_summary.txt
tail -3 gpurun_out/a_kernels_fwd.log gpurun_out/a_kernels.log gpurun_out/a_parity.log
grep -E "fwd|seq_|after" gpurun_out/a_bench_kernels.log
cat gpurun_out/a_bench.json
