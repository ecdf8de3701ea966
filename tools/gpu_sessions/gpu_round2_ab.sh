#!/bin/bash
# short-K forwarding epilogue (proj) on 8 warps in two tile-alternating groups: tests, kernel time, ABAB whole forward against the 4-warp build
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q -k "forward or fwd or stat or residual or bound" --timeout 180 > gpurun_out/ab_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/ab_tests.log
timeout -s KILL 300 python tools/bench_kernels.py --images 64 --only fwd --tag _ab8 2>/dev/null | grep -E "residual|seq_" | cut -c1-150
timeout -s KILL 300 python tools/bench_kernels.py --images 64 --only fwd --tag _ab4 --lib duoformer_tcga_b200/libduoformer_sm100_fwd_proj4.so 2>/dev/null | grep -E "residual|seq_" | cut -c1-150
: > gpurun_out/ab_fwd.log
for rnd in 1 2 3; do
  timeout -s KILL 300 python tools/fwd_time.py --tag proj8 >> gpurun_out/ab_fwd.log 2>/dev/null
  timeout -s KILL 300 python tools/fwd_time.py --lib duoformer_tcga_b200/libduoformer_sm100_fwd_proj4.so --tag proj4 >> gpurun_out/ab_fwd.log 2>/dev/null
done
cut -c1-150 gpurun_out/ab_fwd.log
timeout -s KILL 1200 python -m pytest tests/test_parity_gpu.py -q -x -k "wo4 or forwarding or full_bench" > gpurun_out/ab_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/ab_parity.log
