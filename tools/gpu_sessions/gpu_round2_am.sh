#!/bin/bash
# scale attention: probabilities by packed ex2.approx.f16x2, P as the fp16 A operand against the bf16 V^T (mixed-format UMMA)
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q -k "attention" --timeout 120 > gpurun_out/am_tests.log 2>&1; echo "attn tests rc=$?"; grep -E "passed|failed|^E  " gpurun_out/am_tests.log | tail -6
timeout -s KILL 300 python tools/bench_kernels.py --images 64 --only attn --tag _am 2>/dev/null | grep -E "algo3" | cut -c1-160
timeout -s KILL 300 python tools/bench_kernels.py --images 64 --only attn --tag _am_old --lib duoformer_tcga_b200/libduoformer_sm100_fwd_oldattn.so 2>/dev/null | grep -E "algo3" | cut -c1-160
: > gpurun_out/am_fwd.log
for rnd in 1 2 3; do
  timeout -s KILL 300 python tools/fwd_time.py --tag f16exp >> gpurun_out/am_fwd.log 2>/dev/null
  timeout -s KILL 300 python tools/fwd_time.py --lib duoformer_tcga_b200/libduoformer_sm100_fwd_oldattn.so --tag f32exp >> gpurun_out/am_fwd.log 2>/dev/null
done
cut -c1-150 gpurun_out/am_fwd.log
