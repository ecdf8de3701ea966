#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/e_summary.txt; : > $S
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q > gpurun_out/e_kernels.log 2>&1; echo "kernels rc=$?" >> $S
for v in "" _fwd_gelu8; do
  timeout 300 python tools/bench_kernels.py --images 64 --only fwd,fc1 --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag _e$v > gpurun_out/e_bench_kernels$v.log 2>&1; echo "bench_kernels$v rc=$?" >> $S
done
timeout 900 python tools/ab_forwarding.py 256 > gpurun_out/e_ab.log 2>&1; echo "ab rc=$?" >> $S
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x > gpurun_out/e_parity.log 2>&1; echo "parity rc=$?" >> $S
cat $S; tail -4 gpurun_out/e_kernels.log; tail -4 gpurun_out/e_parity.log
for v in "" _fwd_gelu8; do echo "== lib$v"; grep -E "fc1|ln_applied|_plain|seq_proj" gpurun_out/e_bench_kernels$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['kernel'].ljust(34), d['ms'], d.get('tflops'))"; done
tail -1 gpurun_out/e_ab.log
