#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/b_summary.txt; : > $S
timeout 900 python -m pytest tests/test_kernels_gpu.py -q > gpurun_out/b_kernels.log 2>&1; echo "kernels rc=$?" >> $S
timeout 2400 python -m pytest tests/test_parity_gpu.py tests/test_fp16_range_gpu.py tests/test_library_bar_gpu.py tests/test_distributed_gpu.py -q > gpurun_out/b_parity.log 2>&1; echo "parity rc=$?" >> $S
for v in "" _fwd_deep _fwd_longk; do
  timeout 300 python tools/bench_kernels.py --images 64 --only fwd --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag _b$v > gpurun_out/b_bench_kernels$v.log 2>&1; echo "bench_kernels$v rc=$?" >> $S
done
timeout 900 python bench.py > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" >> $S
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_models.py > gpurun_out/b_memcheck.log 2>&1; echo "memcheck rc=$?" >> $S
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_models.py wo4 > gpurun_out/b_racecheck.log 2>&1; echo "racecheck rc=$?" >> $S
cat $S
tail -5 gpurun_out/b_kernels.log; tail -15 gpurun_out/b_parity.log
for v in "" _fwd_deep _fwd_longk; do echo "== lib$v"; grep -E "residual_fwd|ln_applied|_plain|seq_" gpurun_out/b_bench_kernels$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['kernel'].ljust(34), d['ms'], d.get('tflops'))"; done
cat gpurun_out/b_bench.json | cut -c1-3000
tail -5 gpurun_out/b_memcheck.log; tail -5 gpurun_out/b_racecheck.log
