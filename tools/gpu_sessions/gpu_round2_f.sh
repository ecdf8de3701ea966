#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/f_summary.txt; : > $S
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q > gpurun_out/f_kernels.log 2>&1; echo "kernels rc=$?" >> $S
for v in "" _fwd_gelu16; do
  timeout 300 python tools/bench_kernels.py --images 64 --only fwd,fc1 --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag _f$v > gpurun_out/f_bench_kernels$v.log 2>&1; echo "bench_kernels$v rc=$?" >> $S
done
timeout 900 python tools/ab_forwarding.py 256 > gpurun_out/f_ab.log 2>&1; echo "ab rc=$?" >> $S
timeout 1800 python -m pytest tests/test_parity_gpu.py tests/test_fp16_range_gpu.py tests/test_library_bar_gpu.py -q > gpurun_out/f_parity.log 2>&1; echo "parity rc=$?" >> $S
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?" >> $S
cat $S; tail -4 gpurun_out/f_kernels.log; tail -4 gpurun_out/f_parity.log
for v in "" _fwd_gelu16; do echo "== lib$v"; grep -E "fc1|ln_applied|_plain|seq_|residual" gpurun_out/f_bench_kernels$v.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['kernel'].ljust(34), d['ms'], d.get('tflops'))"; done
tail -1 gpurun_out/f_ab.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/f_bench.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['roofline']['by_shape_NxK_epi'].items():
    if v['launches']>6: print(k.ljust(24), round(v['tflops'],1), round(v['ms_per_launch'],3))
PY
