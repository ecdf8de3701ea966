#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/m_summary.txt; : > $S
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q -k "attention" > gpurun_out/m_kernels.log 2>&1; echo "kernels rc=$?" >> $S
timeout 300 python tools/bench_kernels.py --images 64 --only attn --tag _m > gpurun_out/m_bench_kernels.log 2>&1; echo "bench_kernels rc=$?" >> $S
timeout 300 python tools/bench_kernels.py --images 128 --S 6 --only attn --tag _m6 > gpurun_out/m_bench_kernels6.log 2>&1; echo "bench_kernels6 rc=$?" >> $S
timeout 900 python -m pytest tests/test_parity_gpu.py -q -k "wo4 or wo2_d12 or full_bench or config4" > gpurun_out/m_parity.log 2>&1; echo "parity rc=$?" >> $S
timeout 900 python bench.py --no-cpu-baseline --no-library-bar > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?" >> $S
cat $S; tail -4 gpurun_out/m_kernels.log; tail -3 gpurun_out/m_parity.log; grep attention gpurun_out/m_bench_kernels.log gpurun_out/m_bench_kernels6.log | cut -c1-200
python - <<'PY'
import json
d=json.load(open('gpurun_out/m_bench.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['roofline_hbm']['kernels'].items(): print(k, round(v['frac'],3), round(v['ms_per_step'],2))
PY
