#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?"
timeout 1800 python -m pytest tests/test_parity_gpu.py -q > gpurun_out/t_parity.log 2>&1; echo "parity rc=$?"
tail -3 gpurun_out/t_kernels.log; tail -3 gpurun_out/t_parity.log
timeout 300 python tools/bench_kernels.py --images 64 --only fwd --tag _t 2>/dev/null | grep -E "residual|seq_" | cut -c1-120
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/t_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'warmup', d['warmup'], d['clocks'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k,v in d['roofline']['by_shape_NxK_epi'].items():
    if v['launches']>6: print(k.ljust(24), round(v['tflops'],1), round(v['ms_per_launch'],3))
PY
