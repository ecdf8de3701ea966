#!/bin/bash
# last scale block: q projection on the live (s = 0) rows only — parity suite + bench
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 1800 python -m pytest tests/test_parity_gpu.py -q > gpurun_out/ak_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/ak_parity.log
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-library-bar > gpurun_out/ak_bench.json 2> gpurun_out/ak_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/ak_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['by_shape_NxK_epi'].items(): print(k.ljust(24), round(v['tflops'],1), round(v['ms_per_launch'],3), v['launches'])
PY
