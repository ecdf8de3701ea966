#!/bin/bash
# whole GPU suite + bench with the trunk on the own convolution kernel; A/B of the whole forward against the cuDNN backend
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 2400 python -m pytest tests -q -m gpu > gpurun_out/x_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?"; tail -4 gpurun_out/x_pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/x_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/x_smoke.log
timeout -s KILL 600 python tools/ab_trunk_backend.py > gpurun_out/x_ab_trunk.json 2> gpurun_out/x_ab_trunk.err; echo "ab rc=$?"; cat gpurun_out/x_ab_trunk.json
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/x_bench.json 2> gpurun_out/x_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/x_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['by_shape_NxK_epi'].items():
    if v['launches']>6: print(k.ljust(24), round(v['tflops'],1), round(v['ms_per_launch'],3))
for k,v in d['roofline_hbm']['kernels'].items(): print(k, round(v['frac'],3), round(v['ms_per_step'],2))
print(d.get('trunk_convs')); print(d.get('library_bar')); print(d.get('cpu_baseline'))
PY
