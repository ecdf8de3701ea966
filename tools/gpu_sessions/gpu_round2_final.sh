#!/bin/bash
# final validation of the round: what the driver runs (pytest -m gpu, smoke, bench both arms) + launch list of the timed region
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/z_summary.txt; : > $S
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?" >> $S
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?" >> $S
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/z_bench_reference.json 2> gpurun_out/z_bench_reference.err; echo "bench_ref rc=$?" >> $S
B="bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar"
timeout 900 ncu --nvtx --nvtx-include "duo.bench_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_final.csv python $B > gpurun_out/z_ncu_launches.log 2>&1; echo "ncu_launches rc=$?" >> $S
cat $S; tail -4 gpurun_out/z_pytest_gpu.log; tail -1 gpurun_out/z_smoke.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/z_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], d['clocks'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'])
for k,v in d['roofline']['by_shape_NxK_epi'].items():
    if v['launches']>6: print(k.ljust(24), round(v['tflops'],1), round(v['ms_per_launch'],3))
for k,v in d['roofline_hbm']['kernels'].items(): print(k, round(v['frac'],3), round(v['ms_per_step'],2))
print(d.get('library_bar')); print(d.get('cpu_baseline'))
r=json.load(open('gpurun_out/z_bench_reference.json')); print('reference arm', r['value'], r['cpu_baseline']['sample'])
PY
