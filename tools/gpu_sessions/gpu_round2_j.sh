#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
S=gpurun_out/j_summary.txt; : > $S
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q > gpurun_out/j_kernels.log 2>&1; echo "kernels rc=$?" >> $S
for v in "" _fwd_nohint; do
  timeout 300 python tools/bench_kernels.py --images 64 --only fwd,attn --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag _j$v > gpurun_out/j_bench_kernels$v.log 2>&1; echo "bench_kernels$v rc=$?" >> $S
done
for v in "" _fwd_nohint; do
  timeout 300 python tools/bench_kernels.py --images 64 --only fwd --lib duoformer_tcga_b200/libduoformer_sm100$v.so --tag _j2$v > gpurun_out/j2_bench_kernels$v.log 2>&1
done
B="bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-library-bar"
timeout 900 ncu --nvtx --nvtx-include "duo.bench_timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python $B > gpurun_out/j_ncu_launches.log 2>&1; echo "ncu_launches rc=$?" >> $S
timeout 900 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err; echo "bench rc=$?" >> $S
cat $S; tail -3 gpurun_out/j_kernels.log
for f in j_bench_kernels j_bench_kernels_fwd_nohint j2_bench_kernels j2_bench_kernels_fwd_nohint; do echo "== $f"; grep -E "residual|attention" gpurun_out/$f.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['kernel'].ljust(34), d['ms'], d.get('tflops'), d.get('gbs'))"; done
python - <<'PY'
import json
d=json.load(open('gpurun_out/j_bench.json'))
print(d['value'], d['ms_per_step'], d['clocks'], d['e2e']['value'], d['roofline']['frac'])
for k,v in d['roofline_hbm']['kernels'].items(): print(k, round(v['frac'],3), round(v['ms_per_step'],2))
PY
wc -l gpurun_out/r02_launches.csv
