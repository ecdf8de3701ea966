#!/bin/bash
# convolution kernel with 128 x 256 tiles where Cout % 256 == 0 and there are two waves of them: tests, per-layer trunk times
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_conv_gpu.py -q --timeout 180 > gpurun_out/ah_conv.log 2>&1
echo "conv rc=$?"; grep -E "passed|failed|^E  |Timeout" gpurun_out/ah_conv.log | tail -6
timeout -s KILL 300 python tools/trunk_ab.py 256 --layers > gpurun_out/ah_trunk_ab.json 2> gpurun_out/ah_trunk_ab.err; echo "trunk_ab rc=$?"
timeout -s KILL 300 python tools/trunk_ab.py 128 > gpurun_out/ah_trunk_ab_b128.json 2>/dev/null; cat gpurun_out/ah_trunk_ab_b128.json
python - <<'PY'
import json
d = json.load(open('gpurun_out/ah_trunk_ab.json'))
print({k: v for k, v in d.items() if k != 'layers'})
agg = {}
for r in d.get('layers', []):
    a = agg.setdefault(r['tag'], [0, 0.0, r['tflops'], r['gbs']]); a[0] += 1; a[1] += r['ms']
for k, v in agg.items(): print(k.ljust(28), v[0], round(v[1], 3), 'ms total', v[2], 'TF/s', v[3], 'GB/s')
PY
