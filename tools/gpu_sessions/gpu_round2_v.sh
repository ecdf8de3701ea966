#!/bin/bash
# conv kernel v2 (TMA residual ring, 5-D stem map): kernel tests, trunk A/B with per-layer times, model parity tests
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_conv_gpu.py -q --timeout 180 > gpurun_out/v_conv.log 2>&1
echo "conv rc=$?"; grep -E "passed|failed|^E  |Timeout" gpurun_out/v_conv.log | tail -8
timeout -s KILL 300 python tools/trunk_ab.py 256 --layers > gpurun_out/v_trunk_ab.json 2> gpurun_out/v_trunk_ab.err; echo "trunk_ab rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/v_trunk_ab.json'))
    print({k: v for k, v in d.items() if k != 'layers'})
    agg = {}
    for r in d.get('layers', []):
        a = agg.setdefault(r['tag'], [0, 0.0, r['tflops'], r['gbs']]); a[0] += 1; a[1] += r['ms']
    for k, v in agg.items(): print(k.ljust(28), v[0], round(v[1], 3), 'ms total', v[2], 'TF/s', v[3], 'GB/s')
except Exception as e:
    print('no trunk_ab', e); print(open('gpurun_out/v_trunk_ab.err').read()[-2000:])
PY
timeout -s KILL 1500 python -m pytest tests/test_parity_gpu.py tests/test_fp16_range_gpu.py -q -x > gpurun_out/v_parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/v_parity.log
