#!/bin/bash
# first light of the implicit-GEMM convolution kernel: kernel tests per group (each under its own hard timeout), trunk A/B
cd /root/repo
mkdir -p gpurun_out
for grp in "conv2d_vs_torch and 1-1" "conv2d_vs_torch" "fp16_operands or writes_nothing" "stem" "own_trunk"; do
  tag=$(echo "$grp" | tr ' ' '_')
  timeout -s KILL 420 python -m pytest tests/test_conv_gpu.py -q -k "$grp" --timeout 180 > gpurun_out/u_conv_$tag.log 2>&1
  echo "[$grp] rc=$?"; grep -E "passed|failed|error|Error|Timeout" gpurun_out/u_conv_$tag.log | tail -4
done
timeout -s KILL 300 python tools/trunk_ab.py 256 --layers > gpurun_out/u_trunk_ab.json 2> gpurun_out/u_trunk_ab.err; echo "trunk_ab rc=$?"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/u_trunk_ab.json'))
    print({k: v for k, v in d.items() if k != 'layers'})
    for r in d.get('layers', []): print(r)
except Exception as e:
    print('no trunk_ab', e); print(open('gpurun_out/u_trunk_ab.err').read()[-2000:])
PY
