#!/bin/bash
# scale attention S = 86 restructured for three CTAs per SM (single Q|K|V buffer, O aliases S in TMEM): tests, kernel time, step
cd /root/repo
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py tests/test_bounds_gpu.py -q -k "attention" --timeout 120 > gpurun_out/y_attn_tests.log 2>&1; echo "attn tests rc=$?"; tail -3 gpurun_out/y_attn_tests.log
timeout -s KILL 300 python tools/bench_kernels.py --images 64 --only attn --tag _y 2>gpurun_out/y_bk.err | grep -E "attention" | cut -c1-160
timeout -s KILL 300 python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from duoformer_tcga_b200 import ops
# occupancy / time at batch 256 (one scale block's worth of attention)
M = 256 * 49 * 86
qkv = torch.randn(M, 2304, device='cuda').to(torch.bfloat16)
ao = torch.empty(M, 768, dtype=torch.bfloat16, device='cuda')
for _ in range(3): ops.group_attention(qkv, ao, 86, 12, 0.125)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.group_attention(qkv, ao, 86, 12, 0.125)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"attention S86 batch 256: {ms:.3f} ms, {M*768*8/ms/1e6:.0f} GB/s")
PY
: > gpurun_out/y_fwd.log
for rnd in 1 2 3; do
  timeout -s KILL 300 python tools/fwd_time.py --tag new >> gpurun_out/y_fwd.log 2>/dev/null
  timeout -s KILL 300 python tools/fwd_time.py --lib duoformer_tcga_b200/libduoformer_sm100_fwd_oldattn.so --tag old_2cta >> gpurun_out/y_fwd.log 2>/dev/null
done
cat gpurun_out/y_fwd.log | cut -c1-140
timeout -s KILL 1500 python -m pytest tests/test_parity_gpu.py -q -x -k "wo4 or config4 or full_bench" > gpurun_out/y_parity.log 2>&1; echo "parity rc=$?"; tail -2 gpurun_out/y_parity.log
