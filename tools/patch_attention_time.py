"""Time of the split-precision tcgen05 patch attention (N = 50) at batch 128 / 256: `patch_attention_time.py [--lib path]`."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser(); ap.add_argument("--lib", default=""); a = ap.parse_args()
if a.lib:
    from duoformer_tcga_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(a.lib)
from duoformer_tcga_b200 import ops
out = {"lib": os.path.basename(a.lib) or "default"}
for B in (128, 256, 1024):
    rows = B * 50
    qkv = torch.randn(rows, 6 * 768, device="cuda").to(torch.bfloat16)
    ao = torch.empty(rows, 2 * 768, dtype=torch.bfloat16, device="cuda")
    for _ in range(3): ops.group_attention(qkv, ao, 50, 12, 0.125, split_in=True)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.group_attention(qkv, ao, 50, 12, 0.125, split_in=True)
    e1.record(); torch.cuda.synchronize()
    out[f"b{B}_ms"] = round(e0.elapsed_time(e1) / 20, 4)
print(json.dumps(out))
