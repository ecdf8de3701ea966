"""Time of the fp32 FMA attention (cooperative CTA per problem) at the patch-attention size of 384 x 384 tiles (N = 145,
batch 128, fp32 qkv in, split out) and at S = 86 (fp32 mode): `qb_attention_time.py [--lib path]`."""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser(); ap.add_argument("--lib", default=""); a = ap.parse_args()
if a.lib:
    from duoformer_tcga_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(a.lib)
from duoformer_tcga_b200 import ops
out = {"lib": os.path.basename(a.lib) or "default"}
for S, groups in ((145, 128), (86, 64 * 49), (50, 256)):
    qkv = torch.randn(groups * S, 2304, device="cuda")
    ao = torch.empty(groups * S, 1536, dtype=torch.bfloat16, device="cuda")
    for _ in range(3): ops.group_attention(qkv, ao, S, 12, 0.125)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.group_attention(qkv, ao, S, 12, 0.125)
    e1.record(); torch.cuda.synchronize()
    out[f"S{S}_g{groups}_ms"] = round(e0.elapsed_time(e1) / 10, 4)
print(json.dumps(out))
