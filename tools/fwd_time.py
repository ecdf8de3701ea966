"""Forward time of the bench model (batch 256) with a given build of the library: `fwd_time.py [--lib path] [--tag t]`.
Used ABAB-interleaved from a shell loop to compare tuning builds on one box (prints one JSON line)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default="")
ap.add_argument("--tag", default="")
ap.add_argument("--batch", type=int, default=256)
a = ap.parse_args()
if a.lib:
    from duoformer_tcga_b200 import _lib
    _lib.LIB_PATH = os.path.abspath(a.lib)
import duoformer_tcga_b200 as duo  # noqa: E402

torch.manual_seed(0)
model = duo.build_model_no_extra_params(pretrained=False, depth=12, embed_dim=768, num_heads=12, num_classes=10,
                                        num_layers=4, proj_dim=768).cuda().eval()
x = torch.randn(a.batch, 3, 224, 224, device="cuda")
ts = []
with torch.no_grad():
    for _ in range(4):
        model(x)
    torch.cuda.synchronize()
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
ts.sort()
print(json.dumps({"lib": os.path.basename(a.lib) or "default", "tag": a.tag, "median_ms": round(ts[len(ts) // 2], 2),
                  "min_ms": round(ts[0], 2), "all": [round(t, 1) for t in ts]}), flush=True)
