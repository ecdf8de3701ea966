"""Same-box A/B of the LayerNorm statistics forwarding: forward time of the bench model (batch 256) with both links
forwarded, only norm1 (fc2 -> QKV), only norm2 (proj -> fc1), and none (LayerNorm launches), interleaved."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import duoformer_tcga_b200 as duo  # noqa: E402
from duoformer_tcga_b200 import engine  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
model = duo.build_model_no_extra_params(pretrained=False, depth=12, embed_dim=768, num_heads=12, num_classes=10,
                                        num_layers=4, proj_dim=768).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
settings = {"both": ("norm1", "norm2"), "norm1_only": ("norm1",), "norm2_only": ("norm2",), "none": ()}
res = {k: [] for k in settings}
with torch.no_grad():
    for _ in range(3):
        model(x)
    for rep in range(4):
        for name, links in settings.items():
            engine.FORWARD_LINKS = links
            model(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                model(x)
            e1.record()
            torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 3)
out = {k: {"ms": sorted(v)[len(v) // 2], "all": [round(t, 2) for t in v]} for k, v in res.items()}
print(json.dumps(out))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/ab_forwarding.json", "w"), indent=1)
