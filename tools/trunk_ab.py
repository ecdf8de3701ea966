"""Trunk pass on the own convolution kernels against the fused cuDNN calls: CUDA-event time per backend (interleaved
rounds, batch 256 by default) and, with --layers, per convolution of the own path (ops.PROFILE)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import duoformer_tcga_b200 as duo
from duoformer_tcga_b200 import ops, token_builder as tb

B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
size = 224
torch.manual_seed(0)
m = duo.MyModel_no_extra_params(depth=1, num_layers=4, pretrained=False, embed_dim=768, num_heads=12, num_classes=10, proj_dim=768).cuda().eval()
x = torch.randn(B, 3, size, size, device="cuda")
runners = {}
for name in ("own", "cudnn"):
    r = tb.TrunkRunner()
    r.backend = name
    runners[name] = r
res = {k: [] for k in runners}
with torch.no_grad():
    for name, r in runners.items():
        for _ in range(2):
            r.features(m.resnet_projector, x, "bf16", False)  # calibration / verification / warm-up
    for rep in range(5):
        for name, r in runners.items():
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                r.features(m.resnet_projector, x, "bf16", False)
            e1.record(); torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 3)
    out = {"batch": B, **{k: round(sorted(v)[len(v) // 2], 3) for k, v in res.items()}}
    if "--layers" in sys.argv:
        ops.PROFILE = []
        runners["own"].features(m.resnet_projector, x, "bf16", False)
        torch.cuda.synchronize()
        rows = []
        for e0, e1, kind, flops, nbytes, tag in ops.PROFILE:
            ms = e0.elapsed_time(e1)
            rows.append({"tag": tag, "ms": round(ms, 4), "tflops": round(flops / ms / 1e9, 1), "gbs": round(nbytes / ms / 1e6, 1)})
        ops.PROFILE = None
        out["layers"] = rows
        out["sum_conv_ms"] = round(sum(r["ms"] for r in rows), 3)
print(json.dumps(out))
