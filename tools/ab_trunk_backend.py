"""Whole 4-scale forward (batch 256) with the trunk on the own convolution kernels against the cuDNN backend:
interleaved rounds on one box (the step is power-capped: only same-box A/B comparisons mean anything)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import duoformer_tcga_b200 as duo

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
cfg = dict(depth=12, embed_dim=768, num_heads=12, num_classes=10, num_layers=4, proj_dim=768)
m = duo.build_model_no_extra_params(pretrained=False, **cfg).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
res = {"own": [], "cudnn": []}
with torch.no_grad():
    for name in res:
        m._trunk_runner.backend = name
        for _ in range(3):
            m(x)
    torch.cuda.synchronize()
    for rnd in range(4):
        for name in res:
            m._trunk_runner.backend = name
            m(x)  # re-pack for the backend (untimed)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                m(x)
            e1.record(); torch.cuda.synchronize()
            res[name].append(e0.elapsed_time(e1) / 4)
print(json.dumps({"batch": B, "ms_per_forward": {k: [round(t, 2) for t in v] for k, v in res.items()},
                  "median": {k: round(sorted(v)[len(v) // 2], 2) for k, v in res.items()}}))
