"""Composed patch linears on / off, interleaved on one box: the 2-scale model (batch 128: the patch stage is a fifth of the
forward) and the 4-scale bench model (batch 256)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import duoformer_tcga_b200 as duo

out = {}
for name, layers, B in (("2scale_b128", 2, 128), ("4scale_b256", 4, 256)):
    torch.manual_seed(0)
    m = duo.build_model_no_extra_params(pretrained=False, depth=12, embed_dim=768, num_heads=12, num_classes=10, num_layers=layers, proj_dim=768).cuda().eval()
    x = torch.randn(B, 3, 224, 224, device="cuda")
    vt = m.vision_transformer
    res = {True: [], False: []}
    with torch.no_grad():
        for f in res:
            vt.fuse_patch_linears = f
            for _ in range(3): m(x)
        torch.cuda.synchronize()
        for rnd in range(5):
            for f in res:
                vt.fuse_patch_linears = f
                m(x)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(8 if layers == 2 else 3): m(x)
                e1.record(); torch.cuda.synchronize()
                res[f].append(round(e0.elapsed_time(e1) / (8 if layers == 2 else 3), 3))
    out[name] = {"fused_ms": res[True], "sequential_ms": res[False],
                 "median": {"fused": sorted(res[True])[2], "sequential": sorted(res[False])[2]}}
    del m, x
    torch.cuda.empty_cache()
print(json.dumps(out))
