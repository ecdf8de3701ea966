"""CUDA-event timing of the stages of one forward (trunk, token builder, scale stage, patch stage + head)
at the bench configuration (4-scale, depth 12, batch 256)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import duoformer_tcga_b200 as duo
from duoformer_tcga_b200 import engine, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
if os.environ.get("CUDNN_BENCHMARK") == "1":
    torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
m = duo.MyModel_no_extra_params(depth=12, num_layers=4, pretrained=False, embed_dim=768, num_heads=12, num_classes=10, proj_dim=768).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
vt = m.vision_transformer

def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e

with torch.no_grad():
    for _ in range(3):
        m(x)
    res = {}
    for rep in range(3):
        t0 = ev()
        feats = m._trunk_runner.features(m.resnet_projector, x, "bf16", False)
        t1 = ev()
        X = m._token_builder.build(feats, m.projection, 4, m.channel_token.detach().reshape(-1).float(), vt.pos_scale_table(), "bf16")
        t2 = ev()
        engine.scale_stage(X, [b.pack("bf16") for b in vt.scaleBlocks], 12, 0.125, 1e-6, "bf16", vt.workspace(X.device),
                           live_only_last=True)
        t3 = ev()
        # patch stage + head, as in MultiscaleFormer.forward_prepared
        saved = vt.scaleBlocks
        vt.scaleBlocks = torch.nn.Sequential()
        y = vt.forward_prepared(X)
        vt.scaleBlocks = saved
        t4 = ev()
        torch.cuda.synchronize()
        for k, a, b in (("trunk", t0, t1), ("token_builder", t1, t2), ("scale_stage", t2, t3), ("patch_stage_head", t3, t4), ("total", t0, t4)):
            res.setdefault(k, []).append(a.elapsed_time(b))
    out = {k: round(sorted(v)[1], 3) for k, v in res.items()}
    out["batch"] = B
    print(json.dumps(out))
