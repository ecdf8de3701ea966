"""CUDA-event timing of the pieces of the trunk pass (batch 256): input conversion, stem, max-pool, layer1..4."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import duoformer_tcga_b200 as duo
from duoformer_tcga_b200 import token_builder as tb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = duo.MyModel_no_extra_params(depth=1, num_layers=4, pretrained=False, embed_dim=768, num_heads=12, num_classes=10, proj_dim=768).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
with torch.no_grad():
    m(x)
    t = m._trunk_runner._packed_trunk(m.resnet_projector, "bf16", x, False)
    ch = dict(t.named_children())
    by_scale = "conv1" in ch
    stem, pool = (t.conv1, t.maxpool) if by_scale else (ch["0"], ch["3"])
    layers = [t.layer1, t.layer2, t.layer3, t.layer4] if by_scale else [ch["4"], ch["5"], ch["6"], ch["7"]]

    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    res = {}
    for rep in range(4):
        e = [ev()]
        xh = x.to(dtype=torch.float16).contiguous(memory_format=torch.channels_last); e.append(ev())
        y = torch.cudnn_convolution_relu(xh, stem.weight, stem.bias, *tb._conv_args(stem)); e.append(ev())
        y = tb._stem_pool(pool, y); e.append(ev())
        for layer in layers:
            for blk in layer:
                y = tb._fused_bottleneck(blk, y)
            e.append(ev())
        torch.cuda.synchronize()
        names = ["input_to_fp16_nhwc", "stem", "maxpool", "layer1", "layer2", "layer3", "layer4"]
        for k, a, b in zip(names, e[:-1], e[1:]):
            res.setdefault(k, []).append(a.elapsed_time(b))
    out = {k: round(sorted(v)[1], 3) for k, v in res.items()}
    out["total"] = round(sum(out.values()), 3)
    print(json.dumps(out))
