"""One pass of the own-kernel trunk (batch 256 by default) — the target of ncu captures of conv_tcgen05_kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import duoformer_tcga_b200 as duo
from duoformer_tcga_b200 import token_builder as tb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(0)
m = duo.MyModel_no_extra_params(depth=1, num_layers=4, pretrained=False, embed_dim=768, num_heads=12, num_classes=10, proj_dim=768).cuda().eval()
x = torch.randn(B, 3, 224, 224, device="cuda")
r = tb.TrunkRunner()
r.act_scale_override = 1.0  # no calibration pass
with torch.no_grad():
    f = r.features(m.resnet_projector, x, "bf16", False)
torch.cuda.synchronize()
print({k: tuple(v.shape) for k, v in f.items()})
