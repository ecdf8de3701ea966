"""Run under compute-sanitizer (memcheck / racecheck / synccheck): one tiny forward of every model family and
precision, which between them launch every kernel of libduoformer_sm100.so (tcgen05 GEMMs with all epilogues incl.
statistics forwarding, the three scale-attention kernels, split-precision patch attention, LayerNorm, token scatter,
im2col / pool, head).  Prints the launch count; the sanitizer's own summary is the result."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import duoformer_tcga_b200 as duo  # noqa: E402
from duoformer_tcga_b200 import ops  # noqa: E402

COMMON = dict(embed_dim=768, num_heads=12, num_classes=10, proj_dim=768)
torch.manual_seed(0)
x = torch.randn(1, 3, 224, 224, device="cuda")
ops.launch_count_reset()
cases = [("wo4", lambda: duo.MyModel_no_extra_params(depth=2, num_layers=4, pretrained=False, **COMMON)),
         ("wo3", lambda: duo.MyModel_no_extra_params(depth=2, num_layers=3, pretrained=False, **COMMON)),
         ("wo2_channel", lambda: duo.MyModel_no_extra_params(depth=2, num_layers=2, scale_token="channel", pretrained=False, **COMMON)),
         ("mm2", lambda: duo.MyModel(depth=2, patch_size=32, init_values=1e-5, num_layers=2, model_ver="scaleformer",
                                     pretrained=False, **COMMON))]
only = sys.argv[1].split(",") if len(sys.argv) > 1 else None
for name, make in cases:
    if only and name not in only:
        continue
    model = make().cuda().eval()
    for prec in ("bf16", "fp32"):
        model.set_precision(prec)
        with torch.no_grad():
            y = model(x).float()
        torch.cuda.synchronize()
        print(name, prec, "finite" if bool(torch.isfinite(y).all()) else "NON-FINITE", flush=True)
    del model
print("launches:", ops.launch_count())
