"""GPU check: fused cuDNN trunk path vs the plain module path (agreement + time per batch of 256)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torchvision
from duoformer_tcga_b200.token_builder import TrunkRunner

trunk = torch.nn.Sequential(*list(torchvision.models.resnet50().children())[:-2]).cuda().eval()
x = torch.randn(256, 3, 224, 224, device="cuda")
for fused in (True, False):
    r = TrunkRunner()
    if not fused:
        r.fused_ok = False
    for _ in range(3):
        f = r.features(trunk, x, "bf16", False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f = r.features(trunk, x, "bf16", False)
    e1.record(); torch.cuda.synchronize()
    print("fused" if fused else "plain", "fused_ok=", r.fused_ok, "ms/batch256 =", e0.elapsed_time(e1) / 5,
          {k: (tuple(v.shape), v.is_contiguous(memory_format=torch.channels_last)) for k, v in f.items()})
