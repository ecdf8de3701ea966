"""Diagnostic (run by hand on a GPU box): which stage breaks batch-permutation invariance bit for bit
(finding: only the cuDNN trunk, by one fp16 ulp in layer4; the own kernels are bit-exact)."""
import os, sys
_here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_here)); sys.path.insert(0, _here)
import torch
from common import build_product, load_golden
from oracle import synth
gold = load_golden("wo4_d12"); case = gold["case"]
model = build_product(case); model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0)); model = model.cuda().eval()
x = synth.synth_images(256, seed=5).cuda()
perm = torch.randperm(256, generator=torch.Generator().manual_seed(1)).cuda()
with torch.no_grad():
    f1 = model._trunk_runner.features(model.resnet_projector, x, "bf16", False)
    f2 = model._trunk_runner.features(model.resnet_projector, x[perm], "bf16", False)
    f3 = model._trunk_runner.features(model.resnet_projector, x, "bf16", False)
    for k in f1:
        print("stage", k, "perm max abs diff", (f1[k][perm].float() - f2[k].float()).abs().max().item(), "rerun diff", (f1[k].float()-f3[k].float()).abs().max().item(), "absmax", f1[k].float().abs().max().item())
    X1 = model.build_tokens(x); X2 = model.build_tokens(x[perm])
    print("tokens perm diff", (X1[perm]-X2).abs().max().item(), X1.abs().max().item())
    y1 = model.vision_transformer.forward_prepared(X1.clone()); y2 = model.vision_transformer.forward_prepared(X1[perm].clone().contiguous())
    print("transformer-only perm diff", (y1[perm]-y2).abs().max().item(), "rerun", (model.vision_transformer.forward_prepared(X1.clone())-y1).abs().max().item())
