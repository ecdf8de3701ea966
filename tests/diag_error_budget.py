"""Diagnostic (run by hand on a GPU box): per-stage relative error of the bf16 path vs the fp32
oracle for the bench architecture (4-scale, depth 12, batch 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from common import build_product, load_golden, oracle_forward, relerr
from oracle import synth

torch.set_num_threads(os.cpu_count() or 1)
gold = load_golden("wo4_d12"); case = gold["case"]
model = build_product(case)
sd = synth.synth_state_dict(model.state_dict(), seed=0)
model.load_state_dict(sd)
x = synth.synth_images(case["batch"], seed=gold["input_seed"])
ocap = {}
with torch.no_grad():
    yo = oracle_forward(case, x, sd, capture=ocap)
model = model.cuda().eval()
for variant in sys.argv[1:] or ["default"]:
    model.set_precision("bf16")
    model._trunk_runner.trunk_dtype = "fp32" if "trunk32" in variant else ("bf16" if "trunkbf16" in variant else "fp16")
    model._trunk_runner._trunk = None
    model.vision_transformer.patch_precision = None if "patchbf16" in variant else ("mixed" if "patchmixed" in variant else "fp32")
    cap = {}
    model.vision_transformer._capture = cap
    with torch.no_grad():
        y = model(x.cuda()).float().cpu()
    print("==", variant, "logits rel err", round(relerr(y, yo), 5))
    for k, t in cap.items():
        ref = ocap[k[:-3]][:, :, 0, :] if k.endswith("_s0") else (ocap[k[:-4]][:, 0, :] if k.endswith("_cls") else ocap[k])
        print(f"   {k:18s} {relerr(t, ref):.4e}")
