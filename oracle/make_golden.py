"""Generate tests/golden/*.pt from the REAL reference modules — TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):  python -m oracle.make_golden
Each fixture holds the case config, the weight/input seeds (weights are regenerated from the
seed by oracle/synth.py, not stored), the reference logits and probes (256 sampled values +
L2 norm) of intermediates captured with hooks on the reference modules.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import reference_adapter as ra  # noqa: E402
from oracle import synth  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

CASES = {
    # BASELINE.json configs[0]: main_toy 2-scale, batch 2 (wo-extra kwargs, main_toy.py:84-98)
    "wo2_d12": dict(kind="wo", depth=12, num_layers=2, batch=2, backbone="r50", scale_token="random"),
    "wo4_d2": dict(kind="wo", depth=2, num_layers=4, batch=2, backbone="r50", scale_token="random"),
    # BASELINE.json configs[1] model (4-scale, depth 12) at batch 2: the bench workload's architecture
    "wo4_d12": dict(kind="wo", depth=12, num_layers=4, batch=2, backbone="r50", scale_token="random"),
    "wo3_d2": dict(kind="wo", depth=2, num_layers=3, batch=1, backbone="r50", scale_token="random"),
    "wo2_channel_d2": dict(kind="wo", depth=2, num_layers=2, batch=2, backbone="r50", scale_token="channel"),
    "wo2_swav_d2": dict(kind="wo", depth=2, num_layers=2, batch=2, backbone="r50_Swav", scale_token="random"),
    # main_toy literally calls build_model -> MyModel (patch_size=32 so pos_embed has 50 rows)
    "mm2_d12": dict(kind="mm", depth=12, num_layers=2, batch=2),
    "mm2_d1": dict(kind="mm", depth=1, num_layers=2, batch=3),
    "mm2_d2_b1": dict(kind="mm", depth=2, num_layers=2, batch=1),  # .squeeze() -> [ncls]
}
COMMON = dict(embed_dim=768, num_heads=12, num_classes=10, proj_dim=768)


def probe(t: torch.Tensor, n: int = 256) -> dict:
    flat = t.detach().float().reshape(-1)
    g = torch.Generator().manual_seed(flat.numel())
    idx = torch.randperm(flat.numel(), generator=g)[:n].clone()
    return {"shape": tuple(t.shape), "idx": idx, "values": flat[idx].clone(), "norm": flat.norm().item(),
            "absmax": flat.abs().max().item()}


def build_reference(case: dict):
    if case["kind"] == "wo":
        mods = ra.load_reference()
        if case["backbone"] == "r50_Swav":  # no network: construct the SSL trunk without downloading
            orig = mods["wo"].resnet50FeatureExtractor
            mods["wo"].resnet50FeatureExtractor = lambda pretrained, progress, key, **kw: orig(False, progress, key, **kw)
            # D17: the reference hands backbone="r50_Swav" to Projection, which only knows "r50"/"r18"
            # (projection_head.py:13,60) and so creates no heads -> AttributeError at :142.  The SSL
            # trunk has ResNet-50 channel counts, so the obvious fix is the "r50" projection.
            proj = mods["ph"].Projection
            mods["wo"].Projection = lambda num_layers, proj_dim, backbone: proj(num_layers=num_layers, proj_dim=proj_dim, backbone="r50")
        return ra.build_wo_extra(depth=case["depth"], num_layers=case["num_layers"], backbone=case["backbone"],
                                 scale_token=case["scale_token"], **COMMON)
    return ra.build_mymodel(depth=case["depth"], patch_size=32, init_values=1e-5, num_layers=case["num_layers"],
                            model_ver="scaleformer", pretrained=False, **COMMON)


def run_case(name: str, case: dict) -> dict:
    ref = build_reference(case)
    sd = synth.synth_state_dict(ref.state_dict(), seed=0)
    ref.load_state_dict(sd)
    x = synth.synth_images(case["batch"])
    probes = {}
    vt = ref.vision_transformer
    hooks = [vt.register_forward_pre_hook(lambda m, a: probes.__setitem__("tokens_in", probe(a[0])))]
    if case["kind"] == "wo":
        for i, blk in enumerate(vt.scaleBlocks):
            hooks.append(blk.register_forward_hook(lambda m, a, o, i=i: probes.__setitem__(f"scale_block_{i}", probe(o))))
        for i, blk in enumerate(vt.blocks):
            hooks.append(blk.register_forward_hook(lambda m, a, o, i=i: probes.__setitem__(f"patch_block_{i}", probe(o))))
    with torch.no_grad():
        logits = ref(x)
    for h in hooks:
        h.remove()
    return {"name": name, "case": case, "common": COMMON, "weight_seed": 0, "input_seed": 1234,
            "keys": sorted(sd.keys()), "logits": logits.clone(), "probes": probes,
            "torch_version": torch.__version__}


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        out = run_case(name, case)
        torch.save(out, os.path.join(GOLDEN_DIR, f"{name}.pt"))
        print(name, tuple(out["logits"].shape), out["logits"].flatten()[:4].tolist(), sorted(out["probes"].keys())[:3])


if __name__ == "__main__":
    main()
