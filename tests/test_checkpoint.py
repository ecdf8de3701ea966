"""CPU (-m "not gpu"): checkpoint ingestion (duoformer_tcga_b200/checkpoint.py).

The reference's own checkpoint format is a pickle of the whole module (main_toy.py:139-149).  The
round trip through the REAL reference classes only runs where /root/reference exists (build
container); the container-independent cases cover the key normalisation logic."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

from common import COMMON, ROOT
import duoformer_tcga_b200 as duo
from duoformer_tcga_b200 import checkpoint
from oracle import synth


def _small(backbone="r50", **kw):
    return duo.MyModel_no_extra_params(depth=1, num_layers=2, backbone=backbone, pretrained=False, **COMMON, **kw).eval()


def test_plain_and_wrapped_state_dicts_and_dataparallel_prefix():
    src = _small()
    sd = synth.synth_state_dict(src.state_dict(), seed=5)
    for container in (sd, {"model": sd, "epoch": 3}, {"state_dict": {"module." + k: v for k, v in sd.items()}}):
        dst = _small()
        missing, unexpected = checkpoint.load_checkpoint(dst, container)
        assert not missing and not unexpected
        assert all(torch.equal(dst.state_dict()[k], sd[k]) for k in sd)


def test_trunk_naming_is_converted_both_ways():
    idx_model, named_model = _small("r50"), _small("r50_Swav")
    sd_idx = synth.synth_state_dict(idx_model.state_dict(), seed=6)
    checkpoint.load_checkpoint(named_model, sd_idx)  # index-based keys into the name-based trunk
    assert torch.equal(named_model.state_dict()["resnet_projector.layer4.2.bn3.running_var"],
                       sd_idx["resnet_projector.7.2.bn3.running_var"])
    sd_named = named_model.state_dict()
    back = _small("r50")
    checkpoint.load_checkpoint(back, sd_named)
    assert torch.equal(back.state_dict()["resnet_projector.0.weight"], sd_idx["resnet_projector.0.weight"])


def test_dead_keys_are_tolerated_and_real_mismatches_are_not():
    src = _small()
    sd = dict(src.state_dict())
    del sd["vision_transformer.fc_norm.weight"], sd["vision_transformer.fc_norm.bias"]  # dead in the forward
    sd["vision_transformer.patch_embed.proj.weight"] = torch.zeros(768, 3, 32, 32)     # timm leftovers
    sd["vision_transformer.blocks.0.attn.q_norm.weight"] = torch.zeros(64)
    missing, unexpected = checkpoint.load_checkpoint(_small(), sd)
    assert not missing and not unexpected
    del sd["vision_transformer.head.weight"]
    with pytest.raises(RuntimeError, match="missing"):
        checkpoint.load_checkpoint(_small(), sd)
    bad = dict(src.state_dict())
    bad["vision_transformer.head.weight"] = torch.zeros(3, 768)
    with pytest.raises(RuntimeError, match="shape mismatch"):
        checkpoint.load_checkpoint(_small(), bad)


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="needs the reference (build container only)")
def test_whole_module_pickle_of_the_real_reference_loads_without_reference_or_timm(tmp_path):
    ckpt = str(tmp_path / "duoformer_epoch_7.pt")
    save = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {ROOT!r})
        from oracle import reference_adapter as ra, synth
        m = ra.build_wo_extra(depth=1, embed_dim=768, num_heads=12, num_classes=10, num_layers=2, proj_dim=768, backbone="r50")
        m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=9))
        # the adapter's D4/D5 patch is a local subclass; a real checkpoint pickles the reference's own class path
        m.vision_transformer.__class__ = ra.load_reference()["sa"].MultiscaleFormer
        torch.save({{"epoch": 7, "model": m, "train_acc": 0.8, "test_acc": 0.76}}, {ckpt!r})   # main_toy.py:139-149
    """)
    subprocess.run([sys.executable, "-c", save], check=True, capture_output=True)
    load = textwrap.dedent(f"""
        import sys, torch
        sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})
        assert not any('reference' in p for p in sys.path)
        import duoformer_tcga_b200 as duo
        from duoformer_tcga_b200 import checkpoint
        from oracle import synth
        m = duo.MyModel_no_extra_params(depth=1, num_layers=2, pretrained=False, embed_dim=768, num_heads=12, num_classes=10, proj_dim=768)
        try:  # a whole-module pickle executes code on load: refused unless the caller vouches for it
            checkpoint.load_checkpoint(m, {ckpt!r})
            raise SystemExit("untrusted pickle was loaded")
        except RuntimeError as e:
            assert "trust_pickle" in str(e)
        missing, unexpected = checkpoint.load_checkpoint(m, {ckpt!r}, trust_pickle=True)
        assert not missing and not unexpected, (missing, unexpected)
        want = synth.synth_state_dict(m.state_dict(), seed=9)
        assert all(torch.equal(m.state_dict()[k], want[k]) for k in want)
        assert 'timm' not in sys.modules and 'model_wo_extra_params' not in sys.modules
        print('OK')
    """)
    r = subprocess.run([sys.executable, "-c", load], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_state_dict_file_loads_with_weights_only(tmp_path):
    """A tensors-only checkpoint file needs no trust flag (torch.load(weights_only=True))."""
    src = _small()
    path = str(tmp_path / "sd.pt")
    torch.save({"epoch": 3, "model_state_dict": src.state_dict()}, path)
    dst = _small()
    missing, unexpected = checkpoint.load_checkpoint(dst, path)
    assert not missing and not unexpected
    assert all(torch.equal(dst.state_dict()[k], v) for k, v in src.state_dict().items())
