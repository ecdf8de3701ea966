"""fp16 hazards of the bf16-mode trunk (-m gpu).  The cuDNN trunk copy and the token-builder GEMM run on IEEE fp16
(11-bit mantissa, maximum 65504): these tests drive it with the inputs real TCGA pipelines produce — BatchNorm
statistics with small running_var (BN-folded weights scaled by gamma / sqrt(var) up to ~30x) and un-normalised /
badly scaled tiles — and check the range guard (token_builder.TrunkRunner: power-of-two activation rescaling chosen
by a bf16 calibration pass) against the fp32 oracle.  Tolerance: the bf16 bar of BASELINE.json, 2e-2 relative."""
import pytest
import torch

from common import build_product, load_golden, oracle_forward, relerr
from oracle import duoformer_oracle as orc
from oracle import synth

pytestmark = pytest.mark.gpu


def _small_running_var(sd, seed=0):
    """Re-parametrise every trunk BatchNorm to running_var in [1e-3, 1] (log-uniform per channel) WITHOUT changing the
    network function: the convolution feeding it is scaled per output channel by s = sqrt((v + eps) / (var + eps)) and
    the running mean by s, as a trained checkpoint whose pre-BN activations are small would look."""
    g = torch.Generator().manual_seed(seed)
    out = dict(sd)
    eps = 1e-5
    for k in list(sd.keys()):
        if not (k.startswith("resnet_projector.") and k.endswith("running_var")):
            continue
        bn = k[: -len("running_var")]
        parent, leaf = bn[:-1].rsplit(".", 1)
        if leaf.startswith("bn"):
            conv = f"{parent}.conv{leaf[2:]}.weight"
        else:  # Sequential index: '1' after '0' (stem, downsample)
            conv = f"{parent}.{int(leaf) - 1}.weight"
        assert conv in sd, (k, conv)
        v = 10.0 ** (-3.0 * torch.rand(sd[k].shape, generator=g))
        s = torch.sqrt((v + eps) / (sd[k] + eps))
        out[k] = v
        out[bn + "running_mean"] = sd[bn + "running_mean"] * s
        out[conv] = sd[conv] * s[:, None, None, None]
    return out


def _model(sd_fn=None):
    gold = load_golden("wo4_d2")
    case = gold["case"]
    model = build_product(case)
    sd = synth.synth_state_dict(model.state_dict(), seed=gold["weight_seed"])
    if sd_fn is not None:
        sd = sd_fn(sd)
    model.load_state_dict(sd)
    return case, sd, model.cuda().eval()


def test_small_running_var_and_16x_inputs():
    case, sd, model = _model(_small_running_var)
    assert min(float(v.min()) for k, v in sd.items() if k.endswith("running_var") and k.startswith("resnet_projector")) < 2e-3
    x = synth.synth_images(2, seed=31) * 16.0
    with torch.no_grad():
        yo = oracle_forward(case, x, sd)
        y = model(x.cuda()).float().cpu()
        feats = model.get_features(x.cuda())
    assert all(torch.isfinite(f).all() for f in feats.values())
    assert torch.isfinite(y).all()
    assert relerr(y, yo) < 2e-2
    assert torch.equal(y.argmax(-1), yo.argmax(-1))


@pytest.mark.parametrize("scale", [1024.0, 4096.0])
def test_range_guard_engages_on_unnormalised_tiles(scale):
    """Inputs three orders of magnitude above ImageNet normalisation: an unguarded fp16 trunk overflows (stage maps
    reach 3.4e4 x scale / 1024); the guard rescales by a power of two and the logits still meet the bf16 bar."""
    case, sd, model = _model()
    x = synth.synth_images(2, seed=32) * scale
    with torch.no_grad():
        ocap = {}
        yo = oracle_forward(case, x, sd, capture=ocap)
        y = model(x.cuda()).float().cpu()
        tr = model._trunk_runner
        assert tr.act_scale < 1.0 and tr.calibration_max > 4096.0, (tr.act_scale, tr.calibration_max)
        feats = model.get_features(x.cuda())  # public API: true magnitudes
    for k, f in feats.items():
        assert torch.isfinite(f).all()
        assert relerr(f, ocap["features"][k]) < 1e-2, k
    assert torch.isfinite(y).all()
    assert relerr(y, yo) < 2e-2
    assert torch.equal(y.argmax(-1), yo.argmax(-1))
    # the same weights on ordinary inputs recalibrate only when the weights change: the pinned factor stays valid
    with torch.no_grad():
        x1 = synth.synth_images(2, seed=33)
        y1 = model(x1.cuda()).float().cpu()
        assert relerr(y1, oracle_forward(case, x1, sd)) < 2e-2


def test_unguarded_fp16_trunk_would_overflow():
    """Documents the hazard: with the guard pinned off (act_scale_override = 1) the same inputs give non-finite fp16
    stage maps, and the path RAISES instead of returning garbage."""
    case, sd, model = _model()
    model._trunk_runner.act_scale_override = 1.0
    x = synth.synth_images(2, seed=32) * 4096.0
    with torch.no_grad(), pytest.raises(RuntimeError, match="non-finite"):
        model(x.cuda())


@pytest.mark.parametrize("backend", ["own", "cudnn"])
def test_fused_trunk_mismatch_raises(monkeypatch, backend):
    """The fast trunk path (own implicit-GEMM convolutions, or cuDNN's fused conv+bias(+add)+ReLU calls) is verified once
    per weight set against the module path; a mismatch is an error — there is no silent fallback."""
    from duoformer_tcga_b200 import token_builder as tb
    from duoformer_tcga_b200 import trunk_convs

    case, sd, model = _model()
    model._trunk_runner.backend = backend

    def skew(real):
        def skewed(*args, **kwargs):
            f = real(*args, **kwargs)
            f[2] = f[2] * 1.05
            return f
        return skewed

    if backend == "own":
        monkeypatch.setattr(trunk_convs.OwnTrunk, "features", skew(trunk_convs.OwnTrunk.features))
    else:
        monkeypatch.setattr(tb, "_fused_trunk_forward", skew(tb._fused_trunk_forward))
    with torch.no_grad(), pytest.raises(RuntimeError, match="differs from the module path"):
        model(synth.synth_images(2, seed=34).cuda())

