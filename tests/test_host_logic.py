"""CPU (-m "not gpu"): host-side logic — C-ABI surface, state_dict schema, factories' signatures,
index tables / scatter maps, weight packing helpers, error behaviour without a GPU."""
import ctypes
import inspect
import os
import re

import pytest
import torch

from common import ROOT, build_product, COMMON
import duoformer_tcga_b200 as duo
from duoformer_tcga_b200 import _lib, index_tables, ops
from oracle import duoformer_oracle as orc


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "duoformer_sm100.h")).read()
    declared = set(re.findall(r"\b(duo_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found in the header"
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    lib = _lib.load()  # built by __graft_entry__.build(); fails loudly if missing
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.duo_abi_version() == _lib.ABI_VERSION == 4
    assert lib.duo_last_error() is not None


def test_gemm_args_struct_layout_matches_header():
    # 7 pointers, 4 int64, 7 int32 + float + 2 int32, 4 pointers -> 8*7 + 8*4 + 4*10 + 8*4 = 160 bytes
    assert ctypes.sizeof(_lib.GemmArgs) == 160
    assert _lib.GemmArgs.xb_out.offset == 128 and _lib.GemmArgs.ln_stats.offset == 144 and _lib.GemmArgs.shift_stats.offset == 152
    assert _lib.GemmArgs.M.offset == 56 and _lib.GemmArgs.N.offset == 88


def test_conv2d_args_struct_layout_matches_header():
    # struct duo_conv2d_args: five pointers, ten int32 (B H W Cin Cout ksize stride relu fp16 out_fp16), in2, four int32
    assert ctypes.sizeof(_lib.Conv2dArgs) == 104
    assert _lib.Conv2dArgs.out.offset == 32 and _lib.Conv2dArgs.B.offset == 40 and _lib.Conv2dArgs.out_fp16.offset == 76
    assert _lib.Conv2dArgs.in2.offset == 80 and _lib.Conv2dArgs.stride2.offset == 100


def test_own_trunk_eligibility():
    """ResNet-50 (Bottlenecks, BatchNorm folded) runs on the own convolution kernels; r18 BasicBlocks and un-folded
    trunks do not (they keep the cuDNN path)."""
    import torchvision

    from duoformer_tcga_b200 import token_builder, trunk_convs

    r18 = torch.nn.Sequential(*list(torchvision.models.resnet18(weights=None).children())[:-2]).eval()
    token_builder._fold_batchnorm_(r18)
    assert not trunk_convs.eligible(r18, False)
    r50 = torch.nn.Sequential(*list(torchvision.models.resnet50(weights=None).children())[:-2]).eval()
    assert not trunk_convs.eligible(r50, False)  # BatchNorm not folded yet
    token_builder._fold_batchnorm_(r50)
    assert trunk_convs.eligible(r50, False)
    own = trunk_convs.OwnTrunk(r50, False, torch.float16)  # packing works on the host
    assert own.stem_w.shape == (64, 256) and len(own.layers) == 4 and [len(l) for l in own.layers] == [3, 4, 6, 3]
    assert own.layers[1][0].fused_shortcut and own.layers[1][0].c2.stride == 2 and not own.layers[1][1].fused_shortcut
    assert own.layers[1][0].c3.weight.shape == (512, 128 + 256) and own.layers[1][0].c3.stride2 == 2
    assert own.layers[0][0].c2.weight.shape == (64, 9 * 64)


def test_conv_tile_coordinates_multiply_shift_is_exact():
    """csrc/conv_tcgen05.cu takes tile coordinates with floor(x / d) = ((x << 24) * ceil(2^40 / d)) >> 64 instead of integer
    divisions; exact for x < 2^24 tiles and d < 2^15 (the host refuses larger problems).  Restated here on Python integers."""
    import random

    def fast_div(x, d):
        return ((x << 24) * (((1 << 40) + d - 1) // d)) >> 64

    rnd = random.Random(0)
    for _ in range(200000):
        d, x = rnd.randint(1, (1 << 15) - 1), rnd.randint(0, (1 << 24) - 1)
        assert fast_div(x, d) == x // d, (x, d)
    for d in (1, 2, 3, 7, 49, 98, 392, 6272, 25088, 32767):
        for x in list(range(0, 300)) + [(1 << 24) - 1, d * 511 - 1, d * 511]:
            if x < (1 << 24):
                assert fast_div(x, d) == x // d, (x, d)


def test_ops_refuse_cpu_tensors():
    a = torch.zeros(128, 64, dtype=torch.bfloat16)
    w = torch.zeros(128, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(a, w, None, torch.zeros(128, 128, dtype=torch.bfloat16), ops.EPI_BF16)


def test_models_refuse_cpu_forward():
    m = duo.MyModel_no_extra_params(depth=1, num_layers=2, pretrained=False, **COMMON).eval()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 3, 224, 224))
    with pytest.raises(NotImplementedError):
        m.vision_transformer(torch.zeros(1, 49, 6, 768))


REF_SIGNATURES = {
    # reference defaults, models/__init__.py:12-24, :40-54 ; model.py:23-36 ; model_wo_extra_params.py:30-47
    "build_model": dict(depth=12, patch_size=49, embed_dim=256, num_heads=6, init_values=1e-5, num_classes=100,
                        num_layers=4, proj_dim=384, model_ver="scaleformer", pretrained=True, freeze=True),
    "build_model_no_extra_params": dict(depth=12, embed_dim=256, num_heads=6, num_classes=100, num_layers=4,
                                        num_patches=49, proj_dim=384, mlp_ratio=4.0, attn_drop_rate=0.0,
                                        proj_drop_rate=0.0, freeze_backbone=True, backbone="r50", pretrained=True),
    "MyModel": dict(depth=None, patch_size=49, embed_dim=256, num_heads=6, init_values=1e-5, num_classes=2,
                    num_layers=4, proj_dim=512, model_ver="originalViT", pretrained=True, freeze=True),
    "MyModel_no_extra_params": dict(depth=None, embed_dim=768, num_heads=12, init_values=1e-5, num_classes=2,
                                    num_layers=4, num_patches=49, mlp_ratio=4.0, attn_drop_rate=0.0,
                                    proj_drop_rate=0.0, proj_dim=768, freeze_backbone=True, backbone="r50",
                                    scale_token="random", patch_attn=True),
}


@pytest.mark.parametrize("name", sorted(REF_SIGNATURES))
def test_public_signatures_keep_reference_names_order_defaults(name):
    obj = getattr(duo, name)
    params = [p for p in inspect.signature(obj).parameters.values() if p.name != "self"]
    want = REF_SIGNATURES[name]
    got = [(p.name, p.default) for p in params[: len(want)]]
    assert got == list(want.items())
    for p in params[len(want):]:  # additive extensions must be optional keywords
        assert p.default is not inspect.Parameter.empty


def test_state_dict_schema_counts():
    """526 tensors / 139.98 M params (wo-extra 4-scale), 629 / 187.72 M (MyModel)."""
    wo = duo.build_model_no_extra_params(depth=12, num_layers=4, pretrained=False, **COMMON)
    sd = wo.state_dict()
    assert len(sd) == 526
    assert abs(sum(p.numel() for p in wo.parameters()) / 1e6 - 139.98) < 0.01
    assert sd["vision_transformer.pos_embed_for_scale"].shape == (1, 1, 86, 768)
    assert sd["vision_transformer.scaleBlocks.0.attn.qkv.weight"].shape == (2304, 768)
    assert "vision_transformer.fc_norm.weight" in sd and "channel_token" in sd
    assert sd["projection.proj_heads0.weight"].shape == (768, 256, 1, 1)
    mm = duo.build_model(depth=12, patch_size=32, embed_dim=768, num_heads=12, num_classes=10, num_layers=2,
                         proj_dim=768, pretrained=False)
    sd = mm.state_dict()
    assert len(sd) == 629
    assert abs(sum(p.numel() for p in mm.parameters()) / 1e6 - 187.72) < 0.01
    for k in ("vision_transformer.patch_embed.proj.weight", "vision_transformer.blocks.0.attn.q_norm.weight",
              "vision_transformer.blocks.3.attn.qkv1.weight", "vision_transformer.blocks.0.ls1.gamma",
              "vision_transformer.norm.weight", "chann_proj_all.nConvs.0.norm.running_var"):
        assert k in sd, k
    assert sd["vision_transformer.pos_embed"].shape == (1, 50, 768)


def test_trunk_key_naming_by_backbone():
    a = duo.MyModel_no_extra_params(depth=1, num_layers=2, backbone="r50", pretrained=False, **COMMON)
    b = duo.MyModel_no_extra_params(depth=1, num_layers=2, backbone="r50_Swav", pretrained=False, **COMMON)
    assert "resnet_projector.0.weight" in a.state_dict() and "resnet_projector.7.2.bn3.running_var" in a.state_dict()
    assert "resnet_projector.conv1.weight" in b.state_dict() and "resnet_projector.layer4.2.bn3.running_var" in b.state_dict()


def test_constructor_validation():
    with pytest.raises(ValueError):
        duo.build_model_no_extra_params(pretrained=False)  # reference defaults 256 != 384
    with pytest.raises(AssertionError):
        duo.MyModel_no_extra_params(depth=1, embed_dim=768, proj_dim=768, num_heads=7, pretrained=False)
    with pytest.raises(NotImplementedError):
        duo.MyModel_no_extra_params(depth=1, num_layers=2, attn_drop_rate=0.1, pretrained=False, **COMMON)
    m = duo.MyModel_no_extra_params(depth=1, num_layers=2, pretrained=False, **COMMON)
    with pytest.raises(ValueError):
        m.set_precision("fp16")
    assert m.set_precision("fp32").precision == "fp32"
    assert m.name == "scaleformer" and m.num_layers == 2


@pytest.mark.parametrize("g", [7, 12])
@pytest.mark.parametrize("k", [3, 2, 1, 0])
def test_gather_index_equals_oracle_table(k, g):
    assert torch.equal(index_tables.gather_index(k, g), orc.index_table(k, g))


@pytest.mark.parametrize("num_layers", [2, 3, 4])
@pytest.mark.parametrize("g", [7, 12])
def test_token_row_maps_are_a_bijection_onto_non_scale_rows(num_layers, g):
    S = index_tables.num_scale_tokens(num_layers)
    assert S == orc.num_scale_tokens(num_layers) == {2: 6, 3: 22, 4: 86}[num_layers]
    maps = index_tables.token_row_maps(num_layers, g)
    assert sorted(maps) == sorted(index_tables.stages_used(num_layers))
    rows = torch.cat([m for m in maps.values()]).long()
    P = g * g
    assert rows.numel() == P * (S - 1) and rows.unique().numel() == rows.numel()
    assert (rows % S != 0).all() and rows.min() >= 1 and rows.max() == P * S - 1
    # consistency with the reference-style gather: token (p, s) of stage k comes from pixel idx[p, j]
    s_off = 1
    for k in index_tables.stages_used(num_layers):
        idx = orc.index_table(k, g)
        w2 = idx.shape[1]
        for p in (0, P // 2, P - 1):
            for j in (0, w2 - 1):
                assert maps[k][idx[p, j]].item() == p * S + s_off + j
        s_off += w2


def test_split_weight_reconstructs_to_16_bits():
    w = torch.randn(64, 128)
    s = ops.split_weight(w)
    assert s.shape == (64, 256) and s.dtype == torch.bfloat16
    rec = s[:, :128].float() + s[:, 128:].float()
    assert ((rec - w).abs() / w.abs().clamp_min(1e-3)).max().item() < 2 ** -15


def test_pack_cache_invalidates_on_parameter_update():
    from duoformer_tcga_b200 import engine

    blk = duo.PatchBlock(768, 12, qkv_bias=True)
    p1 = blk.pack("bf16")
    assert blk.pack("bf16") is p1
    with torch.no_grad():
        blk.attn.qkv.weight.add_(1.0)
    p2 = blk.pack("bf16")
    assert p2 is not p1 and not torch.equal(p1["qkv"][0], p2["qkv"][0])
    p3 = blk.pack("fp32")
    assert p3["qkv"][0].shape == (2304, 1536)


def test_token_row_maps_properties_over_random_grids():
    """hypothesis: for any patch grid g and 1..4 scales the scatter maps tile the non-scale-token rows
    exactly once, and every pixel lands in the patch that geometrically contains it."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=25, deadline=None)
    @given(g=st.integers(min_value=1, max_value=13), num_layers=st.integers(min_value=1, max_value=4))
    def check(g, num_layers):
        S = index_tables.num_scale_tokens(num_layers)
        maps = index_tables.token_row_maps(num_layers, g)
        rows = torch.cat(list(maps.values())).long()
        assert rows.numel() == g * g * (S - 1) == rows.unique().numel()
        for k, m in maps.items():
            w = 2 ** (3 - k)
            G = g * w
            ys, xs = torch.meshgrid(torch.arange(G), torch.arange(G), indexing="ij")
            patch_of_pixel = (ys // w) * g + (xs // w)
            assert torch.equal((m.long() // S).view(G, G), patch_of_pixel)

    check()


def test_pack_ln_linear_centres_the_rounded_rows_and_matches_layernorm_linear():
    """Consumer-side operands of the LayerNorm statistics forwarding (engine.pack_ln_linear), emulated on the CPU:
    the bf16 VALUES of every weight row sum to ~0 (so a per-row constant in x — its mean, or the producer's shift —
    drops out of x W''^T), and rstd * (bf16(x - shift) W''^T) + bias' reproduces Linear(LayerNorm(x)) within bf16
    rounding even when the row mean is 100x the row spread, provided the shift is near the mean."""
    from duoformer_tcga_b200 import engine

    g = torch.Generator().manual_seed(0)
    K, N = 768, 3072
    W = torch.randn(N, K, generator=g) * 0.02
    b = torch.randn(N, generator=g) * 0.02
    lw = 1 + 0.1 * torch.randn(K, generator=g)
    lb = 0.05 * torch.randn(K, generator=g)
    wp, bp = engine.pack_ln_linear(W, b, lw, lb)
    assert wp.dtype == torch.bfloat16 and wp.shape == (N, K) and bp.shape == (N,)
    row_sums = wp.double().sum(dim=1).abs().max().item()
    naive = ((W * lw) - (W * lw).mean(dim=1, keepdim=True)).to(torch.bfloat16).double().sum(dim=1).abs().max().item()
    assert row_sums < 1e-5 * wp.float().abs().mean().item() * K and row_sums < 1e-3 * naive
    x = torch.randn(256, K, generator=g) * 3
    for offset in (0.0, 30.0, 300.0):
        xo = x + offset
        ref = torch.nn.functional.layer_norm(xo, (K,), lw, lb, 1e-6) @ W.t() + b
        rstd = torch.rsqrt(xo.var(dim=-1, unbiased=False, keepdim=True) + 1e-6)
        shift = xo.mean(dim=-1, keepdim=True) + 0.3  # an estimate of the mean (the previous block's)
        out = rstd * ((xo - shift).to(torch.bfloat16).float() @ wp.float().t()) + bp
        ln_round = torch.nn.functional.layer_norm(xo, (K,), lw, lb, 1e-6).to(torch.bfloat16).float() @ W.to(torch.bfloat16).float().t() + b
        e_fwd = ((out - ref).abs().max() / ref.abs().max()).item()
        e_ln = ((ln_round - ref).abs().max() / ref.abs().max()).item()
        assert e_fwd < 2.5 * e_ln + 1e-3, (offset, e_fwd, e_ln)


def test_library_sources_have_no_environment_switches_or_cpu_paths():
    """The product library selects its code paths from its arguments only: no getenv-controlled variants (round-1
    scaffolding), and the Python package never imports the oracle."""
    import glob
    import os

    from common import ROOT

    for path in glob.glob(os.path.join(ROOT, "duoformer_tcga_b200", "csrc", "*.cu*")):
        text = open(path).read()
        assert "getenv" not in text, f"{os.path.basename(path)} reads the environment"
    for path in glob.glob(os.path.join(ROOT, "duoformer_tcga_b200", "*.py")):
        text = open(path).read()
        assert "import oracle" not in text and "from oracle" not in text, f"{os.path.basename(path)} imports the oracle"
