"""Implicit-GEMM convolution kernel (csrc/conv_tcgen05.cu) and the own-kernel ResNet trunk (-m gpu).

Kernel tests compare duo_conv2d / duo_stem_* with torch's fp32 convolution of the SAME 16-bit-rounded inputs (the only
difference left is fp32 accumulation order and the final 16-bit rounding: 2e-3 relative of the maximum for fp16 outputs,
1e-2 for bf16).  Shapes: every (map size, channels, kernel, stride) class of the ResNet-50 trunk at 224 and 384 pixels, the
channel-token branch (3840 -> 768 on 7 x 7), ragged batches (batch rows past B are clipped by the TMA box store)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from duoformer_tcga_b200 import ops, token_builder, trunk_convs  # noqa: E402


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def _gen(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


CASES = [
    # B, H, Cin, Cout, k, stride, residual
    (4, 56, 64, 64, 1, 1, False),      # layer1 conv1 (first block)
    (4, 56, 64, 64, 3, 1, False),      # layer1 conv2
    (4, 56, 64, 256, 1, 1, True),      # layer1 conv3 + identity
    (4, 56, 256, 64, 1, 1, False),     # layer1 conv1
    (3, 56, 256, 128, 1, 1, False),    # layer2 conv1, ragged batch (box batch 2)
    (4, 56, 128, 128, 3, 2, False),    # layer2 conv2 stride 2
    (4, 56, 256, 512, 1, 2, False),    # layer2 downsample 1x1 stride 2
    (9, 28, 128, 512, 1, 1, True),     # layer2 conv3, batch not a multiple of the box batch 8
    (5, 28, 256, 256, 3, 2, False),    # layer3 conv2 stride 2
    (33, 14, 256, 1024, 1, 1, True),   # layer3 conv3 (box batch 32)
    (6, 14, 512, 512, 3, 2, False),    # layer4 conv2 stride 2
    (130, 7, 512, 2048, 1, 1, True),   # layer4 conv3, two batch tiles (box batch 128)
    (6, 7, 512, 512, 3, 1, False),     # layer4 conv2
    (2, 96, 64, 64, 3, 1, False),      # 384 x 384 input: layer1 map
    (2, 12, 512, 512, 3, 1, False),    # 384 x 384 input: layer4 map
    (2, 24, 512, 512, 3, 2, False),    # 384 x 384 input: layer4 stride 2
    (3, 7, 3840, 768, 3, 1, False),    # channel-token branch, first 3x3
    (2, 10, 64, 128, 3, 1, True),      # map size with a single factor of two
    (200, 14, 256, 256, 3, 1, False),  # enough tiles for the 128 x 256 tile variant (layer 3 at batch >= 200)
    (200, 14, 256, 512, 1, 1, True),   # same size with a residual (keeps 128 x 128 tiles)
]


@pytest.mark.parametrize("B,H,Cin,Cout,k,stride,res", CASES)
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_conv2d_vs_torch(B, H, Cin, Cout, k, stride, res, dtype):
    x = _gen((B, H, H, Cin), 1).to(dtype)
    w = _gen((Cout, Cin, k, k), 2, (2.0 / (Cin * k * k)) ** 0.5).to(dtype)
    bias = _gen((Cout,), 3)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=stride, padding=k // 2)
    Ho = ref.shape[2]
    r = _gen((B, Ho, Ho, Cout), 4).to(dtype) if res else None
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    wk = w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous()
    tol = 2e-3 if dtype == torch.float16 else 1e-2
    for relu in (True, False):
        out = ops.conv2d(x, wk, bias, k, stride, relu, r)
        want = (ref.clamp_min(0) if relu else ref).permute(0, 2, 3, 1)
        assert out.shape == want.shape
        assert relerr(out, want) < tol, (relu, relerr(out, want))
    out = ops.conv2d(x, wk, None, k, stride, False, r)
    assert relerr(out, (ref - bias[None, :, None, None]).permute(0, 2, 3, 1)) < tol


@pytest.mark.parametrize("B,H,Cmid,Cin2,Cout,stride2", [(4, 56, 64, 64, 256, 1), (5, 28, 128, 256, 512, 2), (3, 7, 512, 1024, 2048, 2),
                                                       (2, 14, 256, 512, 1024, 2)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_conv2d_fused_projection_shortcut(B, H, Cmid, Cin2, Cout, stride2, dtype):
    """conv3 (1x1) + downsample (1x1, stride s) of the block input in one K loop == the sum of the two convolutions."""
    o = _gen((B, H, H, Cmid), 1).to(dtype)
    x = _gen((B, H * stride2, H * stride2, Cin2), 2).to(dtype)
    w3 = _gen((Cout, Cmid, 1, 1), 3, (1.0 / Cmid) ** 0.5).to(dtype)
    wd = _gen((Cout, Cin2, 1, 1), 4, (1.0 / Cin2) ** 0.5).to(dtype)
    bias = _gen((Cout,), 5)
    ref = F.conv2d(o.float().permute(0, 3, 1, 2), w3.float(), bias) + F.conv2d(x.float().permute(0, 3, 1, 2), wd.float(), None, stride=stride2)
    ref = ref.clamp_min(0).permute(0, 2, 3, 1)
    w = torch.cat([w3.reshape(Cout, Cmid), wd.reshape(Cout, Cin2)], dim=1).contiguous()
    out = ops.conv2d(o, w, bias, 1, 1, True, in2=x, stride2=stride2)
    assert relerr(out, ref) < (2e-3 if dtype == torch.float16 else 1e-2)


@pytest.mark.parametrize("B,H,W,Cin,Cout,k,stride", [(2, 12, 20, 64, 64, 3, 1), (3, 9, 15, 64, 128, 3, 2), (1, 5, 24, 128, 64, 1, 1),
                                                     (2, 16, 6, 64, 64, 3, 2), (1, 1, 1, 64, 64, 3, 1)])
def test_conv2d_non_square_maps(B, H, W, Cin, Cout, k, stride):
    """Rectangular and odd-sized maps: the tile box takes whatever powers of two divide the OUTPUT width / height (down to
    1 x 1 x 128 pixels), the rest comes from the batch."""
    x = _gen((B, H, W, Cin), 1).half()
    w = _gen((Cout, Cin, k, k), 2, (2.0 / (Cin * k * k)) ** 0.5).half()
    bias = _gen((Cout,), 3)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), bias, stride=stride, padding=k // 2).clamp_min(0).permute(0, 2, 3, 1)
    out = ops.conv2d(x, w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous(), bias, k, stride, True)
    assert out.shape == ref.shape
    assert relerr(out, ref) < 2e-3


def test_conv2d_fp16_operands_bf16_output():
    x = _gen((4, 28, 28, 128), 1, 30.0).to(torch.float16)
    w = _gen((256, 128, 3, 3), 2, 1.0).to(torch.float16)  # sums far beyond the fp16 maximum
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), None, stride=2, padding=1).permute(0, 2, 3, 1)
    out = ops.conv2d(x, w.permute(0, 2, 3, 1).reshape(256, -1).contiguous(), None, 3, 2, False, out_dtype=torch.bfloat16)
    assert out.dtype == torch.bfloat16 and torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 1e-2


def test_conv2d_writes_nothing_outside_its_output():
    B, H, Cin, Cout = 3, 14, 64, 128  # box batch 32: 29 clipped batch rows per tile
    x = _gen((B, H, H, Cin), 1).to(torch.float16)
    w = _gen((Cout, 9 * Cin), 2, 0.05).to(torch.float16)
    guard = 4096
    buf = torch.full((B * H * H * Cout + 2 * guard,), 7.0, dtype=torch.float16, device="cuda")
    out = buf[guard:guard + B * H * H * Cout].view(B, H, H, Cout)
    ops.conv2d(x, w, None, 3, 1, False, out=out)
    torch.cuda.synchronize()
    assert (buf[:guard] == 7.0).all() and (buf[-guard:] == 7.0).all()


@pytest.mark.parametrize("B,H,W", [(2, 224, 224), (3, 64, 96), (1, 384, 384)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_stem_vs_torch(B, H, W, dtype):
    x = _gen((B, 3, H, W), 1)
    w = _gen((64, 3, 7, 7), 2, 0.1).to(dtype)
    bias = _gen((64,), 3)
    scale = 0.5
    packed = ops.stem_pack(x, scale, dtype)
    assert packed.shape == (B, H + 8, W + 8, 8)
    want = torch.zeros(B, H + 8, W + 8, 8, device="cuda")
    want[:, 3:3 + H, 3:3 + W, :3] = (x * scale).permute(0, 2, 3, 1)    # padded pixel (R, X)
    want[:, 2:2 + H, 3:3 + W, 4:7] = (x * scale).permute(0, 2, 3, 1)   # padded pixel (R + 1, X)
    assert torch.equal(packed.float(), want.to(dtype).float())
    # channels-last input gives the same packed tensor
    assert torch.equal(ops.stem_pack(x.contiguous(memory_format=torch.channels_last), scale, dtype), packed)
    out = ops.stem_conv7x7(packed, ops.pack_stem_weight(w, dtype), bias, relu=True)
    ref = F.conv2d((x * scale).to(dtype).float(), w.float(), bias, stride=2, padding=3).clamp_min(0).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert relerr(out, ref) < (2e-3 if dtype == torch.float16 else 1e-2)


def _resnet50_folded(seed=0):
    import torchvision

    torch.manual_seed(seed)
    m = torchvision.models.resnet50(weights=None)
    for mod in m.modules():  # non-trivial BatchNorm statistics
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
            mod.weight.data.uniform_(0.5, 1.5)
            mod.bias.data.normal_(0, 0.1)
    trunk = torch.nn.Sequential(*list(m.children())[:-2]).eval().cuda()
    return trunk


@pytest.mark.parametrize("size,B", [(224, 3), (384, 2)])
def test_own_trunk_matches_module_path(size, B):
    """The four stage maps of the own-kernel trunk against the fp32 torch modules (fp16 trunk: 5e-3 of the maximum) and
    against the cuDNN backend at the same precision."""
    trunk = _resnet50_folded()
    x = _gen((B, 3, size, size), 5)
    with torch.no_grad():
        ref = token_builder.TrunkRunner._plain_forward(trunk, x, False)
    own = token_builder.TrunkRunner()
    f_own = own.features(trunk, x, "bf16", False)
    assert own._own is not None, "ResNet-50 must be eligible for the own convolution kernels"
    lib = token_builder.TrunkRunner()
    lib.backend = "cudnn"
    f_lib = lib.features(trunk, x, "bf16", False)
    assert lib._own is None
    assert own.act_scale == lib.act_scale
    for k in range(4):  # the maps come multiplied by the fp16 range-guard factor
        assert f_own[k].shape == ref[k].shape and f_own[k].dtype == torch.float16
        e_own, e_lib = relerr(f_own[k].float() / own.act_scale, ref[k]), relerr(f_lib[k].float() / lib.act_scale, ref[k])
        assert e_own < 5e-3, (k, e_own)
        assert e_own < 2 * e_lib + 1e-3, (k, e_own, e_lib)
    # second call: the verified fast path alone
    f2 = own.features(trunk, x, "bf16", False)
    for k in range(4):
        assert torch.equal(f2[k], f_own[k])
