"""Thin torch-tensor front-ends of the C ABI (one function per entry point).

PyTorch is used for device memory and streams only; every function enqueues hand-written
sm_100a kernels on ``torch.cuda.current_stream()``.  No CPU / eager fallback exists: a
non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes
import functools
from typing import Optional

import torch

from . import _lib
from ._lib import (  # noqa: F401  (re-exported)
    ACT_BF16,
    ACT_F32,
    ACT_SPLIT,
    EPI_BF16,
    EPI_F32,
    EPI_GELU_BF16,
    EPI_GELU_SPLIT_BF16,
    EPI_RESIDUAL_F32,
    EPI_SCATTER_F32,
    EPI_SPLIT_BF16,
)


# Optional per-launch CUDA-event timing (bench.py's profiled pass): when a list is installed here, gemm(),
# layernorm() and group_attention() append (start_event, end_event, kind, algorithmic flops, algorithmic bytes, tag).
PROFILE = None


def _prof_begin():
    if PROFILE is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return e0


def _prof_end(e0, kind, flops, nbytes, tag):
    if e0 is not None and PROFILE is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        PROFILE.append((e0, e1, kind, float(flops), float(nbytes), tag))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_operand_device(fn):
    """Run `fn` with the operands' device current: the library launches on the CURRENT device (kernel
    attributes, SM count, tensor maps), so tensors living on another device must switch it first.  All
    tensor operands must share one device."""

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for v in list(args) + list(kwargs.values()):
            if isinstance(v, torch.Tensor) and v.is_cuda:
                if dev is None:
                    dev = v.device
                elif v.device != dev:
                    raise RuntimeError(f"{fn.__name__}: operands on different devices ({dev} and {v.device})")
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapper


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("duoformer_tcga_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()


def _act_kind(t: torch.Tensor, cols: int) -> int:
    if t.dtype == torch.float32:
        return ACT_F32
    if t.dtype != torch.bfloat16:
        raise RuntimeError(f"unsupported activation dtype {t.dtype}")
    return ACT_SPLIT if t.shape[-1] == 2 * cols else ACT_BF16


@_on_operand_device
def gemm(
    A: torch.Tensor,
    W: torch.Tensor,
    bias: Optional[torch.Tensor],
    out: torch.Tensor,
    epilogue: int,
    *,
    split3=False,
    gamma: Optional[torch.Tensor] = None,
    row_map: Optional[torch.Tensor] = None,
    rows_per_group: int = 0,
    dest_rows_per_group: int = 0,
    pos: Optional[torch.Tensor] = None,
    pos_period: int = 0,
    relu: bool = False,
    xb_out: Optional[torch.Tensor] = None,
    stats_out: Optional[torch.Tensor] = None,
    ln_stats: Optional[torch.Tensor] = None,
    ln_eps: float = 1e-6,
    shift_stats: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """out = epilogue(A @ W.T + bias).  A [M,K] / W [N,K] bf16 (or [.,2K] split when split3).

    LayerNorm statistics forwarding (include/duoformer_sm100.h):
      producer  EPI_RESIDUAL_F32 with xb_out (bf16 [M,N]) + stats_out (fp32 [M, N/256, 2]): the updated rows are
                also written un-normalised in bf16 together with their per-256-column (mean, M2) pairs; with
                shift_stats (the rows' PREVIOUS statistics, same layout) the copy is bf16(x - previous row mean);
      consumer  EPI_BF16 / EPI_GELU_BF16 with ln_stats: A is such a copy, W = W * ln_weight with centred rows
                (engine.pack_ln_linear), bias = W ln_bias + b; the epilogue applies rstd."""
    split3 = int(split3)  # 0 plain, 1 both operands split (hi|lo), 2 only W split (A exact bf16)
    assert A.dim() == 2 and W.dim() == 2 and W.dtype == A.dtype, "A and W must share one 16-bit format"
    assert A.dtype == torch.bfloat16 or (A.dtype == torch.float16 and split3 == 0), "operands must be bf16 (or fp16 in plain mode)"
    assert A.stride(1) == 1 and W.stride(1) == 1 and out.stride(-1) == 1
    M = A.shape[0]
    N = W.shape[0]
    K = W.shape[1] // 2 if split3 else W.shape[1]
    assert A.shape[1] == (2 * K if split3 == 1 else K), (A.shape, W.shape, split3)
    a = _lib.GemmArgs()
    a.A, a.W, a.bias, a.out = _ptr(A), _ptr(W), _ptr(bias), _ptr(out)
    a.gamma, a.row_map, a.pos = _ptr(gamma), _ptr(row_map), _ptr(pos)
    a.M, a.N, a.K = M, N, K
    a.lda, a.ldw = A.stride(0), W.stride(0)
    a.ldo = out.stride(-2) if out.dim() >= 2 else out.shape[-1]
    a.split3 = split3
    a.relu = 1 if relu else 0
    a.fp16_operands = 1 if A.dtype == torch.float16 else 0
    a.epilogue = epilogue
    a.rows_per_group, a.dest_rows_per_group, a.pos_period = rows_per_group, dest_rows_per_group, pos_period
    if xb_out is not None or stats_out is not None:
        assert xb_out.dtype == torch.bfloat16 and xb_out.is_contiguous() and xb_out.numel() == M * N
        assert stats_out.dtype == torch.float32 and stats_out.is_contiguous() and stats_out.numel() >= M * (N // 256) * 2
        a.xb_out, a.stats_out = _ptr(xb_out), _ptr(stats_out)
        if shift_stats is not None:
            assert shift_stats.dtype == torch.float32 and shift_stats.is_contiguous() and shift_stats.numel() >= M * (N // 256) * 2
            assert shift_stats.data_ptr() != stats_out.data_ptr()
            a.shift_stats = _ptr(shift_stats)
    if ln_stats is not None:
        assert ln_stats.dtype == torch.float32 and ln_stats.is_contiguous() and ln_stats.numel() >= M * (K // 256) * 2
        a.ln_stats, a.ln_eps = _ptr(ln_stats), float(ln_eps)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
    if row_map is not None:
        assert row_map.dtype == torch.int32
    e0 = _prof_begin()
    _lib.check(_lib.load().duo_gemm(ctypes.byref(a), _stream()), "duo_gemm")
    if e0 is not None:
        passes = 3 if split3 == 1 else (2 if split3 == 2 else 1)
        out_bytes = {EPI_BF16: 2, EPI_GELU_BF16: 2, EPI_RESIDUAL_F32: 8, EPI_SCATTER_F32: 4, EPI_F32: 4, EPI_SPLIT_BF16: 4,
                     EPI_GELU_SPLIT_BF16: 4}[epilogue] + (2 if xb_out is not None else 0)
        tag = f"{N}x{K}:epi{epilogue}" + ("+fwd" if xb_out is not None else "") + ("+ln" if ln_stats is not None else "")
        _prof_end(e0, "gemm", 2.0 * M * N * K * passes, M * (A.shape[1] * 2 + N * out_bytes) + W.numel() * 2, tag)
    return out


@_on_operand_device
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, out: torch.Tensor, eps: float,
              stats_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out (bf16 [rows,D] or split bf16 [rows,2D], dense) = LayerNorm(x fp32 [rows,D]).
    x may be a 2-D row-strided view (e.g. the s = 0 token of every patch).
    stats_out (fp32 [rows, D/256, 2]): also the per-256-column (mean, M2) pairs of every row."""
    D = x.shape[-1]
    assert x.dtype == torch.float32 and out.is_contiguous() and x.stride(-1) == 1
    if x.is_contiguous():
        rows, ldx = x.numel() // D, D
    else:
        assert x.dim() == 2, "strided LayerNorm input must be a 2-D view"
        rows, ldx = x.shape[0], x.stride(0)
    kind = _act_kind(out, D)
    assert kind in (ACT_BF16, ACT_SPLIT)
    if stats_out is not None:
        assert stats_out.dtype == torch.float32 and stats_out.is_contiguous() and stats_out.numel() >= rows * (D // 256) * 2
    e0 = _prof_begin()
    _lib.check(
        _lib.load().duo_layernorm(_ptr(x), _ptr(gamma), _ptr(beta), _ptr(out), kind, rows, D, ldx, float(eps),
                                  _ptr(stats_out), _stream()),
        "duo_layernorm",
    )
    _prof_end(e0, "layernorm", 0.0, rows * (D * 4 + out.shape[-1] * 2), f"layernorm:{rows}x{D}")
    return out


@_on_operand_device
def group_attention(
    qkv: torch.Tensor, out: torch.Tensor, S: int, num_heads: int, scale: float, algo: int = 0, q_rows: int = 0,
    split_in: bool = False,
) -> torch.Tensor:
    """softmax(q k^T * scale) v per (group of S rows, head); qkv [rows, 3*D], out [rows, D | 2D].
    q_rows > 0: only the first q_rows query rows per group; out is [groups * q_rows, D | 2D].
    split_in: qkv is split bf16 [rows, 2*3*D] (hi | lo, duo_gemm's SPLIT epilogue), out split [rows, 2D]:
    the split-precision tcgen05 kernel (S <= 64)."""
    assert qkv.is_contiguous() and out.is_contiguous()
    D = qkv.shape[-1] // (6 if split_in else 3)
    rows = qkv.numel() // ((6 if split_in else 3) * D)
    assert rows % S == 0 and D == 64 * num_heads
    if split_in:
        assert qkv.dtype == torch.bfloat16 and out.dtype == torch.bfloat16 and out.shape[-1] == 2 * D
    in_kind = ACT_SPLIT if split_in else (ACT_F32 if qkv.dtype == torch.float32 else ACT_BF16)
    out_kind = _act_kind(out, D)
    e0 = _prof_begin()
    _lib.check(
        _lib.load().duo_group_attention(
            _ptr(qkv), in_kind, _ptr(out), out_kind, rows // S, S, num_heads, float(scale), algo,
            q_rows if q_rows > 0 else S, _stream()
        ),
        "duo_group_attention",
    )
    if e0 is not None:
        qr = q_rows if q_rows > 0 else S
        in_bytes = rows * qkv.shape[-1] * qkv.element_size() * (1.0 if qr == S else (2.0 + qr / S) / 3.0)
        _prof_end(e0, "attention", 4.0 * (rows // S) * qr * S * D, in_bytes + out.numel() * out.element_size(),
                  f"attention:S{S}:q{qr}:in{in_kind}")
    return out


@_on_operand_device
def fill_scale_token(X: torch.Tensor, tok: torch.Tensor, pos0: torch.Tensor) -> torch.Tensor:
    """X[b,p,0,:] = tok[b,p,:] + pos0.  X fp32 [B,P,S,D]; tok [D] (broadcast) or [B,P,D]."""
    B, P, S, D = X.shape
    assert X.dtype == torch.float32 and X.is_contiguous() and tok.dtype == torch.float32
    if tok.numel() == D:
        sb, sp = 0, 0
    else:
        assert tok.shape == (B, P, D) and tok.stride(2) == 1
        sb, sp = tok.stride(0), tok.stride(1)
    _lib.check(
        _lib.load().duo_fill_scale_token(_ptr(X), _ptr(tok), sb, sp, _ptr(pos0), B, P, S, D, _stream()),
        "duo_fill_scale_token",
    )
    return X


@_on_operand_device
def add_pos(x: torch.Tensor, pos: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[..., s, :] = x[..., s, :] + pos[s, :]  (x fp32 [..., S, D] contiguous)."""
    S, D = x.shape[-2], x.shape[-1]
    assert x.dtype == torch.float32 and x.is_contiguous() and out.is_contiguous() and pos.is_contiguous()
    _lib.check(
        _lib.load().duo_add_pos(_ptr(x), _ptr(pos), _ptr(out), x.numel() // D, S, D, _stream()), "duo_add_pos"
    )
    return out


@_on_operand_device
def assemble_patch_tokens(X: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, Z: torch.Tensor) -> torch.Tensor:
    """Z[b,0]=cls+pos[0]; Z[b,1+p]=X[b,p,0]+pos[1+p].  Z bf16 [B,P+1,D] or split [B,P+1,2D]."""
    B, P, S, D = X.shape
    assert X.is_contiguous() and Z.is_contiguous() and pos.is_contiguous()
    kind = _act_kind(Z, D)
    _lib.check(
        _lib.load().duo_assemble_patch_tokens(_ptr(X), _ptr(cls), _ptr(pos), _ptr(Z), kind, B, P, S, D, _stream()),
        "duo_assemble_patch_tokens",
    )
    return Z


@_on_operand_device
def head(
    inp: torch.Tensor,
    row_stride: int,
    W: torch.Tensor,
    bias: Optional[torch.Tensor],
    logits: torch.Tensor,
    ln_gamma: Optional[torch.Tensor] = None,
    ln_beta: Optional[torch.Tensor] = None,
    eps: float = 1e-6,
) -> torch.Tensor:
    B, ncls = logits.shape
    D = W.shape[1]
    assert inp.dtype == torch.float32 and W.dtype == torch.float32 and W.is_contiguous()
    _lib.check(
        _lib.load().duo_head(
            _ptr(inp), row_stride, _ptr(ln_gamma), _ptr(ln_beta), float(eps), _ptr(W), _ptr(bias), _ptr(logits),
            B, D, ncls, _stream(),
        ),
        "duo_head",
    )
    return logits


@_on_operand_device
def convert(inp: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [rows, cols] or split bf16 [rows, 2*cols]."""
    assert inp.dtype == torch.float32 and inp.dim() == 2 and inp.stride(1) == 1 and out.is_contiguous()
    rows, cols = inp.shape
    kind = _act_kind(out, cols)
    _lib.check(
        _lib.load().duo_convert(_ptr(inp), inp.stride(0), _ptr(out), kind, rows, cols, _stream()), "duo_convert"
    )
    return out


_IN_KIND = {torch.bfloat16: _lib.ACT_BF16, torch.float16: _lib.ACT_F16, torch.float32: ACT_F32}


@_on_operand_device
def im2col3x3(x_nhwc: torch.Tensor, stride: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """NHWC [B,H,W,C] (bf16/f16/f32, contiguous) -> bf16 [B*Ho*Wo, 9*C] patches of a 3x3 / pad 1 / stride s conv."""
    assert x_nhwc.dim() == 4 and x_nhwc.is_contiguous()
    B, H, W, C = x_nhwc.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if out is None:
        out = torch.empty(B * Ho * Wo, 9 * C, dtype=torch.bfloat16, device=x_nhwc.device)
    assert out.is_contiguous() and out.shape == (B * Ho * Wo, 9 * C) and out.dtype == torch.bfloat16
    _lib.check(_lib.load().duo_im2col3x3(_ptr(x_nhwc), _IN_KIND[x_nhwc.dtype], _ptr(out), B, H, W, C, stride, _stream()),
               "duo_im2col3x3")
    return out


@_on_operand_device
def pool_to_slice(x_nhwc: torch.Tensor, out_slice: torch.Tensor, pool: int) -> torch.Tensor:
    """2x2 max-pool (pool=2) or copy (pool=1) of NHWC [B,H,W,C] into out_slice = wide[:, c0:c0+C]
    (a bf16 [B*Ho*Wo, C] column slice of a wider contiguous matrix)."""
    assert x_nhwc.dim() == 4 and x_nhwc.is_contiguous() and out_slice.dtype == torch.bfloat16
    B, H, W, C = x_nhwc.shape
    assert out_slice.shape == (B * (H // pool) * (W // pool), C) and out_slice.stride(1) == 1
    _lib.check(_lib.load().duo_pool_to_slice(_ptr(x_nhwc), _IN_KIND[x_nhwc.dtype], _ptr(out_slice), out_slice.stride(0),
                                             B, H, W, C, pool, _stream()), "duo_pool_to_slice")
    return out_slice


@_on_operand_device
def maxpool3x3s2(x: torch.Tensor) -> torch.Tensor:
    """nn.MaxPool2d(3, 2, 1) of a channels-last fp16 / bf16 NCHW tensor; returns a channels-last tensor."""
    assert x.dim() == 4 and x.dtype in (torch.float16, torch.bfloat16)
    assert x.is_contiguous(memory_format=torch.channels_last), "channels-last input expected"
    B, C, H, W = x.shape
    out = torch.empty(B, C, (H - 1) // 2 + 1, (W - 1) // 2 + 1, dtype=x.dtype, device=x.device,
                      memory_format=torch.channels_last)
    _lib.check(_lib.load().duo_maxpool3x3s2(_ptr(x), _IN_KIND[x.dtype], _ptr(out), B, H, W, C, _stream()),
               "duo_maxpool3x3s2")
    return out


@_on_operand_device
def conv2d(x_nhwc: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], ksize: int, stride: int,
           relu: bool, residual: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
           out_dtype: Optional[torch.dtype] = None, in2: Optional[torch.Tensor] = None, stride2: int = 1) -> torch.Tensor:
    """Implicit-GEMM convolution (tcgen05, TMA box loads; include/duoformer_sm100.h duo_conv2d).
    x_nhwc [B,H,W,Cin] fp16 / bf16 contiguous; weight [Cout, ksize*ksize*Cin] in (ky, kx, c) column order, same dtype;
    bias fp32 [Cout] or None; residual / out NHWC [B,Ho,Wo,Cout] fp16 / bf16 (out_dtype, default = the input's);
    padding ksize // 2.  in2 [B,H2,W2,Cin2] (+ stride2): fused 1x1 projection shortcut of another tensor, its weights are
    the last Cin2 columns of `weight` ([Cout, ksize*ksize*Cin + Cin2])."""
    assert x_nhwc.dim() == 4 and x_nhwc.is_contiguous() and x_nhwc.dtype in (torch.float16, torch.bfloat16)
    B, H, W, Cin = x_nhwc.shape
    Cout = weight.shape[0]
    Cin2 = 0
    if in2 is not None:
        assert in2.dim() == 4 and in2.is_contiguous() and in2.dtype == x_nhwc.dtype and in2.shape[0] == B
        Cin2 = in2.shape[3]
    assert weight.dtype == x_nhwc.dtype and weight.is_contiguous() and weight.shape == (Cout, ksize * ksize * Cin + Cin2)
    pad = ksize // 2
    Ho, Wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    if out is None:
        out = torch.empty(B, Ho, Wo, Cout, dtype=out_dtype or x_nhwc.dtype, device=x_nhwc.device)
    assert out.shape == (B, Ho, Wo, Cout) and out.is_contiguous() and out.dtype in (torch.float16, torch.bfloat16)
    if residual is not None:
        assert residual.shape == out.shape and residual.is_contiguous() and residual.dtype == out.dtype
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == Cout and bias.is_contiguous()
    a = _lib.Conv2dArgs()
    a.inp, a.weight, a.bias, a.residual, a.out = _ptr(x_nhwc), _ptr(weight), _ptr(bias), _ptr(residual), _ptr(out)
    a.B, a.H, a.W, a.Cin, a.Cout = B, H, W, Cin, Cout
    a.ksize, a.stride, a.relu, a.fp16 = ksize, stride, 1 if relu else 0, 1 if x_nhwc.dtype == torch.float16 else 0
    a.out_fp16 = 1 if out.dtype == torch.float16 else 0
    if in2 is not None:
        a.in2, a.H2, a.W2, a.Cin2, a.stride2 = _ptr(in2), in2.shape[1], in2.shape[2], Cin2, stride2
    e0 = _prof_begin()
    _lib.check(_lib.load().duo_conv2d(ctypes.byref(a), _stream()), "duo_conv2d")
    if e0 is not None:
        rows = B * Ho * Wo
        nbytes = x_nhwc.numel() * 2 + weight.numel() * 2 + rows * Cout * 2 * (2 if residual is not None else 1)
        nbytes += 0 if in2 is None else in2.numel() * 2 // (stride2 * stride2)
        tag = f"conv{ksize}x{ksize}s{stride}:{Cin}->{Cout}@{Ho}" + ("+res" if residual is not None else "") + (f"+ds{Cin2}" if in2 is not None else "")
        _prof_end(e0, "conv", 2.0 * rows * Cout * weight.shape[1], nbytes, tag)
    return out


@_on_operand_device
def stem_pack(x: torch.Tensor, scale: float, dtype: torch.dtype) -> torch.Tensor:
    """fp32 image [B,3,H,W] (any strides) * scale -> zero-padded row-pair tensor [B, H + 8, W + 8, 8]: channels 0..3 =
    pixel (R - 3, X - 3), channels 4..7 = pixel (R - 2, X - 3) (duo_stem_pack)."""
    assert x.dim() == 4 and x.shape[1] == 3 and x.dtype == torch.float32 and dtype in (torch.float16, torch.bfloat16)
    B, _, H, W = x.shape
    out = torch.empty(B, H + 8, W + 8, 8, dtype=dtype, device=x.device)
    _lib.check(_lib.load().duo_stem_pack(_ptr(x), x.stride(0), x.stride(1), x.stride(2), x.stride(3), float(scale), _ptr(out),
                                         1 if dtype == torch.float16 else 0, B, H, W, _stream()), "duo_stem_pack")
    return out


@_on_operand_device
def stem_conv7x7(packed: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], relu: bool = True) -> torch.Tensor:
    """7x7 / stride 2 / padding 3 stem convolution of a stem_pack()ed image; weight [Cout, 256] (pack_stem_weight)."""
    assert packed.dim() == 4 and packed.is_contiguous() and packed.shape[3] == 8
    B, Hp, Wp, _ = packed.shape
    H, W = Hp - 8, Wp - 8
    Cout = weight.shape[0]
    assert weight.shape == (Cout, 256) and weight.is_contiguous() and weight.dtype == packed.dtype
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == Cout and bias.is_contiguous()
    out = torch.empty(B, H // 2, W // 2, Cout, dtype=packed.dtype, device=packed.device)
    e0 = _prof_begin()
    _lib.check(_lib.load().duo_stem_conv7x7(_ptr(packed), _ptr(weight), _ptr(bias), _ptr(out), B, H, W, Cout, 1 if relu else 0,
                                            1 if packed.dtype == torch.float16 else 0, _stream()), "duo_stem_conv7x7")
    if e0 is not None:
        rows = B * (H // 2) * (W // 2)
        _prof_end(e0, "conv", 2.0 * rows * Cout * 256, packed.numel() * 2 + rows * Cout * 2, f"stem7x7:{Cout}@{H // 2}")
    return out


def pack_stem_weight(w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[Cout, 3, 7, 7] -> 16-bit [Cout, 256]: tap (ky, kx, c) at column (ky // 2) * 64 + kx * 8 + (ky % 2) * 4 + c
    (filter rows and columns padded to 8, channels to 4 with zeros) — the K layout of duo_stem_conv7x7."""
    w = w.detach().float()
    full = torch.zeros(w.shape[0], 8, 8, 4, dtype=torch.float32, device=w.device)  # (ky, kx, c)
    full[:, :7, :7, :3] = w.permute(0, 2, 3, 1)
    packed = full.view(w.shape[0], 4, 2, 8, 4).permute(0, 1, 3, 2, 4)  # (ky / 2, kx, ky % 2, c)
    return packed.reshape(w.shape[0], 256).to(dtype).contiguous()


def launch_count() -> int:
    return int(_lib.load().duo_launch_count())


def launch_count_reset() -> None:
    _lib.load().duo_launch_count_reset()


def split_weight(w: torch.Tensor) -> torch.Tensor:
    """fp32 [N,K] -> split bf16 [N,2K] (hi | lo).  Host-side weight packing helper (any device)."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.to(torch.float32)).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=1).contiguous()
