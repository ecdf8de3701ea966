"""Host-side executor of the DuoFormer forward path: weight packing and kernel sequencing.

This is orchestration only — every arithmetic step is a kernel of libduoformer_sm100.so
(see ops.py).  Two precisions:

  "bf16"  (default, the benchmarked path) bf16 GEMM operands, fp32 accumulation / residual
          stream / LayerNorm statistics / softmax.
  "fp32"  fp32-accuracy mode (north-star tolerance 1e-3): activations and weights are split into
          bf16 hi|lo halves and every Linear runs as a 3-pass tensor-core GEMM (hi*hi+hi*lo+lo*hi).

Stage-wise parity tests hook `capture` dicts (name -> tensor clone).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops

PRECISIONS = ("bf16", "fp32")


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise NotImplementedError(
            f"{what}: the B200 path has no CPU / eager fallback — move the model and inputs to a CUDA device"
        )


def param_signature(module: nn.Module, precision: str) -> tuple:
    sig: List = [precision]
    for p in list(module.parameters()) + list(module.buffers()):
        sig.append((p.data_ptr(), p._version))
    return tuple(sig)


def pack_linear(weight: torch.Tensor, bias: Optional[torch.Tensor], precision: str,
                dtype: torch.dtype = torch.bfloat16) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """[N, K] fp32 master weight -> bf16 (or fp16, or split hi|lo bf16) K-major TMA operand + fp32 bias."""
    w = weight.detach().reshape(weight.shape[0], -1).to(torch.float32)
    wp = ops.split_weight(w) if precision == "fp32" else w.to(dtype).contiguous()
    b = None if bias is None else bias.detach().to(torch.float32).contiguous()
    return wp, b


def _f32(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    return None if t is None else t.detach().to(torch.float32).contiguous()


class Workspace:
    """One reusable byte buffer carved into the per-chunk activation tensors."""

    def __init__(self, device: torch.device):
        self.device = device
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self.buf

    @staticmethod
    def view(buf: torch.Tensor, offset: int, shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return buf[offset : offset + nbytes].view(dtype).view(*shape)


def _align(n: int, a: int = 1024) -> int:
    return (n + a - 1) // a * a


# bf16 mode: run each LayerNorm inside the preceding residual GEMM (duo_gemm's ln_out: extra LayerNorm
# warps re-read every finished 256-row panel and write the bf16 operand of the next GEMM).
# Correct (tests/test_kernels_gpu.py, parity suite with the flag on) but OFF by default: on the
# power-capped B200 the LayerNorm's traffic costs the same time inside the GEMM as beside it —
# per 64 images fc2+LN1 1.13 ms fused vs 1.15 ms as two launches, proj+LN2 0.60 vs 0.55 ms
# (profiles/r01_notes.md; the first version, LN in the epilogue warps, was 2x slower still).
FUSE_LAYERNORM = False


_LN_SYNC = {}


def _ln_sync(device, rows, slot):
    """Zeroed row-panel counters for duo_gemm's fused LayerNorm (8 per 256 rows).  Every launch leaves
    them zero, so one buffer per (device, concurrently running chunk) is allocated once and reused."""
    need = 8 * ((rows + 255) // 256)
    key = (device.index, slot)
    t = _LN_SYNC.get(key)
    if t is None or t.numel() < need:
        t = torch.zeros(need, dtype=torch.int32, device=device)
        _LN_SYNC[key] = t
    return t


def _scale_chunk_ops(X, b0, nb, buf, blocks, num_heads, scale, eps, precision, capture, attn_algo, live_only_last):
    """Generator over the kernel launches of all scale blocks for images [b0, b0+nb): yields after
    every launch so that two chunks can be issued interleaved on two streams (see scale_stage).

    bf16 mode, per block (5 launches): QKV GEMM, attention, proj GEMM (+residual +LayerNorm2),
    fc1 GEMM (+GELU), fc2 GEMM (+residual +LayerNorm1 of the NEXT block); only block 0 needs a
    standalone LayerNorm launch.  fp32 mode keeps the LayerNorm launches (split hi|lo outputs)."""
    B, P, S, D = X.shape
    fp32 = precision == "fp32"
    fuse = FUSE_LAYERNORM and not fp32
    kd = 2 if fp32 else 1
    hidden = blocks[0]["fc1"][0].shape[0]
    T = nb * P * S
    hn_bytes = _align(T * kd * D * 2)
    Xc = X[b0 : b0 + nb].view(T, D)
    Ha = Workspace.view(buf, 0, (T, kd * D), torch.bfloat16)             # LayerNorm outputs
    Hb = Workspace.view(buf, hn_bytes, (T, kd * D), torch.bfloat16)      # attention output
    QKV = Workspace.view(buf, 2 * hn_bytes, (T, 3 * D), torch.float32 if fp32 else torch.bfloat16)
    HID = Workspace.view(buf, 2 * hn_bytes, (T, kd * hidden), torch.bfloat16)  # aliases QKV (dead after attention)
    gelu = ops.EPI_GELU_SPLIT_BF16 if fp32 else ops.EPI_GELU_BF16
    L = len(blocks)
    sync = _ln_sync(X.device, T, b0) if fuse else None
    for i, blk in enumerate(blocks):
        last = i == L - 1
        if i == 0 or not fuse:
            ops.layernorm(Xc, blk["n1w"], blk["n1b"], Ha, eps)
            yield
        ops.gemm(Ha, blk["qkv"][0], blk["qkv"][1], QKV, ops.EPI_F32 if fp32 else ops.EPI_BF16, split3=fp32)
        yield
        if live_only_last and last:
            # Last scale block: only the scale token (s = 0) of every patch is consumed downstream
            # (scale_attention.py:183-185), so K/V are needed for all tokens but the query,
            # attention output, proj, MLP and both residual updates only for the s = 0 rows
            # (SURVEY.md App. A.3).  X0 is the strided view of those rows inside X.
            R = nb * P
            X0 = Xc.view(R, S, D)[:, 0, :]
            A0 = Workspace.view(buf, hn_bytes, (R, kd * D), torch.bfloat16)
            N0 = Workspace.view(buf, 0, (R, kd * D), torch.bfloat16)
            ops.group_attention(QKV, A0, S, num_heads, scale, algo=attn_algo, q_rows=1)
            yield
            if fuse:
                ops.gemm(A0, blk["proj"][0], blk["proj"][1], X0, ops.EPI_RESIDUAL_F32, gamma=blk["g1"],
                         ln_gamma=blk["n2w"], ln_beta=blk["n2b"], ln_out=N0, ln_eps=eps, ln_sync=sync)
                yield
            else:
                ops.gemm(A0, blk["proj"][0], blk["proj"][1], X0, ops.EPI_RESIDUAL_F32, gamma=blk["g1"], split3=fp32)
                yield
                ops.layernorm(X0, blk["n2w"], blk["n2b"], N0, eps)
                yield
            H0 = Workspace.view(buf, 2 * hn_bytes, (R, kd * hidden), torch.bfloat16)
            ops.gemm(N0, blk["fc1"][0], blk["fc1"][1], H0, gelu, split3=fp32)
            yield
            ops.gemm(H0, blk["fc2"][0], blk["fc2"][1], X0, ops.EPI_RESIDUAL_F32, gamma=blk["g2"], split3=fp32)
            yield
            if capture is not None:
                capture[f"scale_block_{i}_s0"] = X[:, :, 0, :].clone()
            continue
        ops.group_attention(QKV, Hb, S, num_heads, scale, algo=attn_algo)
        yield
        if fuse:
            ops.gemm(Hb, blk["proj"][0], blk["proj"][1], Xc, ops.EPI_RESIDUAL_F32, gamma=blk["g1"],
                     ln_gamma=blk["n2w"], ln_beta=blk["n2b"], ln_out=Ha, ln_eps=eps, ln_sync=sync)
            yield
        else:
            ops.gemm(Hb, blk["proj"][0], blk["proj"][1], Xc, ops.EPI_RESIDUAL_F32, gamma=blk["g1"], split3=fp32)
            yield
            ops.layernorm(Xc, blk["n2w"], blk["n2b"], Ha, eps)
            yield
        ops.gemm(Ha, blk["fc1"][0], blk["fc1"][1], HID, gelu, split3=fp32)
        yield
        if fuse and not last:
            nxt = blocks[i + 1]
            ops.gemm(HID, blk["fc2"][0], blk["fc2"][1], Xc, ops.EPI_RESIDUAL_F32, gamma=blk["g2"],
                     ln_gamma=nxt["n1w"], ln_beta=nxt["n1b"], ln_out=Ha, ln_eps=eps, ln_sync=sync)
        else:
            ops.gemm(HID, blk["fc2"][0], blk["fc2"][1], Xc, ops.EPI_RESIDUAL_F32, gamma=blk["g2"], split3=fp32)
        yield
        if capture is not None:
            capture[f"scale_block_{i}"] = X.clone()


def _chunk_workspace_bytes(nb, P, S, D, hidden, fp32):
    T = nb * P * S
    kd = 2 if fp32 else 1
    return 2 * _align(T * kd * D * 2) + _align(max(T * 3 * D * (4 if fp32 else 2), T * kd * hidden * 2))


# Optional two-lane issue: the HBM-bound kernels (LayerNorm) of one half of the batch overlap the
# tensor-bound GEMMs of the other half (LN CTAs need no shared memory, so they co-reside with the
# persistent GEMM CTAs).  Lane 1 trails lane 0 by _LANE_OFFSET launches.  Measured on a
# power-capped B200 (profiles/r01_notes.md): 1 217 vs 1 210 images/s — within noise, because the
# step is limited by the 1 kW cap rather than by idle pipes — so it is OFF by default.
OVERLAP_LANES = False
_LANE_OFFSET = 1
_side_streams: Dict[int, Tuple[torch.cuda.Stream, torch.cuda.Stream]] = {}


def scale_stage(
    X: torch.Tensor,
    blocks: List[Dict],
    num_heads: int,
    scale: float,
    eps: float,
    precision: str,
    ws: Workspace,
    capture: Optional[Dict[str, torch.Tensor]] = None,
    attn_algo: int = 0,
    max_chunk_tokens: int = 1 << 21,
    live_only_last: bool = True,
) -> torch.Tensor:
    """L x { X += g1*Attn(LN1 X) ; X += g2*MLP(LN2 X) } in place on the fp32 token tensor
    X [B, P, S, D]  (scale_attention.py:90-93 / multiscale_attn.py:282-285)."""
    B, P, S, D = X.shape
    if not blocks:
        return X
    fp32 = precision == "fp32"
    hidden = blocks[0]["fc1"][0].shape[0]
    tokens_per_image = P * S
    chunk_images = max(1, min(B, max_chunk_tokens // tokens_per_image))
    if capture is not None:
        chunk_images = B  # captures want whole-batch tensors after every block
    two_lanes = OVERLAP_LANES and capture is None and B >= 2 and B * tokens_per_image >= (1 << 13)
    if two_lanes:
        chunk_images = max(1, min(chunk_images, (B + 1) // 2))
    chunks = [(b0, min(chunk_images, B - b0)) for b0 in range(0, B, chunk_images)]
    per_chunk = _chunk_workspace_bytes(chunk_images, P, S, D, hidden, fp32)
    args = (blocks, num_heads, scale, eps, precision, capture, attn_algo, live_only_last)
    if not two_lanes:
        buf = ws.get(per_chunk)
        for b0, nb in chunks:
            for _ in _scale_chunk_ops(X, b0, nb, buf, *args):
                pass
        return X

    buf = ws.get(2 * per_chunk)
    bufs = (buf[:per_chunk], buf[per_chunk:])
    dev = X.device.index if X.device.index is not None else torch.cuda.current_device()
    if dev not in _side_streams:
        _side_streams[dev] = (torch.cuda.Stream(device=X.device), torch.cuda.Stream(device=X.device))
    lanes = _side_streams[dev]
    main = torch.cuda.current_stream(X.device)
    start = torch.cuda.Event()
    start.record(main)
    for s in lanes:
        s.wait_event(start)
    for c in range(0, len(chunks), 2):
        gens = [_scale_chunk_ops(X, chunks[c][0], chunks[c][1], bufs[0], *args)]
        if c + 1 < len(chunks):
            gens.append(_scale_chunk_ops(X, chunks[c + 1][0], chunks[c + 1][1], bufs[1], *args))
        alive = [True] * len(gens)

        def step(k):
            if alive[k]:
                with torch.cuda.stream(lanes[k]):
                    try:
                        next(gens[k])
                    except StopIteration:
                        alive[k] = False

        for _ in range(_LANE_OFFSET):
            step(0)
        while any(alive):
            if len(gens) > 1:
                step(1)
            step(0)
    for s in lanes:
        done = torch.cuda.Event()
        done.record(s)
        main.wait_event(done)
    return X


def region_attention(
    Z: torch.Tensor,
    blk: Dict,
    N: int,
    num_heads: int,
    scale: float,
    precision: str,
    out_f32: bool,
    cls_only: bool = False,
) -> torch.Tensor:
    """One patch ("region") attention block without residual / norm / MLP:
    Z <- proj(softmax(q k^T * scale) v)   (scale_attention.py:195-209, multiscale_attn.py:205-219).

    precision "bf16":  Z bf16 [B*N, D], everything bf16.
              "fp32":  Z split [B*N, 2D]; 3-pass split GEMMs; attention in split precision on tcgen05 (N <= 64:
                       q, k, v as hi | lo pairs, three UMMAs per product) or fp32 qkv + fp32 FMA attention
                       (N > 64, and the CLS-only query of the last block).
              "mixed": Z split [B*N, 2D] (the tensor handed from block to block keeps ~16 mantissa
                       bits, weights are split too), but qkv and the attention output are bf16 so the
                       attention runs on the mma.sync kernel; proj multiplies the exact bf16
                       attention output with the split weight (2-pass).
    Returns the same kind as Z, or fp32 [rows, D] if out_f32."""
    split_io = precision in ("fp32", "mixed")
    fp32 = precision == "fp32"
    kd_io = 2 if split_io else 1
    kd_ao = 2 if fp32 else 1
    rows = Z.shape[0]
    D = Z.shape[1] // kd_io
    dev = Z.device
    if fp32 and not cls_only and N <= 64:
        # global attention on tcgen05 / TMEM in split precision: q, k, v stay hi | lo bf16 pairs end to end
        QKV = torch.empty(rows, 6 * D, dtype=torch.bfloat16, device=dev)
        ops.gemm(Z, blk["qkv"][0], blk["qkv"][1], QKV, ops.EPI_SPLIT_BF16, split3=1)
        AO = torch.empty(rows, 2 * D, dtype=torch.bfloat16, device=dev)
        ops.group_attention(QKV, AO, N, num_heads, scale, split_in=True)
    else:
        QKV = torch.empty(rows, 3 * D, dtype=torch.float32 if fp32 else torch.bfloat16, device=dev)
        ops.gemm(Z, blk["qkv"][0], blk["qkv"][1], QKV, ops.EPI_F32 if fp32 else ops.EPI_BF16, split3=1 if split_io else 0)
        AO = None
    if AO is not None:
        pass
    elif cls_only:
        # last patch block: only the CLS query row of every image reaches the head (scale_attention.py:341)
        rows = rows // N
        AO = torch.empty(rows, kd_ao * D, dtype=torch.bfloat16, device=dev)
        ops.group_attention(QKV, AO, N, num_heads, scale, q_rows=1)
    else:
        AO = torch.empty(rows, kd_ao * D, dtype=torch.bfloat16, device=dev)
        ops.group_attention(QKV, AO, N, num_heads, scale)
    proj_split = 1 if fp32 else (2 if precision == "mixed" else 0)
    if out_f32:
        out = torch.empty(rows, D, dtype=torch.float32, device=dev)
        ops.gemm(AO, blk["proj"][0], blk["proj"][1], out, ops.EPI_F32, split3=proj_split)
    else:
        out = torch.empty(rows, kd_io * D, dtype=torch.bfloat16, device=dev)
        ops.gemm(AO, blk["proj"][0], blk["proj"][1], out, ops.EPI_SPLIT_BF16 if split_io else ops.EPI_BF16,
                 split3=proj_split)
    return out


def unsplit(t: torch.Tensor, precision: str) -> torch.Tensor:
    """Debug/capture helper: fp32 view of a bf16 / split-bf16 activation."""
    if precision in ("fp32", "mixed") and t.dtype == torch.bfloat16:
        D = t.shape[-1] // 2
        return t[..., :D].float() + t[..., D:].float()
    return t.float()


class PackCache:
    """Mixin: lazily (re)build packed device weights when parameters or precision change."""

    _packed = None
    _packed_sig = None

    def packed(self, build: Callable[[], Dict], module: nn.Module, precision: str) -> Dict:
        sig = param_signature(module, precision)
        if self._packed is None or self._packed_sig != sig:
            with torch.no_grad():
                self._packed = build()
            self._packed_sig = sig
        return self._packed
