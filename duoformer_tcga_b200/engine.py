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

import contextlib
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import ops

PRECISIONS = ("bf16", "fp32")


@contextlib.contextmanager
def nvtx(name: str):
    """NVTX range around a stage of the forward (trunk / token builder / scale block i / patch stage / head): shows
    up on the timeline of nsys / ncu --nvtx; a host-side marker only (safe under CUDA-graph capture, ~100 ns idle)."""
    torch.cuda.nvtx.range_push("duo." + name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


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


# bf16 mode: LayerNorm statistics forwarding (duo_gemm's xb_out / stats_out / ln_stats).  Every residual GEMM
# (proj, fc2) also writes the bf16 copy of the updated rows and their per-row partial statistics; the GEMM that
# follows (fc1, next block's QKV) runs on the row-centred W * diag(ln_weight) and applies rstd in its epilogue, so no
# LayerNorm kernel re-reads the fp32 stream (only block 0's norm1 and the live rows of the last block run as
# standalone launches).  Precision study: tests/diag_ln_forwarding.py (same error as LayerNorm-then-round for
# |row mean| <= row spread; fp32 mode keeps the standalone LayerNorm).  Needs D % 256 == 0.
FORWARD_LN_STATS = True
# which LayerNorms are forwarded when FORWARD_LN_STATS is on: "norm1" (fc2 -> next block's QKV), "norm2" (proj -> fc1)
FORWARD_LINKS = ("norm1", "norm2")

# Token rows per chunk of the scale stage (engine.scale_stage).
SCALE_CHUNK_TOKENS = 1 << 21

STAT_COLS = 256  # columns per forwarded (mean, M2) pair (csrc/gemm_tcgen05.cu kStatCols)


def pack_ln_linear(weight: torch.Tensor, bias: Optional[torch.Tensor], ln_weight: torch.Tensor, ln_bias: torch.Tensor):
    """Linear(LayerNorm(x)) with the norm folded into the operands (consumer side of the statistics forwarding).

    W'' = W * ln_weight with every row centred: sum_k W''[n, k] = 0 makes the row mean of x cancel inside the
    tensor-core product (x W''^T = (x - mean) W'^T), so the GEMM epilogue only scales by rstd.  The centring must hold
    for the bf16 values the MMA actually multiplies: after rounding, the residual row sum (~sqrt(K) ulps) is absorbed
    by the row's smallest element, whose fine ulp leaves |sum_k W''[n, k]| ~ 1e-6 of a typical weight — the leftover
    mean leakage is then < 1e-5 of the output scale even for |mean| = 10 x the row spread.
    bias' = W ln_bias + b (fp32, exact W)."""
    w = weight.detach().to(torch.float32)
    wg = w * ln_weight.detach().to(torch.float32)[None, :]
    wg = wg - wg.mean(dim=1, keepdim=True)
    wp = wg.to(torch.bfloat16)
    for _ in range(2):
        resid = wp.to(torch.float64).sum(dim=1)  # exact row sums of the rounded values
        k = wp.abs().argmin(dim=1, keepdim=True)
        fixed = (wp.gather(1, k).to(torch.float64) - resid[:, None]).to(torch.bfloat16)
        wp.scatter_(1, k, fixed)
    b = w @ ln_bias.detach().to(torch.float32)
    if bias is not None:
        b = b + bias.detach().to(torch.float32)
    return wp.contiguous(), b.contiguous()


def pack_scale_block(precision: str, n1w, n1b, n2w, n2b, qkv, proj, fc1, fc2, g1=None, g2=None) -> Dict:
    """Packed device operands of one scale block ((weight, bias) pairs of the four Linears + the two norms)."""
    d = {
        "n1w": _f32(n1w), "n1b": _f32(n1b), "n2w": _f32(n2w), "n2b": _f32(n2b),
        "qkv": pack_linear(qkv[0], qkv[1], precision), "proj": pack_linear(proj[0], proj[1], precision),
        "fc1": pack_linear(fc1[0], fc1[1], precision), "fc2": pack_linear(fc2[0], fc2[1], precision),
        "g1": _f32(g1), "g2": _f32(g2),
    }
    if precision == "bf16":
        d["qkv_ln"] = pack_ln_linear(qkv[0], qkv[1], n1w, n1b)
        d["fc1_ln"] = pack_ln_linear(fc1[0], fc1[1], n2w, n2b)
    return d


def _scale_chunk(X, g0, ng, buf, blocks, num_heads, scale, eps, precision, capture, attn_algo, live_only_last):
    """All scale blocks for the patches (groups of S token rows) [g0, g0+ng) of the flattened batch.

    bf16 mode with statistics forwarding, per block (5 launches): QKV GEMM (LayerNorm applied in its epilogue),
    attention, proj GEMM (+residual, emits bf16 rows + statistics), fc1 GEMM (LayerNorm in the epilogue, +GELU),
    fc2 GEMM (+residual, emits the next block's bf16 rows + statistics).  fp32 mode (and D % 256 != 0): 7 launches
    with standalone LayerNorm kernels (split hi|lo outputs in fp32 mode)."""
    B, P, S, D = X.shape
    fp32 = precision == "fp32"
    fwd = FORWARD_LN_STATS and not fp32 and D % 256 == 0 and "qkv_ln" in blocks[0]
    fwd1, fwd2 = fwd and "norm1" in FORWARD_LINKS, fwd and "norm2" in FORWARD_LINKS
    kd = 2 if fp32 else 1
    hidden = blocks[0]["fc1"][0].shape[0]
    T = ng * S
    hn_bytes = _align(T * kd * D * 2)
    Xc = X.view(B * P * S, D)[g0 * S : (g0 + ng) * S]
    Ha = Workspace.view(buf, 0, (T, kd * D), torch.bfloat16)             # LayerNorm output / bf16 copy of the stream
    Hb = Workspace.view(buf, hn_bytes, (T, kd * D), torch.bfloat16)      # attention output
    QKV = Workspace.view(buf, 2 * hn_bytes, (T, 3 * D), torch.float32 if fp32 else torch.bfloat16)
    HID = Workspace.view(buf, 2 * hn_bytes, (T, kd * hidden), torch.bfloat16)  # aliases QKV (dead after attention)
    big = _align(max(T * 3 * D * (4 if fp32 else 2), T * kd * hidden * 2))
    # two statistics buffers, written alternately: a forwarding GEMM writes the new statistics into one while its CTAs
    # still read the rows' previous ones (the shift of the bf16 copy) from the other
    st_bytes = _align(T * (D // STAT_COLS) * 8)
    STS = [Workspace.view(buf, 2 * hn_bytes + big + k * st_bytes, (T, D // STAT_COLS, 2), torch.float32) for k in (0, 1)] if fwd else None
    st_cur = [0]  # index of the buffer holding the most recent statistics of Xc

    def forward_residual(A, blk, key, gamma_key):
        """Xc += gamma * (A W^T + b) with statistics forwarding: bf16(Xc - previous row mean) -> Ha, new statistics."""
        prev, nxt = STS[st_cur[0]], STS[st_cur[0] ^ 1]
        ops.gemm(A, blk[key][0], blk[key][1], Xc, ops.EPI_RESIDUAL_F32, gamma=blk[gamma_key], xb_out=Ha, stats_out=nxt,
                 shift_stats=prev)
        st_cur[0] ^= 1
        return nxt

    gelu = ops.EPI_GELU_SPLIT_BF16 if fp32 else ops.EPI_GELU_BF16
    L = len(blocks)

    def run_block(i: int, blk: Dict, have_ln1: bool) -> bool:
        """One scale block; have_ln1: Ha / ST hold this block's forwarded norm1 input.  Returns the same for the next."""
        last = i == L - 1
        if have_ln1 and live_only_last and last:
            # Last scale block: keys and values for every token, but only the s = 0 QUERY of each patch is ever used
            # (below) — the q third of the projection runs on those rows alone (A, output and statistics are the strided
            # s = 0 views; the statistics are gathered densely first: the kernel indexes them by row)
            w, b = blk["qkv_ln"]
            st = STS[st_cur[0]]
            ops.gemm(Ha, w[D:], b[D:], QKV[:, D:], ops.EPI_BF16, ln_stats=st, ln_eps=eps)
            st0 = st.view(ng, S, D // STAT_COLS, 2)[:, 0].contiguous()
            ops.gemm(Ha.view(ng, S, D)[:, 0, :], w[:D], b[:D], QKV.view(ng, S, 3 * D)[:, 0, :D], ops.EPI_BF16, ln_stats=st0, ln_eps=eps)
        elif have_ln1:
            w, b = blk["qkv_ln"]
            ops.gemm(Ha, w, b, QKV, ops.EPI_BF16, ln_stats=STS[st_cur[0]], ln_eps=eps)
        else:
            # (in forwarding mode the LayerNorm launch also leaves the row statistics: the next producer's shift)
            ops.layernorm(Xc, blk["n1w"], blk["n1b"], Ha, eps, stats_out=STS[st_cur[0]] if fwd else None)
            ops.gemm(Ha, blk["qkv"][0], blk["qkv"][1], QKV, ops.EPI_F32 if fp32 else ops.EPI_BF16, split3=fp32)
        if live_only_last and last:
            # Last scale block: only the scale token (s = 0) of every patch is consumed downstream
            # (scale_attention.py:183-185), so K/V are needed for all tokens but the query,
            # attention output, proj, MLP and both residual updates only for the s = 0 rows
            # (SURVEY.md App. A.3).  X0 is the strided view of those rows inside X.
            R = ng
            X0 = Xc.view(R, S, D)[:, 0, :]
            A0 = Workspace.view(buf, hn_bytes, (R, kd * D), torch.bfloat16)
            N0 = Workspace.view(buf, 0, (R, kd * D), torch.bfloat16)
            ops.group_attention(QKV, A0, S, num_heads, scale, algo=attn_algo, q_rows=1)
            ops.gemm(A0, blk["proj"][0], blk["proj"][1], X0, ops.EPI_RESIDUAL_F32, gamma=blk["g1"], split3=fp32)
            ops.layernorm(X0, blk["n2w"], blk["n2b"], N0, eps)
            H0 = Workspace.view(buf, 2 * hn_bytes, (R, kd * hidden), torch.bfloat16)
            ops.gemm(N0, blk["fc1"][0], blk["fc1"][1], H0, gelu, split3=fp32)
            ops.gemm(H0, blk["fc2"][0], blk["fc2"][1], X0, ops.EPI_RESIDUAL_F32, gamma=blk["g2"], split3=fp32)
            if capture is not None:
                capture[f"scale_block_{i}_s0"] = X[:, :, 0, :].clone()
            return False
        ops.group_attention(QKV, Hb, S, num_heads, scale, algo=attn_algo)
        if fwd2:
            st = forward_residual(Hb, blk, "proj", "g1")
            w, b = blk["fc1_ln"]
            ops.gemm(Ha, w, b, HID, gelu, ln_stats=st, ln_eps=eps)
        else:
            ops.gemm(Hb, blk["proj"][0], blk["proj"][1], Xc, ops.EPI_RESIDUAL_F32, gamma=blk["g1"], split3=fp32)
            ops.layernorm(Xc, blk["n2w"], blk["n2b"], Ha, eps, stats_out=STS[st_cur[0]] if fwd else None)
            ops.gemm(Ha, blk["fc1"][0], blk["fc1"][1], HID, gelu, split3=fp32)
        forward_next = fwd1 and not last
        if forward_next:
            forward_residual(HID, blk, "fc2", "g2")
        else:
            ops.gemm(HID, blk["fc2"][0], blk["fc2"][1], Xc, ops.EPI_RESIDUAL_F32, gamma=blk["g2"], split3=fp32)
        if capture is not None:
            capture[f"scale_block_{i}"] = X.clone()
        return forward_next

    have_ln1 = False
    for i, blk in enumerate(blocks):
        with nvtx(f"scale_block_{i}"):
            have_ln1 = run_block(i, blk, have_ln1)


def _chunk_workspace_bytes(ng, S, D, hidden, fp32):
    T = ng * S
    kd = 2 if fp32 else 1
    stats = 0 if fp32 else 2 * _align(T * (D // STAT_COLS) * 8)
    return 2 * _align(T * kd * D * 2) + _align(max(T * 3 * D * (4 if fp32 else 2), T * kd * hidden * 2)) + stats


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
    max_chunk_tokens: Optional[int] = None,
    live_only_last: bool = False,
) -> torch.Tensor:
    """L x { X += g1*Attn(LN1 X) ; X += g2*MLP(LN2 X) } in place on the fp32 token tensor
    X [B, P, S, D]  (scale_attention.py:90-93 / multiscale_attn.py:282-285).

    The batch is walked in chunks of whole patches (all blocks for one chunk, then the next chunk; every token row
    only interacts with the S rows of its own patch): max_chunk_tokens (default SCALE_CHUNK_TOKENS) bounds the
    workspace and decides whether the activations a kernel hands to the next one are still in L2.
    live_only_last: whole-model callers only — the LAST block updates just the s = 0 row of every patch
    (all that MultiscaleFormer / MultiscaleTransformer consume afterwards); the other rows keep their
    pre-block values."""
    B, P, S, D = X.shape
    if not blocks:
        return X
    fp32 = precision == "fp32"
    hidden = blocks[0]["fc1"][0].shape[0]
    groups = B * P
    chunk_groups = max(1, min(groups, (max_chunk_tokens or SCALE_CHUNK_TOKENS) // S))
    if capture is not None:
        chunk_groups = groups  # captures want whole-batch tensors after every block
    buf = ws.get(_chunk_workspace_bytes(chunk_groups, S, D, hidden, fp32))
    for g0 in range(0, groups, chunk_groups):
        _scale_chunk(X, g0, min(chunk_groups, groups - g0), buf, blocks, num_heads, scale, eps, precision, capture,
                     attn_algo, live_only_last)
    return X


class PatchScratch:
    """Activation buffers of the patch stage carved out of the model's workspace (the scale stage's buffers are dead
    by then): two ping-pong token buffers, qkv, attention output and the fp32 result of the last block — the
    2 x depth launches of the patch stage allocate nothing."""

    def __init__(self, ws: Workspace, rows: int, D: int):
        self.sizes = {"z0": _align(rows * 2 * D * 2), "z1": _align(rows * 2 * D * 2), "qkv": _align(rows * 6 * D * 2),
                      "ao": _align(rows * 2 * D * 2), "f32": _align(rows * D * 4)}
        self.offsets, off = {}, 0
        for k, n in self.sizes.items():
            self.offsets[k] = off
            off += n
        self.buf = ws.get(off)
        self._next_z = 0

    def view(self, name: str, shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        return Workspace.view(self.buf, self.offsets[name], shape, dtype)

    def next_z(self, shape: Tuple[int, ...], dtype: torch.dtype) -> torch.Tensor:
        """The token buffer the current block's input does NOT live in."""
        t = self.view(f"z{self._next_z}", shape, dtype)
        self._next_z ^= 1
        return t


def region_attention(
    Z: torch.Tensor,
    blk: Dict,
    N: int,
    num_heads: int,
    scale: float,
    precision: str,
    out_f32: bool,
    cls_only: bool = False,
    scratch: Optional[PatchScratch] = None,
    skip_proj: bool = False,
) -> torch.Tensor:
    """One patch ("region") attention block without residual / norm / MLP:
    Z <- proj(softmax(q k^T * scale) v)   (scale_attention.py:195-209, multiscale_attn.py:205-219).
    skip_proj: return the attention output (the kind proj would consume) — the caller has composed this block's proj
    with the next block's qkv into one linear map (MultiscaleFormer.fuse_patch_linears).

    precision "bf16":  Z bf16 [B*N, D], everything bf16.
              "fp32":  Z split [B*N, 2D]; 3-pass split GEMMs; attention in split precision on tcgen05 (N <= 64:
                       q, k, v as hi | lo pairs, three UMMAs per product) or fp32 qkv + fp32 FMA attention
                       (N > 64, and the CLS-only query of the last block).
              "mixed": Z split [B*N, 2D] (the tensor handed from block to block keeps ~16 mantissa
                       bits, weights are split too), but qkv and the attention output are bf16 so the
                       attention runs on the mma.sync kernel; proj multiplies the exact bf16
                       attention output with the split weight (2-pass).
    scratch: buffers for qkv / attention output / result (Z itself must be scratch.next_z of the previous block or a
             foreign tensor); None allocates.
    Returns the same kind as Z, or fp32 [rows, D] if out_f32."""
    split_io = precision in ("fp32", "mixed")
    fp32 = precision == "fp32"
    kd_io = 2 if split_io else 1
    kd_ao = 2 if fp32 else 1
    rows = Z.shape[0]
    D = Z.shape[1] // kd_io
    dev = Z.device

    def buf(name, shape, dtype):
        if scratch is None:
            return torch.empty(shape, dtype=dtype, device=dev)
        return scratch.next_z(shape, dtype) if name == "z" else scratch.view(name, shape, dtype)

    if fp32 and not cls_only and N <= 64:
        # global attention on tcgen05 / TMEM in split precision: q, k, v stay hi | lo bf16 pairs end to end
        QKV = buf("qkv", (rows, 6 * D), torch.bfloat16)
        ops.gemm(Z, blk["qkv"][0], blk["qkv"][1], QKV, ops.EPI_SPLIT_BF16, split3=1)
        AO = buf("ao", (rows, 2 * D), torch.bfloat16)
        ops.group_attention(QKV, AO, N, num_heads, scale, split_in=True)
    else:
        QKV = buf("qkv", (rows, 3 * D), torch.float32 if fp32 else torch.bfloat16)
        ops.gemm(Z, blk["qkv"][0], blk["qkv"][1], QKV, ops.EPI_F32 if fp32 else ops.EPI_BF16, split3=1 if split_io else 0)
        if cls_only:
            # last patch block: only the CLS query row of every image reaches the head (scale_attention.py:341)
            rows = rows // N
            AO = buf("ao", (rows, kd_ao * D), torch.bfloat16)
            ops.group_attention(QKV, AO, N, num_heads, scale, q_rows=1)
        else:
            AO = buf("ao", (rows, kd_ao * D), torch.bfloat16)
            ops.group_attention(QKV, AO, N, num_heads, scale)
    if skip_proj:
        return AO
    proj_split = 1 if fp32 else (2 if precision == "mixed" else 0)
    if out_f32:
        out = buf("f32", (rows, D), torch.float32)
        ops.gemm(AO, blk["proj"][0], blk["proj"][1], out, ops.EPI_F32, split3=proj_split)
    else:
        out = buf("z", (rows, kd_io * D), torch.bfloat16)
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
