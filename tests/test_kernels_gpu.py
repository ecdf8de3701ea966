"""Kernel-level parity (-m gpu): every C-ABI entry point against a plain PyTorch fp32 restatement
of the same op on seeded inputs.  Tolerances: bf16 operands -> error of the fp32 result after
rounding the INPUTS to bf16 is compared tightly (1e-2 rel of max for bf16 outputs, 2e-3 for fp32
outputs); split3 (fp32-accuracy) GEMM: 2e-5."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from duoformer_tcga_b200 import ops  # noqa: E402


def relerr(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-30)).item()


def _gen(shape, seed, scale=1.0, device="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(device)


@pytest.mark.parametrize(
    "M,N,K",
    [(128, 256, 64), (128, 128, 64), (256, 768, 768), (294 * 2, 2304, 768), (1000, 768, 3072), (128 * 300, 768, 256),
     (98, 768, 2048), (256 * 75 + 77, 2304, 768), (256 * 160 + 129, 768, 3072)],
)
def test_gemm_bf16_epilogues(M, N, K):
    A = _gen((M, K), 1).to(torch.bfloat16)
    W = _gen((N, K), 2, 0.05).to(torch.bfloat16)
    bias = _gen((N,), 3)
    ref = A.float() @ W.float().t() + bias
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, W, bias, out, ops.EPI_BF16)
    assert relerr(out, ref) < 1e-2
    out32 = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(A, W, bias, out32, ops.EPI_F32)
    assert relerr(out32, ref) < 1e-4
    ops.gemm(A, W, None, out32, ops.EPI_F32)
    assert relerr(out32, ref - bias) < 1e-4
    outg = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(A, W, bias, outg, ops.EPI_GELU_BF16)
    assert relerr(outg, torch.nn.functional.gelu(ref)) < 1e-2
    # residual with LayerScale
    X = _gen((M, N), 4, 3.0)
    gamma = _gen((N,), 5)
    Xr = X + gamma * ref
    ops.gemm(A, W, bias, X, ops.EPI_RESIDUAL_F32, gamma=gamma)
    assert relerr(X, Xr) < 1e-4
    X2 = _gen((M, N), 6, 3.0)
    X2r = X2 + ref
    ops.gemm(A, W, bias, X2, ops.EPI_RESIDUAL_F32)
    assert relerr(X2, X2r) < 1e-4


@pytest.mark.parametrize("M", [300, 256 * 150 + 5])
def test_gemm_split3_fp32_accuracy(M):
    N, K = 768, 768
    A = _gen((M, K), 11)
    W = _gen((N, K), 12, 0.05)
    bias = _gen((N,), 13)
    ref = (A.double() @ W.double().t() + bias.double()).float()
    As = torch.empty(M, 2 * K, dtype=torch.bfloat16, device="cuda")
    ops.convert(A, As)
    assert torch.equal(As, ops.split_weight(A))
    Ws = ops.split_weight(W)
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(As, Ws, bias, out, ops.EPI_F32, split3=True)
    assert relerr(out, ref) < 2e-5
    outs = torch.empty(M, 2 * N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(As, Ws, bias, outs, ops.EPI_SPLIT_BF16, split3=True)
    rec = outs[:, :N].float() + outs[:, N:].float()
    assert relerr(rec, ref) < 3e-5
    ops.gemm(As, Ws, bias, outs, ops.EPI_GELU_SPLIT_BF16, split3=True)
    rec = outs[:, :N].float() + outs[:, N:].float()
    assert relerr(rec, torch.nn.functional.gelu(ref)) < 3e-5


@pytest.mark.parametrize("B", [3, 200])
def test_gemm_scatter_tokens(B):
    # stage-2-like map: 196 source rows per image scattered into the 49*6 token rows per image
    hw, P, S, N, K = 196, 49, 6, 768, 1024
    A = _gen((B * hw, K), 21).to(torch.bfloat16)
    W = _gen((N, K), 22, 0.05).to(torch.bfloat16)
    bias = _gen((N,), 23)
    pos = _gen((S, N), 24)
    g = torch.Generator().manual_seed(5)
    # random injective map of the 196 source rows into the 49*6 token rows, avoiding s == 0
    cand = torch.tensor([p * S + s for p in range(P) for s in range(1, S)])
    row_map = cand[torch.randperm(cand.numel(), generator=g)[:hw]].to(torch.int32).cuda()
    X = torch.zeros(B * P * S, N, device="cuda")
    ops.gemm(A, W, bias, X, ops.EPI_SCATTER_F32, row_map=row_map, rows_per_group=hw,
             dest_rows_per_group=P * S, pos=pos, pos_period=S)
    ref = A.float() @ W.float().t() + bias
    Xr = torch.zeros_like(X)
    for b in range(B):
        dst = b * P * S + row_map.long()
        Xr[dst] = ref[b * hw:(b + 1) * hw] + pos[(row_map.long() % S)]
    assert relerr(X, Xr) < 1e-4


@pytest.mark.parametrize("rows,D", [(1, 768), (1000, 768), (4214 * 3, 768), (77, 128), (33, 1024)])
def test_layernorm(rows, D):
    x = _gen((rows, D), 31, 50.0) + 10.0
    g = _gen((D,), 32)
    b = _gen((D,), 33)
    ref = torch.nn.functional.layer_norm(x, (D,), g, b, 1e-6)
    out = torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    ops.layernorm(x, g, b, out, 1e-6)
    assert relerr(out, ref) < 6e-3
    outs = torch.empty(rows, 2 * D, dtype=torch.bfloat16, device="cuda")
    ops.layernorm(x, g, b, outs, 1e-6)
    assert relerr(outs[:, :D].float() + outs[:, D:].float(), ref) < 2e-5


def _attn_ref(qkv, S, H, scale):
    rows, D3 = qkv.shape
    D = D3 // 3
    G = rows // S
    t = qkv.float().reshape(G, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    a = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
    return (a @ v).transpose(1, 2).reshape(rows, D)


@pytest.mark.parametrize("S,G,algo", [(6, 98, 1), (22, 50, 1), (22, 50, 2), (86, 49, 1), (86, 49, 2), (50, 7, 1), (50, 7, 2), (145, 3, 1), (17, 5, 2), (96, 4, 2),
                                          (86, 49, 3), (86, 700, 3), (96, 4, 3), (80, 6, 3), (71, 5, 3), (65, 3, 3),
                                          (6, 98, 4), (6, 49 * 128, 4), (6, 1, 4), (2, 33, 4), (8, 17, 4), (5, 40, 4), (6, 98, 0)])
def test_group_attention_bf16(S, G, algo):
    H = 12
    qkv = _gen((G * S, 3 * H * 64), 41 + S, 2.0).to(torch.bfloat16)
    ref = _attn_ref(qkv, S, H, 0.125)
    out = torch.empty(G * S, H * 64, dtype=torch.bfloat16, device="cuda")
    ops.group_attention(qkv, out, S, H, 0.125, algo=algo)
    assert relerr(out, ref) < (8e-3 if algo == 1 else 1.5e-2)


@pytest.mark.parametrize("H,S,G,algo", [(6, 86, 33, 3), (6, 86, 33, 2), (1, 90, 7, 3), (16, 70, 5, 3), (6, 50, 9, 1), (6, 6, 77, 4), (1, 6, 9, 4)])
def test_group_attention_other_head_counts(H, S, G, algo):
    """embed_dim = 64 * H for H != 12 (e.g. the 384-wide model): column offsets which*D + h*64 must follow H."""
    qkv = _gen((G * S, 3 * H * 64), 71 + S + H, 2.0).to(torch.bfloat16)
    ref = _attn_ref(qkv, S, H, 0.125)
    out = torch.empty(G * S, H * 64, dtype=torch.bfloat16, device="cuda")
    ops.group_attention(qkv, out, S, H, 0.125, algo=algo)
    assert relerr(out, ref) < 1.5e-2


@pytest.mark.parametrize("N,G,H", [(50, 7, 12), (50, 300, 12), (64, 3, 12), (17, 5, 6), (33, 4, 1)])
def test_patch_attention_split_precision_tcgen05(N, G, H):
    """Global patch attention on tcgen05 in split-bf16 precision: qkv as hi | lo pairs (what duo_gemm's SPLIT epilogue
    writes), out as a hi | lo pair; fp32-grade agreement with the reference on the same (split-rounded) inputs."""
    D = 64 * H
    x = _gen((G * N, 3 * D), 401 + N, 2.0)
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    qkv_split = torch.cat([hi, lo], dim=1).contiguous()
    xs = hi.float() + lo.float()  # what the kernel sees
    t = xs.reshape(G, N, 3, H, 64).permute(2, 0, 3, 1, 4).double()
    a = torch.softmax((t[0] @ t[1].transpose(-2, -1)) * 0.125, dim=-1)
    ref = (a @ t[2]).transpose(1, 2).reshape(G * N, D).float()
    out = torch.empty(G * N, 2 * D, dtype=torch.bfloat16, device="cuda")
    ops.group_attention(qkv_split, out, N, H, 0.125, split_in=True)
    got = out[:, :D].float() + out[:, D:].float()
    assert relerr(got, ref) < 1e-4
    # and against the fp32 FMA kernel it replaces
    out2 = torch.empty(G * N, 2 * D, dtype=torch.bfloat16, device="cuda")
    ops.group_attention(xs.contiguous(), out2, N, H, 0.125, algo=1)
    assert relerr(got, out2[:, :D].float() + out2[:, D:].float()) < 1e-4


@pytest.mark.parametrize("algo", [2, 3])
def test_group_attention_peaked_softmax_stays_finite(algo):
    """Scores of a few hundred (one key dominates every row): exp2 underflows to exactly 0 for the others,
    the result is finite and equals the fp32 reference row by row."""
    S, G, H = 86, 20, 12
    qkv = (_gen((G * S, 3 * H * 64), 301, 2.0) * 6.0).to(torch.bfloat16)
    ref = _attn_ref(qkv, S, H, 0.125)
    out = torch.empty(G * S, H * 64, dtype=torch.bfloat16, device="cuda")
    ops.group_attention(qkv, out, S, H, 0.125, algo=algo)
    assert torch.isfinite(out.float()).all()
    assert relerr(out, ref) < 1.5e-2


@pytest.mark.parametrize("S,G,algo,q_rows", [(86, 49, 2, 1), (86, 49, 1, 1), (22, 10, 2, 3), (6, 30, 1, 1), (50, 5, 2, 17), (6, 30, 4, 1),
                                             (6, 301, 4, 4), (6, 30, 0, 1)])
def test_group_attention_leading_query_rows(S, G, algo, q_rows):
    H = 12
    qkv = _gen((G * S, 3 * H * 64), 91 + S, 2.0).to(torch.bfloat16)
    ref = _attn_ref(qkv, S, H, 0.125).reshape(G, S, H * 64)[:, :q_rows].reshape(G * q_rows, H * 64)
    out = torch.empty(G * q_rows, H * 64, dtype=torch.bfloat16, device="cuda")
    ops.group_attention(qkv, out, S, H, 0.125, algo=algo, q_rows=q_rows)
    assert relerr(out, ref) < 1.5e-2


@pytest.mark.parametrize("rows,D", [(1000, 768), (77, 256), (33, 1024)])
def test_layernorm_statistics_output(rows, D):
    """duo_layernorm's optional stats_out: (mean, sum of squared deviations) of every 256-column part of the row — the
    layout the forwarding GEMMs exchange (used as the first producer's shift_stats)."""
    x = _gen((rows, D), 301, 3.0) + 20.0
    g, b = _gen((D,), 302) + 1.0, _gen((D,), 303)
    out = torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    st = torch.full((rows, D // 256, 2), float("nan"), device="cuda")
    ops.layernorm(x, g, b, out, 1e-6, stats_out=st)
    assert relerr(out, torch.nn.functional.layer_norm(x, (D,), g, b, 1e-6)) < 1e-2
    parts = x.view(rows, D // 256, 256)
    assert (st[:, :, 0] - parts.mean(dim=2)).abs().max().item() < 1e-4
    assert relerr(st[:, :, 1], ((parts - parts.mean(dim=2, keepdim=True)) ** 2).sum(dim=2)) < 1e-4
    mean, var = _merge_stats(st, D)
    assert relerr(var, x.var(dim=1, unbiased=False)) < 1e-4


def test_layernorm_strided_rows_and_residual_into_strided_view():
    R, S, D = 300, 6, 768
    X = _gen((R, S, D), 95, 5.0)
    g, b = _gen((D,), 96), _gen((D,), 97)
    X0 = X[:, 0, :]
    out = torch.empty(R, D, dtype=torch.bfloat16, device="cuda")
    ops.layernorm(X0, g, b, out, 1e-6)
    assert relerr(out, torch.nn.functional.layer_norm(X0, (D,), g, b, 1e-6)) < 6e-3
    A = _gen((R, D), 98).to(torch.bfloat16)
    W = _gen((D, D), 99, 0.05).to(torch.bfloat16)
    bias = _gen((D,), 100)
    ref = X.clone()
    ref[:, 0, :] += A.float() @ W.float().t() + bias
    ops.gemm(A, W, bias, X0, ops.EPI_RESIDUAL_F32)
    assert relerr(X, ref) < 1e-4


@pytest.mark.parametrize("S,G", [(6, 98), (86, 10), (50, 4)])
def test_group_attention_fp32(S, G):
    H = 12
    qkv = _gen((G * S, 3 * H * 64), 51 + S, 2.0)
    ref = _attn_ref(qkv, S, H, 0.0721)
    out = torch.empty(G * S, H * 64, dtype=torch.float32, device="cuda")
    ops.group_attention(qkv, out, S, H, 0.0721)
    assert relerr(out, ref) < 1e-5
    outs = torch.empty(G * S, 2 * H * 64, dtype=torch.bfloat16, device="cuda")
    ops.group_attention(qkv, outs, S, H, 0.0721)
    assert relerr(outs[:, :768].float() + outs[:, 768:].float(), ref) < 2e-5


def test_token_helpers_and_head():
    B, P, S, D, ncls = 3, 49, 6, 768, 10
    X = _gen((B, P, S, D), 61)
    X0 = X.clone()
    tok = _gen((D,), 62)
    pos_s = _gen((S, D), 63)
    ops.fill_scale_token(X, tok, pos_s[0].contiguous())
    ref = X0.clone()
    ref[:, :, 0, :] = tok + pos_s[0]
    assert torch.equal(X, ref)
    tokb = _gen((B, P, D), 64)
    ops.fill_scale_token(X, tokb, pos_s[0].contiguous())
    ref[:, :, 0, :] = tokb + pos_s[0]
    assert torch.equal(X, ref)

    cls = _gen((D,), 65)
    pos = _gen((P + 1, D), 66)
    Z = torch.empty(B, P + 1, D, dtype=torch.bfloat16, device="cuda")
    ops.assemble_patch_tokens(X, cls, pos, Z)
    Zr = torch.cat([cls.expand(B, 1, D), X[:, :, 0, :]], dim=1) + pos
    assert torch.equal(Z, Zr.to(torch.bfloat16))
    Zs = torch.empty(B, P + 1, 2 * D, dtype=torch.bfloat16, device="cuda")
    ops.assemble_patch_tokens(X, cls, pos, Zs)
    assert relerr(Zs[..., :D].float() + Zs[..., D:].float(), Zr) < 2e-5

    Wh = _gen((ncls, D), 67, 0.05)
    bh = _gen((ncls,), 68)
    Zf = _gen((B, P + 1, D), 69)
    logits = torch.empty(B, ncls, device="cuda")
    ops.head(Zf, (P + 1) * D, Wh, bh, logits)
    assert relerr(logits, Zf[:, 0] @ Wh.t() + bh) < 1e-5
    g, b = _gen((D,), 70), _gen((D,), 71)
    ops.head(Zf, (P + 1) * D, Wh, bh, logits, ln_gamma=g, ln_beta=b, eps=1e-6)
    lr = torch.nn.functional.layer_norm(Zf[:, 0], (D,), g, b, 1e-6) @ Wh.t() + bh
    assert relerr(logits, lr) < 1e-5


def test_invalid_arguments_raise():
    A = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    W = torch.zeros(100, 64, dtype=torch.bfloat16, device="cuda")
    out = torch.zeros(128, 100, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(RuntimeError, match="multiple of 128"):
        ops.gemm(A, W, None, out, ops.EPI_BF16)
    with pytest.raises(RuntimeError):
        ops.layernorm(torch.zeros(4, 100, device="cuda"), torch.zeros(100, device="cuda"),
                      torch.zeros(100, device="cuda"), torch.zeros(4, 100, dtype=torch.bfloat16, device="cuda"), 1e-6)


def _merge_stats(st, N):
    """[M, N/256, 2] (mean, M2) partial statistics -> per-row mean, biased variance (Chan's formula)."""
    parts = st.shape[1]
    mean = st[:, :, 0].mean(dim=1)
    m2 = st[:, :, 1].sum(dim=1) + 256.0 * ((st[:, :, 0] - mean[:, None]) ** 2).sum(dim=1)
    return mean, m2 / (256.0 * parts)


@pytest.mark.parametrize("M,K,with_gamma", [(256 * 60 + 77, 768, False), (256 * 60 + 77, 768, True), (256 * 152, 3072, False),
                                            (300, 768, True), (4214, 768, False), (129, 3072, False)])
def test_gemm_residual_statistics_forwarding_producer(M, K, with_gamma):
    """x = x + gamma*(A W^T + b) with xb_out / stats_out: the fp32 rows are updated exactly like the plain residual
    epilogue, xb_out is their bf16 rounding and stats_out holds (mean, M2) of every 256-column part of the fp32 row
    (any M: the forwarding epilogue always runs on the CTA-pair kernel; a large row mean exercises the shifted sums)."""
    N = 768
    A = _gen((M, K), 111).to(torch.bfloat16)
    W = _gen((N, K), 112, 0.05).to(torch.bfloat16)
    bias = _gen((N,), 113)
    gamma = _gen((N,), 114) if with_gamma else None
    X = _gen((M, N), 115, 5.0) + 70.0
    upd = A.float() @ W.float().t() + bias
    Xr = X + (gamma * upd if with_gamma else upd)
    xb = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    st = torch.full((M, N // 256, 2), float("nan"), device="cuda")
    X0 = X.clone()
    for _ in range(2):
        X.copy_(X0)
        ops.gemm(A, W, bias, X, ops.EPI_RESIDUAL_F32, gamma=gamma, xb_out=xb, stats_out=st)
        assert relerr(X, Xr) < 1e-4
        assert torch.equal(xb, X.to(torch.bfloat16))  # the copy is the rounding of the stored fp32 rows
        mean, var = _merge_stats(st, N)
        assert torch.isfinite(st).all()
        assert (mean - X.mean(dim=1)).abs().max().item() < 1e-3
        # (without shift_stats the sums are taken around 0: raw fp32 sums of x and x^2 at |mean| = 14 x spread)
        assert relerr(var, X.var(dim=1, unbiased=False)) < 1e-3
        parts = X.view(M, N // 256, 256)
        assert (st[:, :, 0] - parts.mean(dim=2)).abs().max().item() < 1e-3
        assert relerr(st[:, :, 1], ((parts - parts.mean(dim=2, keepdim=True)) ** 2).sum(dim=2)) < 2e-3


@pytest.mark.parametrize("M", [300, 256 * 60 + 77])
def test_gemm_forwarding_row_shift(M):
    """shift_stats: the bf16 copy is bf16(x - m), m = the row mean according to the PREVIOUS statistics; X and the new
    statistics are unaffected."""
    N = K = 768
    A = _gen((M, K), 211).to(torch.bfloat16)
    W = _gen((N, K), 212, 0.05).to(torch.bfloat16)
    bias = _gen((N,), 213)
    X = _gen((M, N), 214, 3.0) + 40.0
    prev_parts = X.view(M, N // 256, 256)
    prev = torch.stack([prev_parts.mean(dim=2) + 0.25, torch.ones(M, N // 256, device="cuda")], dim=2).contiguous()  # any estimate
    Xr = X + A.float() @ W.float().t() + bias
    xb = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    st = torch.empty(M, N // 256, 2, device="cuda")
    ops.gemm(A, W, bias, X, ops.EPI_RESIDUAL_F32, xb_out=xb, stats_out=st, shift_stats=prev)
    assert relerr(X, Xr) < 1e-4
    shift = prev[:, :, 0].mean(dim=1, keepdim=True)
    # the kernel's shift may differ from torch's mean by an fp32 ulp of its magnitude (40: 4e-6), which matters for
    # elements near zero: compare within one bf16 rounding plus a few such ulps
    want = X - shift
    assert bool(((xb.float() - want).abs() <= want.abs() * 2.0 ** -8 + 3e-5).all())
    assert (xb.float() - X).abs().min().item() > 30.0  # and it is the shifted copy, not x itself
    parts = X.view(M, N // 256, 256)
    assert (st[:, :, 0] - parts.mean(dim=2)).abs().max().item() < 1e-4
    assert relerr(st[:, :, 1], ((parts - parts.mean(dim=2, keepdim=True)) ** 2).sum(dim=2)) < 1e-4  # the shift is the pivot
    with pytest.raises(AssertionError):  # the previous statistics must not be the buffer being written
        ops.gemm(A, W, bias, X, ops.EPI_RESIDUAL_F32, xb_out=xb, stats_out=st, shift_stats=st)


@pytest.mark.parametrize("M,N,epi", [(256 * 60 + 77, 2304, "bf16"), (256 * 60 + 77, 3072, "gelu"), (300, 2304, "bf16"),
                                     (300, 3072, "gelu"), (4214, 3072, "gelu"), (128 * 310 + 5, 768, "bf16")])
def test_gemm_forwarded_layernorm_consumer(M, N, epi):
    """out = f(Linear(LayerNorm(x))) computed from the UN-normalised bf16 copy of x, its forwarded statistics and the
    folded weights (engine.pack_ln_linear), against LayerNorm -> Linear in fp32 on the same x (pair and 1-CTA kernels)."""
    from duoformer_tcga_b200 import engine

    K = 768
    x = _gen((M, K), 121, 4.0) + _gen((M, 1), 122, 2.0)  # per-row means of the order of the spread
    lw, lb = _gen((K,), 123, 0.1) + 1.0, _gen((K,), 124, 0.05)
    W = _gen((N, K), 125, 0.03)
    bias = _gen((N,), 126, 0.1)
    ref = torch.nn.functional.layer_norm(x, (K,), lw, lb, 1e-6) @ W.t() + bias
    if epi == "gelu":
        ref = torch.nn.functional.gelu(ref)
    parts = x.view(M, K // 256, 256)
    pm = parts.mean(dim=2)
    st = torch.stack([pm, ((parts - pm[:, :, None]) ** 2).sum(dim=2)], dim=2).contiguous()
    wp, bp = engine.pack_ln_linear(W, bias, lw, lb)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(x.to(torch.bfloat16), wp, bp, out, ops.EPI_GELU_BF16 if epi == "gelu" else ops.EPI_BF16,
             ln_stats=st, ln_eps=1e-6)
    assert relerr(out, ref) < 1.5e-2


@pytest.mark.parametrize("offset", [0.0, 30.0, 300.0])
def test_gemm_statistics_forwarding_chain_matches_layernorm_path(offset):
    """proj(+residual) -> fc1(+GELU) through the forwarding pair of epilogues equals residual GEMM -> LayerNorm
    kernel -> fc1 GEMM (the round-1 sequence) within bf16 rounding — also when the rows carry a mean of 10x / 100x their
    spread: the producer subtracts the previous row mean (shift_stats, here the LayerNorm kernel's statistics of the
    rows before the update) before rounding, and the consumer's row-centred weights make the shift drop out."""
    from duoformer_tcga_b200 import engine

    M, D, Hd = 256 * 40 + 13, 768, 3072
    A = _gen((M, D), 141).to(torch.bfloat16)
    Wp = _gen((D, D), 142, 0.05).to(torch.bfloat16)
    bp = _gen((D,), 143)
    X = _gen((M, D), 144, 3.0) + offset
    lw, lb = _gen((D,), 145, 0.1) + 1.0, _gen((D,), 146, 0.05)
    W1 = _gen((Hd, D), 147, 0.03)
    b1 = _gen((Hd,), 148, 0.1)
    # round-1 sequence
    X1 = X.clone()
    ops.gemm(A, Wp, bp, X1, ops.EPI_RESIDUAL_F32)
    hn = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    ops.layernorm(X1, lw, lb, hn, 1e-6)
    h1 = torch.empty(M, Hd, dtype=torch.bfloat16, device="cuda")
    ops.gemm(hn, W1.to(torch.bfloat16), b1, h1, ops.EPI_GELU_BF16)
    # forwarding sequence (the rows' previous statistics come from a LayerNorm launch, as in block 0 of the model)
    X2 = X.clone()
    xb = torch.empty(M, D, dtype=torch.bfloat16, device="cuda")
    st = torch.empty(M, D // 256, 2, device="cuda")
    st_prev = torch.empty(M, D // 256, 2, device="cuda")
    ops.layernorm(X2, lw, lb, torch.empty(M, D, dtype=torch.bfloat16, device="cuda"), 1e-6, stats_out=st_prev)
    ops.gemm(A, Wp, bp, X2, ops.EPI_RESIDUAL_F32, xb_out=xb, stats_out=st, shift_stats=st_prev)
    w, b = engine.pack_ln_linear(W1, b1, lw, lb)
    h2 = torch.empty(M, Hd, dtype=torch.bfloat16, device="cuda")
    ops.gemm(xb, w, b, h2, ops.EPI_GELU_BF16, ln_stats=st, ln_eps=1e-6)
    assert relerr(X2, X1) < 1e-5
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(X1, (D,), lw, lb, 1e-6) @ W1.t() + b1)
    e1, e2 = relerr(h1, ref), relerr(h2, ref)
    assert e2 < 1.5e-2 and e2 < 2.0 * e1 + 1e-3, (e1, e2)


def test_gemm_forwarding_argument_checks():
    A = torch.zeros(256, 768, dtype=torch.bfloat16, device="cuda")
    W = torch.zeros(768, 768, dtype=torch.bfloat16, device="cuda")
    X = torch.zeros(256, 768, device="cuda")
    xb = torch.zeros(256, 768, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(256, 3, 2, device="cuda")
    with pytest.raises(RuntimeError, match="statistics forwarding"):  # only with the residual epilogue
        ops.gemm(A, W, None, xb, ops.EPI_BF16, xb_out=xb, stats_out=st)
    with pytest.raises(RuntimeError, match="forwarded LayerNorm"):  # not with the residual epilogue
        ops.gemm(A, W, None, X, ops.EPI_RESIDUAL_F32, ln_stats=st)


@pytest.mark.parametrize("M,N,K", [(300, 768, 256), (256 * 150 + 3, 768, 512)])
def test_gemm_fp16_operands(M, N, K):
    """fp16_operands: A (the cuDNN trunk maps) and W in IEEE fp16, tcgen05.mma kind::f16 (1-CTA and pair kernels)."""
    A = _gen((M, K), 131).to(torch.float16)
    W = _gen((N, K), 132, 0.05).to(torch.float16)
    bias = _gen((N,), 133)
    ref = A.float() @ W.float().t() + bias
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(A, W, bias, out, ops.EPI_F32)
    assert relerr(out, ref) < 2e-5  # exact products, fp32 accumulation


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,C,H,W", [(3, 64, 112, 112), (2, 8, 7, 9), (1, 16, 1, 1)])
def test_maxpool3x3s2_nhwc_matches_torch(B, C, H, W, dtype):
    """duo_maxpool3x3s2 == nn.MaxPool2d(3, 2, 1) bit for bit (the ResNet stem pool), odd sizes included."""
    x = _gen((B, C, H, W), 141).to(dtype).contiguous(memory_format=torch.channels_last)
    ref = torch.nn.functional.max_pool2d(x, 3, 2, 1)
    out = ops.maxpool3x3s2(x)
    assert out.shape == ref.shape and torch.equal(out, ref)


def test_gemm_split_mode_2_plain_A_split_W():
    """split3 == 2: exact bf16 activations times hi|lo-split weights (A*Wh + A*Wl)."""
    M, N, K = 700, 768, 768
    A = _gen((M, K), 121).to(torch.bfloat16)
    W = _gen((N, K), 122, 0.05)
    bias = _gen((N,), 123)
    ref = (A.double() @ W.double().t() + bias.double()).float()
    out = torch.empty(M, N, dtype=torch.float32, device="cuda")
    ops.gemm(A, ops.split_weight(W), bias, out, ops.EPI_F32, split3=2)
    assert relerr(out, ref) < 2e-5


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize("H,C,N,stride", [(28, 256, 256, 2), (14, 512, 512, 2), (7, 768, 768, 1), (12, 256, 128, 1)])
def test_conv3x3_as_im2col_plus_gemm(H, C, N, stride, dtype):
    """Conv2d(k=3, pad=1, stride s) + bias (+ ReLU) == duo_im2col3x3 + duo_gemm with the weight permuted
    to [N, ky, kx, c]  (channel-token branch, projection_head.py:152-268)."""
    B = 3
    x = _gen((B, C, H, H), 131).to(dtype)
    w = _gen((N, C, 3, 3), 132, (2.0 / (9 * C)) ** 0.5)
    bias = _gen((N,), 133, 0.1)
    ref = torch.nn.functional.conv2d(x.float().to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), bias, stride=stride, padding=1)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    a = ops.im2col3x3(x_nhwc, stride)
    Ho = ref.shape[2]
    assert a.shape == (B * Ho * Ho, 9 * C)
    wg = w.permute(0, 2, 3, 1).reshape(N, 9 * C).to(torch.bfloat16).contiguous()
    out = torch.empty(B * Ho * Ho, N, dtype=torch.float32, device="cuda")
    ops.gemm(a, wg, bias, out, ops.EPI_F32)
    assert relerr(out.view(B, Ho, Ho, N).permute(0, 3, 1, 2), ref) < 2e-4
    ops.gemm(a, wg, bias, out, ops.EPI_F32, relu=True)
    assert relerr(out.view(B, Ho, Ho, N).permute(0, 3, 1, 2), ref.clamp_min(0)) < 2e-4
    outb = torch.empty(B * Ho * Ho, N, dtype=torch.bfloat16, device="cuda")
    ops.gemm(a, wg, bias, outb, ops.EPI_BF16, relu=True)
    assert relerr(outb.view(B, Ho, Ho, N).permute(0, 3, 1, 2), ref.clamp_min(0)) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_pool_to_slice_concatenates_like_maxpool_and_cat(dtype):
    B, g = 2, 7
    a = _gen((B, 2 * g, 2 * g, 256), 141).to(dtype)
    b = _gen((B, g, g, 128), 142).to(dtype)
    cat = torch.zeros(B * g * g, 384, dtype=torch.bfloat16, device="cuda")
    ops.pool_to_slice(a, cat[:, :256], 2)
    ops.pool_to_slice(b, cat[:, 256:], 1)
    ra = torch.nn.functional.max_pool2d(a.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).reshape(B * g * g, 256)
    ref = torch.cat([ra, b.float().reshape(B * g * g, 128)], dim=1).to(torch.bfloat16)
    assert torch.equal(cat, ref)
