"""Precision study (CPU, run by hand): LayerNorm statistics forwarding vs LayerNorm-then-round.

Emulates the bf16-mode scale stage of the product in torch on the golden `wo4_d12` case (4-scale, depth 12,
batch 2, |token| up to O(10^2)) with three ways of feeding the QKV / fc1 GEMMs:

  ln_round   (round 1)  A = bf16(LN(x) * g + b), W = bf16(W)
  forward    A = bf16(x), W' = bf16(W * g); out = rstd * (A W'^T - mu * colsum(W')) + (W b_ln + bias)
  forward_shift  as `forward` with A = bf16(x - c_row), c_row = row mean of the block-0 input (fixed per row)
  centred    A = bf16(x), W'' = engine.pack_ln_linear: bf16(W * g) with centred rows, so the mean cancels inside the
             product; out = rstd * (A W''^T) + (W b_ln + bias)
  product    (the product path) as `centred` with A = bf16(x - m_prev), m_prev = the row mean at the PREVIOUS LayerNorm
             point (duo_gemm's shift_stats); block 0's norm1 runs as LayerNorm-then-round (the standalone launch)

  usage: diag_ln_forwarding.py [golden case] [offset added to every token: stress for |mean| >> spread]

and reports the relative max-norm error of every scale block's output and of the logits against the fp32
oracle.  Statistics always come from the fp32 stream (the residual epilogue holds the fp32 row).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.nn.functional as F

from common import COMMON, load_golden, oracle_forward, relerr, build_product
from oracle import synth
from oracle import duoformer_oracle as orc

torch.set_num_threads(os.cpu_count() or 1)


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


PREV_MEAN = [None]  # row mean at the previous LayerNorm point (mode "product")


def lin_ln(x, sd, lnp, linp, mode, shift):
    g, b = sd[lnp + "weight"], sd[lnp + "bias"]
    W, bias = sd[linp + "weight"], sd[linp + "bias"]
    mu = x.mean(-1, keepdim=True)
    var = x.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-6)
    prev, PREV_MEAN[0] = PREV_MEAN[0], mu
    if mode == "ln_round" or (mode == "product" and prev is None):
        a = bf((x - mu) * rstd * g + b)
        return a @ bf(W).t() + bias
    if mode in ("centred", "product"):
        from duoformer_tcga_b200 import engine

        wp, bp = engine.pack_ln_linear(W, bias, g, b)
        c = prev if mode == "product" else torch.zeros_like(mu)
        return rstd * (bf(x - c) @ wp.float().t()) + bp
    Wp = bf(W * g)
    cs = Wp.sum(-1)
    c = shift if mode == "forward_shift" else torch.zeros_like(mu)
    a = bf(x - c)
    acc = a @ Wp.t()
    return rstd * (acc - (mu - c) * cs) + (W @ b + bias)


def scale_stage(x, sd, depth, H, mode):
    p = "vision_transformer."
    C = x.shape[-1]
    scale = (C // H) ** -0.5
    shift = x.mean(-1, keepdim=True)
    PREV_MEAN[0] = None
    outs = []
    stats = []
    for i in range(depth):
        blk = f"{p}scaleBlocks.{i}."
        B, P, S, _ = x.shape
        mu, sdv = x.mean(-1), x.std(-1)
        stats.append((float((mu.abs() / sdv).max()), float((mu.abs() / sdv).mean()), float(x.abs().max()),
                      float(((mu - shift.squeeze(-1)).abs() / sdv).max())))
        qkv = bf(lin_ln(x, sd, blk + "norm1.", blk + "attn.qkv.", mode, shift))
        qkv = qkv.reshape(B, P, S, 3, H, C // H).permute(3, 0, 1, 4, 2, 5)
        q, k, v = qkv[0], qkv[1], qkv[2]
        attn = ((q @ k.transpose(-2, -1)) * scale).softmax(dim=-1)
        o = bf((bf(attn) @ v).transpose(2, 3).reshape(B, P, S, C))
        x = x + (o @ bf(sd[blk + "attn.proj.weight"]).t() + sd[blk + "attn.proj.bias"])
        h = bf(F.gelu(lin_ln(x, sd, blk + "norm2.", blk + "mlp.fc1.", mode, shift)))
        x = x + (h @ bf(sd[blk + "mlp.fc2.weight"]).t() + sd[blk + "mlp.fc2.bias"])
        outs.append(x.clone())
    return x, outs, stats


def tail(x, sd, depth, H):
    p = "vision_transformer."
    C = x.shape[-1]
    scale = (C // H) ** -0.5
    B = x.shape[0]
    cls = sd[p + "cls_token"].expand(B, -1, -1)
    z = torch.cat((cls, x[:, :, 0, :]), dim=1) + sd[p + "pos_embed"]
    for i in range(depth):
        b = f"{p}blocks.{i}."
        z = orc.region_attention(z, sd, b + "attn.qkv.", b + "attn.proj.", H, scale)
    return orc._lin(z[:, 0, :], sd, p + "head.")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "wo4_d12"
    gold = load_golden(name)
    case = gold["case"]
    model = build_product(case)
    sd = synth.synth_state_dict(model.state_dict(), seed=gold["weight_seed"])
    x = synth.synth_images(case["batch"], seed=gold["input_seed"])
    offset = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    ocap = {}
    with torch.no_grad():
        yo = oracle_forward(case, x, sd, capture=ocap)
        tokens = ocap["tokens"] + offset
        if offset:  # fp32 reference of the shifted problem (LayerNorm removes the offset of the block inputs, not of x)
            ocap = {}
            xs_ref, outs_ref, _ = None, [], None
            sd_ref = sd
            from oracle import duoformer_oracle as _o
            xr = tokens.clone()
            p = "vision_transformer."
            scale = (xr.shape[-1] // COMMON["num_heads"]) ** -0.5
            for i in range(case["depth"]):
                bk = f"{p}scaleBlocks.{i}."
                xr = xr + _o.scale_attention(_o._ln(xr, sd, bk + "norm1."), sd, bk + "attn.qkv.", bk + "attn.proj.", COMMON["num_heads"], scale)
                xr = xr + _o._mlp(_o._ln(xr, sd, bk + "norm2."), sd, bk + "mlp.")
                ocap[f"scale_block_{i}"] = xr.clone()
            yo = tail(xr, sd, case["depth"], COMMON["num_heads"])
        for mode in ("ln_round", "forward", "forward_shift", "centred", "product"):
            xs, outs, stats = scale_stage(tokens.clone(), sd, case["depth"], COMMON["num_heads"], mode)
            y = tail(xs, sd, case["depth"], COMMON["num_heads"])
            errs = [relerr(o - offset, ocap[f"scale_block_{i}"] - offset) for i, o in enumerate(outs)]
            print(f"== {mode}: logits rel err {relerr(y, yo):.3e}; worst block {max(errs):.3e}")
            print("   per block:", " ".join(f"{e:.2e}" for e in errs))
            if mode == "ln_round":
                for i, s in enumerate(stats):
                    print(f"   block {i}: max|mu|/sd {s[0]:.2f} mean {s[1]:.3f} max|x| {s[2]:.1f} max|mu-shift|/sd {s[3]:.2f}")


if __name__ == "__main__":
    main()
