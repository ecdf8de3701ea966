"""Pixel -> token maps of the multiscale token builder, generalised to any patch grid g.

The reference hard-codes g = 7 as Python literals (model_wo_extra_params.py:110-212,
model.py:106-211): stage k in {3,2,1,0} has a w x w window per patch, w = 2**(3-k); token order
inside a patch is [scale token, stage 3, stage 2 (column-major 2x2: TL,BL,TR,BR), stage 1
(row-major 4x4), stage 0 (row-major 8x8)] truncated to the first ``num_layers`` stages.
"""
from __future__ import annotations

from typing import Dict, List

import torch

STAGE_CHANNELS_R50 = {3: 2048, 2: 1024, 1: 512, 0: 256}
STAGE_CHANNELS_R18 = {3: 512, 2: 256, 1: 128, 0: 64}


def num_scale_tokens(num_layers: int) -> int:
    """S = 1 + sum_{k<num_layers} 4**k  (2 / 6 / 22 / 86 for 1..4 scales)."""
    return 1 + sum(4**k for k in range(num_layers))


def stages_used(num_layers: int) -> List[int]:
    """Stages in token order: 3, 2, 1, 0 truncated (model_wo_extra_params.py:264,281,294)."""
    return [3, 2, 1, 0][:num_layers]


def window_offsets(k: int) -> List[tuple]:
    """(dr, dc) enumeration order of the w x w window of stage k."""
    w = 2 ** (3 - k)
    if k == 2:  # column-major: dc outer, dr inner (model_wo_extra_params.py:117-124)
        return [(dr, dc) for dc in range(w) for dr in range(w)]
    return [(dr, dc) for dr in range(w) for dc in range(w)]  # row-major (:125-212)


def gather_index(k: int, g: int) -> torch.Tensor:
    """Reference-style gather table: [g*g, w*w] int64 of flat pixel indices of stage k."""
    w = 2 ** (3 - k)
    G = g * w
    offs = window_offsets(k)
    idx = torch.empty(g * g, w * w, dtype=torch.int64)
    for r in range(g):
        for c in range(g):
            idx[r * g + c] = torch.tensor([(w * r + dr) * G + (w * c + dc) for dr, dc in offs])
    return idx


def token_row_maps(num_layers: int, g: int) -> Dict[int, torch.Tensor]:
    """Scatter maps for the fused projection kernel: for each used stage k an int32 tensor
    [ (g*w)**2 ] giving, for flat pixel y*G+x, the token row p*S + s inside one image."""
    S = num_scale_tokens(num_layers)
    maps: Dict[int, torch.Tensor] = {}
    s_off = 1
    for k in stages_used(num_layers):
        w = 2 ** (3 - k)
        G = g * w
        m = torch.empty(G * G, dtype=torch.int32)
        idx = gather_index(k, g)  # [P, w*w]
        P = g * g
        dest = (torch.arange(P).unsqueeze(1) * S + s_off + torch.arange(w * w).unsqueeze(0)).to(torch.int32)
        m[idx.reshape(-1)] = dest.reshape(-1)
        maps[k] = m
        s_off += w * w
    return maps
