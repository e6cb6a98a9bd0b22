"""Rotary tables for the joint [text; image] sequence, fp32 [S, 128] (cos, sin), adjacent-pair layout.

FLUX: diffusers FluxPosEmbed(theta=10000, axes_dim=(16, 56, 56)) evaluated in float64 on ids
(text rows all zero, image rows (0, row, col)), repeat-interleaved, cast to fp32 and then — as the
reference does at lakonlab/models/architecture/arcflow/arcflux.py:171-173 — rounded to bf16.
"""
from __future__ import annotations

from typing import Tuple

import torch


def _axis_freqs(pos: torch.Tensor, dim: int, theta: float = 10000.0) -> torch.Tensor:
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float64)[: dim // 2] / dim))
    return torch.outer(pos.to(torch.float64), freqs)


def flux_rope_tables(txt_len: int, grid_h: int, grid_w: int, axes_dims=(16, 56, 56),
                     round_bf16: bool = True, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    ids = torch.zeros(txt_len + grid_h * grid_w, 3, dtype=torch.float64)
    rows = torch.arange(grid_h, dtype=torch.float64)[:, None].expand(grid_h, grid_w).reshape(-1)
    cols = torch.arange(grid_w, dtype=torch.float64)[None, :].expand(grid_h, grid_w).reshape(-1)
    ids[txt_len:, 1] = rows
    ids[txt_len:, 2] = cols
    cos, sin = [], []
    for i, d in enumerate(axes_dims):
        ang = _axis_freqs(ids[:, i], d)
        cos.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin.append(ang.sin().repeat_interleave(2, dim=1).float())
    cos, sin = torch.cat(cos, dim=-1), torch.cat(sin, dim=-1)
    if round_bf16:
        cos, sin = cos.bfloat16().float(), sin.bfloat16().float()
    return cos.contiguous().to(device), sin.contiguous().to(device)
