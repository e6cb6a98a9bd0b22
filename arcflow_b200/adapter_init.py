"""Initial ArcFlow adapter tensors for a stock (pre-trained or synthetic) base transformer.

Follows the reference's student construction: the three heads start from the base model's velocity head
(`proj_out` repeated K times, a small per-(component, latent channel) bias jitter so the components separate), zero
log-weights, zero log-gamma weight with the log-spaced rate bias, `norm_out` taken over from the base, and peft's
'gaussian' LoRA init (A ~ N(0, (1/r)^2), B = 0) on the configured targets —
lakonlab/models/architecture/arcflow/arcflux.py:92-132, :318-341, :295-302 (arcqwen.py has the same code).
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import torch


def loggamma_bias(num_gaussians: int, logweights_channels: int, lo: float = 0.2, hi: float = 4.0) -> torch.Tensor:
    g = torch.logspace(math.log10(lo), math.log10(hi), num_gaussians - 1, base=10).log()
    return g.unsqueeze(1).repeat(1, logweights_channels).flatten()


def init_arcflow_adapter(base_sd: Dict[str, torch.Tensor], cfg, lora_targets: Iterable[str],
                         generator: Optional[torch.Generator] = None, dtype=torch.bfloat16,
                         bias_jitter: float = 0.05) -> Dict[str, torch.Tensor]:
    """base_sd: stock transformer state dict (diffusers names, with `proj_out.*` and `norm_out.linear.*`).
    lora_targets: full module paths (e.g. 'transformer_blocks.0.ff.net.2'). Returns ONLY the adapter tensors, under the
    on-disk names of export_arcflow_to_diffusers.py:104-127."""
    K, C, L = cfg.num_gaussians, cfg.out_channels, cfg.logweights_channels
    D = cfg.inner_dim
    w, b = base_sd["proj_out.weight"].float().cpu(), base_sd["proj_out.bias"].float().cpu()   # draws are made on the CPU
    if w.shape != (C, D):
        raise ValueError(f"proj_out.weight has shape {tuple(w.shape)}, expected {(C, D)}")
    out: Dict[str, torch.Tensor] = {}
    out["proj_out_means.weight"] = w[None].expand(K, -1, -1).reshape(K * C, D).clone()
    jitter = torch.randn(K * C // L, generator=generator) * bias_jitter
    out["proj_out_means.bias"] = b[None].expand(K, -1).reshape(K * C) + jitter[:, None].expand(-1, L).flatten()
    out["proj_out_logweights.weight"] = torch.zeros(K * L, D)
    out["proj_out_logweights.bias"] = torch.zeros(K * L)
    out["proj_out_loggamma.weight"] = torch.zeros((K - 1) * L, D)
    out["proj_out_loggamma.bias"] = loggamma_bias(K, L)
    out["norm_out.linear.weight"] = base_sd["norm_out.linear.weight"].float().cpu().clone()
    out["norm_out.linear.bias"] = base_sd["norm_out.linear.bias"].float().cpu().clone()
    r = int(cfg.lora_rank)
    if r > 0:
        for name in lora_targets:
            o, i = base_sd[name + ".weight"].shape
            out[name + ".lora_A.weight"] = torch.randn(r, i, generator=generator) / r
            out[name + ".lora_B.weight"] = torch.zeros(o, r)
    return {k: v.to(dtype) for k, v in out.items()}


def flux_lora_target_paths(cfg, target_suffixes: Iterable[str]) -> list:
    """Expands peft-style suffix targets (configs/flux/arcflux_2nfe_k16.py:40-48) to full module paths."""
    from .synthetic import flux_linear_shapes, _is_lora_target
    suffixes = tuple(target_suffixes)
    return [n for n in flux_linear_shapes(cfg) if _is_lora_target(n, suffixes)]
