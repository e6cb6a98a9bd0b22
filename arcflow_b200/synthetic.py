"""Synthetic weights / inputs of the true shapes (no checkpoints exist offline).

State-dict keys are the reference's on-disk names: the diffusers FLUX transformer keys that
`pipe.transformer.state_dict()` yields (lakonlab/pipelines/arcflow_loader.py:242) plus the adapter keys
written by export_arcflow_to_diffusers.py:104-127 (`<path>.lora_A.weight`, `<path>.lora_B.weight`,
`proj_out_means|logweights|loggamma.*`, `norm_out.linear.*`).

Distributions: Linear weights/biases N(0, 0.02^2); RMSNorm scales 1 + N(0, 0.02^2); LoRA A ~ N(0, (1/r)^2)
(peft 'gaussian' init), LoRA B ~ N(0, 0.02^2) (NOT the reference's zero init, so the branch is exercised);
`proj_out_loggamma.bias` = the reference's init ln(logspace(log10 .2, log10 4, K-1)) (arcflux.py:115-132).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from .config import ArcFluxConfig, FLUX_LORA_TARGETS


def _is_lora_target(name: str, targets) -> bool:
    return any(name == t or name.endswith("." + t) for t in targets)


def flux_linear_shapes(cfg: ArcFluxConfig) -> Dict[str, tuple]:
    """name -> (out_features, in_features) of every Linear in the ArcFlow-FLUX student."""
    D, M = cfg.inner_dim, cfg.mlp_dim
    shapes = {
        "x_embedder": (D, cfg.in_channels),
        "context_embedder": (D, cfg.joint_attention_dim),
        "time_text_embed.timestep_embedder.linear_1": (D, 256),
        "time_text_embed.timestep_embedder.linear_2": (D, D),
        "time_text_embed.text_embedder.linear_1": (D, cfg.pooled_projection_dim),
        "time_text_embed.text_embedder.linear_2": (D, D),
    }
    if cfg.guidance_embeds:
        shapes["time_text_embed.guidance_embedder.linear_1"] = (D, 256)
        shapes["time_text_embed.guidance_embedder.linear_2"] = (D, D)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        shapes[p + "norm1.linear"] = (6 * D, D)
        shapes[p + "norm1_context.linear"] = (6 * D, D)
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            shapes[p + "attn." + n] = (D, D)
        shapes[p + "ff.net.0.proj"] = (M, D)
        shapes[p + "ff.net.2"] = (D, M)
        shapes[p + "ff_context.net.0.proj"] = (M, D)
        shapes[p + "ff_context.net.2"] = (D, M)
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}."
        shapes[p + "norm.linear"] = (3 * D, D)
        for n in ("to_q", "to_k", "to_v"):
            shapes[p + "attn." + n] = (D, D)
        shapes[p + "proj_mlp"] = (M, D)
        shapes[p + "proj_out"] = (D, D + M)
    shapes["norm_out.linear"] = (2 * D, D)
    nm, nw, ng = cfg.head_dims
    shapes["proj_out_means"] = (nm, D)
    shapes["proj_out_logweights"] = (nw, D)
    shapes["proj_out_loggamma"] = (ng, D)
    return shapes


def flux_rmsnorm_names(cfg: ArcFluxConfig):
    names = []
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}.attn."
        names += [p + "norm_q", p + "norm_k", p + "norm_added_q", p + "norm_added_k"]
    for i in range(cfg.num_single_layers):
        p = f"single_transformer_blocks.{i}.attn."
        names += [p + "norm_q", p + "norm_k"]
    return names


def make_flux_state_dict(cfg: ArcFluxConfig, seed: int = 1234, device="cpu",
                         dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded synthetic ArcFlow-FLUX student (base trunk + adapter). Values are bf16-representable."""
    g = torch.Generator(device=device).manual_seed(seed)

    def normal(shape, std, mean=0.0):
        t = torch.empty(shape, device=device, dtype=torch.float32).normal_(mean, std, generator=g)
        return t.to(dtype)

    sd: Dict[str, torch.Tensor] = {}
    r = cfg.lora_rank
    for name, (o, i) in flux_linear_shapes(cfg).items():
        sd[name + ".weight"] = normal((o, i), 0.02)
        sd[name + ".bias"] = normal((o,), 0.02)
        if r > 0 and _is_lora_target(name, FLUX_LORA_TARGETS):
            sd[name + ".lora_A.weight"] = normal((r, i), 1.0 / r)
            sd[name + ".lora_B.weight"] = normal((o, r), 0.02)
    for name in flux_rmsnorm_names(cfg):
        sd[name + ".weight"] = normal((cfg.attention_head_dim,), 0.02, mean=1.0)
    gam = torch.logspace(math.log10(0.2), math.log10(4.0), cfg.num_gaussians - 1, base=10).log()
    gam = gam.unsqueeze(1).repeat(1, cfg.logweights_channels).flatten()
    sd["proj_out_loggamma.bias"] = gam.to(device=device, dtype=dtype)
    return sd


def make_flux_inputs(cfg: ArcFluxConfig, batch: int, height: int, width: int, txt_len: int = 512,
                     seed: int = 42, device="cpu"):
    """Seeded synthetic latents (packed tokens, fp32) + cached text embeds (bf16), SURVEY.md §8d."""
    g = torch.Generator(device=device).manual_seed(seed)
    gh, gw = height // 16, width // 16
    lat = torch.empty(batch, 16, 2 * gh, 2 * gw, device=device).normal_(generator=g)
    # FluxPipeline._pack_latents: (B, C, 2h, 2w) -> (B, h*w, C*4) with channel index c*4 + ph*2 + pw
    x = lat.view(batch, 16, gh, 2, gw, 2).permute(0, 2, 4, 1, 3, 5).reshape(batch, gh * gw, 64).contiguous()
    txt = (torch.empty(batch, txt_len, cfg.joint_attention_dim, device=device).normal_(generator=g) * 0.1
           ).to(torch.bfloat16)
    pooled = torch.empty(batch, cfg.pooled_projection_dim, device=device).normal_(generator=g).to(torch.bfloat16)
    return x, txt, pooled


def make_flux_teacher_extras(cfg: ArcFluxConfig, seed: int = 4321, device="cpu", dtype=torch.bfloat16):
    """The teacher-only tensors of the stock FLUX transformer (its `norm_out.linear` and `proj_out`); the trunk is
    tied to the student's frozen base layers (lakonlab/models/base_diffusion.py:93-94)."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = cfg.inner_dim

    def normal(shape, std):
        return torch.empty(shape, device=device, dtype=torch.float32).normal_(0.0, std, generator=g).to(dtype)

    return {"norm_out.linear.weight": normal((2 * D, D), 0.02), "norm_out.linear.bias": normal((2 * D,), 0.02),
            "proj_out.weight": normal((cfg.out_channels, D), 0.02), "proj_out.bias": normal((cfg.out_channels,), 0.02)}


def make_vae_decoder_state_dict(ch: int = 128, ch_mult=(1, 2, 4, 4), z_channels: int = 16, out_ch: int = 3,
                                num_res_blocks: int = 2, seed: int = 7, device="cpu", dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded synthetic decoder weights of the BFL layout (no checkpoint exists offline): conv weights N(0, 1/fan_in) so the
    activations keep O(1) scale through the 30 convolutions, biases N(0, 0.02^2), GroupNorm scales 1 + N(0, 0.05^2)."""
    g = torch.Generator(device=device).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cin, cout, k):
        std = (1.0 / (cin * k * k)) ** 0.5
        sd[name + ".weight"] = (torch.randn(cout, cin, k, k, generator=g, device=device) * std).to(dtype)
        sd[name + ".bias"] = (torch.randn(cout, generator=g, device=device) * 0.02).to(dtype)

    def norm(name, c):
        sd[name + ".weight"] = (1 + torch.randn(c, generator=g, device=device) * 0.05).to(dtype)
        sd[name + ".bias"] = (torch.randn(c, generator=g, device=device) * 0.05).to(dtype)

    def res(p, cin, cout):
        norm(p + "norm1", cin), conv(p + "conv1", cin, cout, 3), norm(p + "norm2", cout), conv(p + "conv2", cout, cout, 3)
        if cin != cout:
            conv(p + "nin_shortcut", cin, cout, 1)

    levels = len(ch_mult)
    block_in = ch * ch_mult[-1]
    conv("decoder.conv_in", z_channels, block_in, 3)
    res("decoder.mid.block_1.", block_in, block_in)
    norm("decoder.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        conv("decoder.mid.attn_1." + n, block_in, block_in, 1)
    res("decoder.mid.block_2.", block_in, block_in)
    for level in reversed(range(levels)):
        block_out = ch * ch_mult[level]
        for i in range(num_res_blocks + 1):
            res(f"decoder.up.{level}.block.{i}.", block_in, block_out)
            block_in = block_out
        if level != 0:
            conv(f"decoder.up.{level}.upsample.conv", block_in, block_in, 3)
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", block_in, out_ch, 3)
    return sd
