"""Cross-check vectors for the FLUX block arithmetic from an INDEPENDENT implementation.

The reference's FLUX blocks live in diffusers==0.35.1 (absent offline), so the oracle restates them.
This image ships `torchtitan.experiments.flux`, Meta's import of the original black-forest-labs/FLUX model code
(DoubleStreamBlock / SingleStreamBlock / EmbedND / LastLayer) — the implementation diffusers' FluxTransformer2DModel was
converted from (diffusers scripts/convert_flux_to_diffusers.py defines the weight mapping used below). Running it here on
seeded weights gives golden outputs that do not depend on anything written in this repo:

    python tools/make_golden_bfl.py            # writes tests/golden/bfl_flux_tiny.npz

tests/test_oracle_bfl.py maps the same seeded weights into the oracle (diffusers naming) and must reproduce the output.
It pins: modulation chunk order (shift, scale, gate), LayerNorm eps, per-head QK RMSNorm, the rotary embedding on the
joint [txt; img] sequence, SDPA scaling, text-first concatenation, the fused single-block linear1/linear2 split, GELU-tanh,
the sinusoidal timestep embedding and the final AdaLN (scale/shift order after the conversion swap).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from arcflow_b200.config import ArcFluxConfig  # noqa: E402
from arcflow_b200.synthetic import make_flux_state_dict, make_flux_teacher_extras  # noqa: E402


def tiny_cfg() -> ArcFluxConfig:
    return ArcFluxConfig(num_layers=2, num_single_layers=2, num_attention_heads=2, joint_attention_dim=192,
                         pooled_projection_dim=96, guidance_embeds=False, lora_rank=0)


def diffusers_state_dict(cfg: ArcFluxConfig, seed: int = 77):
    """Stock FLUX transformer weights under diffusers' names (the reference's `pipe.transformer.state_dict()`), fp32
    values that are bf16-representable."""
    sd = {k: v.float() for k, v in make_flux_state_dict(cfg, seed=seed).items()
          if not k.startswith(("proj_out_", "norm_out."))}
    sd.update({k: v.float() for k, v in make_flux_teacher_extras(cfg, seed=seed + 1).items()})
    return sd


def to_bfl(sd, cfg: ArcFluxConfig):
    """diffusers -> original FLUX names: the inverse of diffusers' convert_flux_to_diffusers.py."""
    D = cfg.inner_dim
    out = {}

    def lin(dst, src):
        out[dst + ".weight"] = sd[src + ".weight"]
        out[dst + ".bias"] = sd[src + ".bias"]

    def cat(dst, srcs):
        out[dst + ".weight"] = torch.cat([sd[s + ".weight"] for s in srcs], 0)
        out[dst + ".bias"] = torch.cat([sd[s + ".bias"] for s in srcs], 0)

    lin("img_in", "x_embedder")
    lin("txt_in", "context_embedder")
    lin("time_in.in_layer", "time_text_embed.timestep_embedder.linear_1")
    lin("time_in.out_layer", "time_text_embed.timestep_embedder.linear_2")
    lin("vector_in.in_layer", "time_text_embed.text_embedder.linear_1")
    lin("vector_in.out_layer", "time_text_embed.text_embedder.linear_2")
    for i in range(cfg.num_layers):
        s, d = f"transformer_blocks.{i}.", f"double_blocks.{i}."
        lin(d + "img_mod.lin", s + "norm1.linear")
        lin(d + "txt_mod.lin", s + "norm1_context.linear")
        cat(d + "img_attn.qkv", [s + "attn.to_q", s + "attn.to_k", s + "attn.to_v"])
        cat(d + "txt_attn.qkv", [s + "attn.add_q_proj", s + "attn.add_k_proj", s + "attn.add_v_proj"])
        out[d + "img_attn.norm.query_norm.weight"] = sd[s + "attn.norm_q.weight"]
        out[d + "img_attn.norm.key_norm.weight"] = sd[s + "attn.norm_k.weight"]
        out[d + "txt_attn.norm.query_norm.weight"] = sd[s + "attn.norm_added_q.weight"]
        out[d + "txt_attn.norm.key_norm.weight"] = sd[s + "attn.norm_added_k.weight"]
        lin(d + "img_attn.proj", s + "attn.to_out.0")
        lin(d + "txt_attn.proj", s + "attn.to_add_out")
        lin(d + "img_mlp.0", s + "ff.net.0.proj")
        lin(d + "img_mlp.2", s + "ff.net.2")
        lin(d + "txt_mlp.0", s + "ff_context.net.0.proj")
        lin(d + "txt_mlp.2", s + "ff_context.net.2")
    for i in range(cfg.num_single_layers):
        s, d = f"single_transformer_blocks.{i}.", f"single_blocks.{i}."
        lin(d + "modulation.lin", s + "norm.linear")
        cat(d + "linear1", [s + "attn.to_q", s + "attn.to_k", s + "attn.to_v", s + "proj_mlp"])
        lin(d + "linear2", s + "proj_out")
        out[d + "norm.query_norm.weight"] = sd[s + "attn.norm_q.weight"]
        out[d + "norm.key_norm.weight"] = sd[s + "attn.norm_k.weight"]
    lin("final_layer.linear", "proj_out")
    # diffusers stores AdaLayerNormContinuous as [scale; shift], the original as [shift; scale] (swap_scale_shift)
    w, b = sd["norm_out.linear.weight"], sd["norm_out.linear.bias"]
    out["final_layer.adaLN_modulation.1.weight"] = torch.cat([w[D:], w[:D]], 0)
    out["final_layer.adaLN_modulation.1.bias"] = torch.cat([b[D:], b[:D]], 0)
    return out


def make_inputs(cfg: ArcFluxConfig, batch: int, txt_len: int, gh: int, gw: int, seed: int = 5):
    g = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, gh * gw, cfg.in_channels, generator=g)
    txt = torch.randn(batch, txt_len, cfg.joint_attention_dim, generator=g) * 0.5
    y = torch.randn(batch, cfg.pooled_projection_dim, generator=g)
    t = torch.rand(batch, generator=g) * 0.9 + 0.05
    return img, txt, y, t


def run_bfl(sd, cfg: ArcFluxConfig, img, txt, y, t, gh: int, gw: int):
    from torchtitan.experiments.flux.model.args import FluxModelArgs
    from torchtitan.experiments.flux.model.model import FluxModel
    args = FluxModelArgs(in_channels=cfg.in_channels, out_channels=cfg.out_channels, vec_in_dim=cfg.pooled_projection_dim,
                         context_in_dim=cfg.joint_attention_dim, hidden_size=cfg.inner_dim, mlp_ratio=float(cfg.mlp_ratio),
                         num_heads=cfg.num_attention_heads, depth=cfg.num_layers, depth_single_blocks=cfg.num_single_layers,
                         axes_dim=tuple(cfg.axes_dims_rope), theta=10_000, qkv_bias=True)
    model = FluxModel(args).float().eval()
    missing, unexpected = model.load_state_dict(to_bfl(sd, cfg), strict=True), None
    for m in model.modules():  # the original FLUX RMSNorm uses eps 1e-6 (torch's nn.RMSNorm default is dtype eps)
        if isinstance(m, torch.nn.RMSNorm):
            m.eps = 1e-6
    B = img.shape[0]
    img_ids = torch.zeros(gh, gw, 3, dtype=torch.float64)
    img_ids[..., 1] += torch.arange(gh, dtype=torch.float64)[:, None]
    img_ids[..., 2] += torch.arange(gw, dtype=torch.float64)[None, :]
    img_ids = img_ids.reshape(1, gh * gw, 3).expand(B, -1, -1)
    txt_ids = torch.zeros(B, txt.shape[1], 3, dtype=torch.float64)
    with torch.no_grad():
        return model(img, img_ids, txt, txt_ids, t, y)


def main():
    cfg = tiny_cfg()
    gh, gw, St, B = 4, 6, 8, 2
    sd = diffusers_state_dict(cfg)
    img, txt, y, t = make_inputs(cfg, B, St, gh, gw)
    out = run_bfl(sd, cfg, img, txt, y, t, gh, gw)
    probe = sum(float(sd[k].double().abs().sum()) for k in sorted(sd))
    path = os.path.join(ROOT, "tests", "golden", "bfl_flux_tiny.npz")
    np.savez_compressed(path, img=img.numpy(), txt=txt.numpy(), y=y.numpy(), t=t.numpy(), out=out.numpy(),
                        grid=np.array([gh, gw]), weight_abs_sum=np.array([probe]))
    print("wrote", path, "out", tuple(out.shape), "mean-abs", float(out.abs().mean()))


if __name__ == "__main__":
    main()
