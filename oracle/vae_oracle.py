"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (arcflow_b200/, lakonlab/).

CPU restatement (plain torch functional ops, any float dtype) of the FLUX VAE DECODER the reference's pipeline runs after
the sampler: `image = self.vae.decode(latents / scaling_factor + shift_factor)` (lakonlab/pipelines/arcflux_pipeline.py:531-534;
training-side wrapper lakonlab/models/architecture/diffusers/pretrained.py:69-76). The arithmetic lives in diffusers'
AutoencoderKL (not vendored, not installable offline); FLUX's VAE is the original black-forest-labs autoencoder, whose
model code IS in this image as torchtitan.experiments.flux.model.autoencoder — tests/test_oracle_vae.py pins this
restatement on that module (same state dict, fp32, <= 1e-5) the way tests/test_oracle_bfl.py pins the transformer.

State-dict names are the BFL ones (`decoder.conv_in`, `decoder.mid.block_1`, `decoder.mid.attn_1.{norm,q,k,v,proj_out}`,
`decoder.up.<level>.block.<i>.{norm1,conv1,norm2,conv2,nin_shortcut}`, `decoder.up.<level>.upsample.conv`,
`decoder.norm_out`, `decoder.conv_out`); the product side (arcflow_b200/vae.py::diffusers_vae_to_bfl) maps diffusers' names onto them.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

SCALE_FACTOR = 0.3611     # vae.config.scaling_factor of FLUX.1 (AutoEncoderParams.scale_factor)
SHIFT_FACTOR = 0.1159     # vae.config.shift_factor


def _gn(sd, name: str, x: Tensor, dtype) -> Tensor:
    return F.group_norm(x, 32, sd[name + ".weight"].to(dtype), sd[name + ".bias"].to(dtype), eps=1e-6)


def _conv(sd, name: str, x: Tensor, dtype, padding: int) -> Tensor:
    return F.conv2d(x, sd[name + ".weight"].to(dtype), sd[name + ".bias"].to(dtype), padding=padding)


def _swish(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def resnet_block(sd, p: str, x: Tensor, dtype) -> Tensor:
    """ResnetBlock.forward: GN -> swish -> conv3x3 -> GN -> swish -> conv3x3, + (1x1 shortcut of) x."""
    h = _conv(sd, p + "conv1", _swish(_gn(sd, p + "norm1", x, dtype)), dtype, 1)
    h = _conv(sd, p + "conv2", _swish(_gn(sd, p + "norm2", h, dtype)), dtype, 1)
    if p + "nin_shortcut.weight" in sd:
        x = _conv(sd, p + "nin_shortcut", x, dtype, 0)
    return x + h


def attn_block(sd, p: str, x: Tensor, dtype) -> Tensor:
    """AttnBlock.forward: GN -> 1x1 q, k, v -> single-head SDPA over the H*W positions (scale 1/sqrt(C)) -> 1x1 proj, + x."""
    h = _gn(sd, p + "norm", x, dtype)
    q, k, v = (_conv(sd, p + n, h, dtype, 0) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).transpose(1, 2)[:, None] for t in (q, k, v))     # b 1 (h w) c
    o = F.scaled_dot_product_attention(q, k, v)
    o = o[:, 0].transpose(1, 2).reshape(b, c, hh, ww)
    return x + _conv(sd, p + "proj_out", o, dtype, 0)


def decoder_levels(sd) -> int:
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.up."))


def vae_decode(sd: Dict[str, Tensor], latents: Tensor, dtype=torch.float32, scale_factor: float = SCALE_FACTOR,
               shift_factor: float = SHIFT_FACTOR, num_res_blocks: int = 2) -> Tensor:
    """AutoEncoder.decode: z / scale + shift -> Decoder.forward. latents [B, 16, h, w] -> image [B, 3, 8h, 8w]."""
    z = (latents.to(torch.float32) / scale_factor + shift_factor).to(dtype)
    h = _conv(sd, "decoder.conv_in", z, dtype, 1)
    h = resnet_block(sd, "decoder.mid.block_1.", h, dtype)
    h = attn_block(sd, "decoder.mid.attn_1.", h, dtype)
    h = resnet_block(sd, "decoder.mid.block_2.", h, dtype)
    for level in reversed(range(decoder_levels(sd))):
        for i in range(num_res_blocks + 1):
            h = resnet_block(sd, f"decoder.up.{level}.block.{i}.", h, dtype)
        if level != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up.{level}.upsample.conv", h, dtype, 1)
    h = _swish(_gn(sd, "decoder.norm_out", h, dtype))
    return _conv(sd, "decoder.conv_out", h, dtype, 1)


# seeded synthetic decoder weights of the BFL layout live with the other synthetic factories
from arcflow_b200.synthetic import make_vae_decoder_state_dict  # noqa: E402,F401
