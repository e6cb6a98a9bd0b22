"""FLUX VAE decoder on the native kernels.

What the reference reaches through diffusers' AutoencoderKL after the sampler —
`latents = latents / vae.config.scaling_factor + vae.config.shift_factor; image = vae.decode(latents)`
(lakonlab/pipelines/arcflux_pipeline.py:531-534; training wrapper lakonlab/models/architecture/diffusers/pretrained.py:69-76)
— i.e. the black-forest-labs autoencoder's Decoder: conv_in, mid (ResnetBlock, single-head AttnBlock, ResnetBlock), four
up levels of three ResnetBlocks (+ nearest-2x Upsample conv on all but the last), GroupNorm + swish + conv_out.

Here: activations NHWC bf16; every 3x3 convolution is an implicit GEMM on the tcgen05 CTA-pair kernel (afb_conv3x3: 9
shifted TMA boxes per K sweep, residual add fused into the epilogue, TMA-store output); 1x1 convolutions, the attention
block's QK^T / PV products are afb_gemm launches; GroupNorm + swish, upsampling and the row softmax are streaming kernels
(csrc/vae.cu). State-dict names: the BFL ones (`decoder.up.<level>.block.<i>.conv1.weight`, ...); diffusers' AutoencoderKL
names are mapped by `diffusers_vae_to_bfl`. There is no PyTorch compute path in this class.
"""
from __future__ import annotations

import re
from typing import Dict, Optional

import torch

from . import _lib, ops
from ._lib import AfbError

BF16 = torch.bfloat16
SCALE_FACTOR = 0.3611     # vae.config.scaling_factor (FLUX.1)
SHIFT_FACTOR = 0.1159     # vae.config.shift_factor


def diffusers_vae_to_bfl(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """diffusers AutoencoderKL decoder keys -> BFL keys (the inverse of diffusers' `convert_ldm_vae_checkpoint`):
    mid_block.resnets.{0,1} -> mid.block_{1,2}; mid_block.attentions.0.{group_norm,to_q,to_k,to_v,to_out.0} ->
    mid.attn_1.{norm,q,k,v,proj_out} (Linear [C, C] -> 1x1 conv [C, C, 1, 1]); up_blocks.{i} -> up.{L-1-i};
    resnets.{j}.conv_shortcut -> block.{j}.nin_shortcut; upsamplers.0.conv -> upsample.conv; conv_norm_out -> norm_out."""
    if any(k.startswith("decoder.up.") or k.startswith("decoder.mid.block_1") for k in sd):
        return {k: v for k, v in sd.items() if k.startswith("decoder.")}
    levels = 1 + max(int(m.group(1)) for k in sd if (m := re.match(r"decoder\.up_blocks\.(\d+)\.", k)))
    out = {}
    for k, v in sd.items():
        if not k.startswith("decoder."):
            continue
        n = k
        n = n.replace("decoder.mid_block.resnets.0.", "decoder.mid.block_1.").replace("decoder.mid_block.resnets.1.", "decoder.mid.block_2.")
        n = n.replace("decoder.mid_block.attentions.0.group_norm.", "decoder.mid.attn_1.norm.")
        for a, b in (("to_q", "q"), ("to_k", "k"), ("to_v", "v"), ("to_out.0", "proj_out")):
            n = n.replace(f"decoder.mid_block.attentions.0.{a}.", f"decoder.mid.attn_1.{b}.")
        m = re.match(r"decoder\.up_blocks\.(\d+)\.(resnets|upsamplers)\.(\d+)\.(.*)", n)
        if m:
            lvl = levels - 1 - int(m.group(1))
            rest = m.group(4).replace("conv_shortcut.", "nin_shortcut.")
            n = f"decoder.up.{lvl}.block.{m.group(3)}.{rest}" if m.group(2) == "resnets" else f"decoder.up.{lvl}.upsample.{rest}"
        n = n.replace("decoder.conv_norm_out.", "decoder.norm_out.")
        if ".mid.attn_1." in n and n.endswith(".weight") and v.dim() == 2 and ".norm." not in n:
            v = v[:, :, None, None]
        out[n] = v
    return out


class FluxVAEDecoder:
    """`decode(latents)`: fp32 NCHW latents [B, 16, h, w] (un-packed, as the pipeline holds them after `_unpack_latents`) ->
    fp32 NCHW image [B, 3, 8h, 8w] in the VAE's output range (the reference's image_processor maps it to [0, 1] / PIL)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device="cuda", scale_factor: float = SCALE_FACTOR,
                 shift_factor: float = SHIFT_FACTOR, num_res_blocks: int = 2):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise AfbError("the VAE decoder needs a CUDA device (there is no CPU path)")
        sd = diffusers_vae_to_bfl(state_dict)
        self.scale_factor, self.shift_factor, self.num_res_blocks = float(scale_factor), float(shift_factor), num_res_blocks
        self.levels = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.up."))
        self.z_channels = sd["decoder.conv_in.weight"].shape[1]
        self.out_ch = sd["decoder.conv_out.weight"].shape[0]
        self.w: Dict[str, torch.Tensor] = {}
        dev = self.device

        def conv3(name, cin_pad=None, cout_pad=None):
            w = sd[name + ".weight"].to(dev)
            if w.shape[2:] != (3, 3):
                raise AfbError(f"{name}: expected a 3x3 convolution, got {tuple(w.shape)}")
            self.w[name + ".w"] = ops.pack_conv3x3_weight(w, cin_pad, cout_pad)
            b = sd[name + ".bias"].to(dev, BF16)
            if cout_pad and cout_pad > b.shape[0]:
                b = torch.cat([b, torch.zeros(cout_pad - b.shape[0], dtype=BF16, device=dev)])
            self.w[name + ".b"] = b.contiguous()

        def conv1(name):
            w = sd[name + ".weight"].to(dev, BF16)
            self.w[name + ".w"] = w.reshape(w.shape[0], w.shape[1]).contiguous()
            self.w[name + ".b"] = sd[name + ".bias"].to(dev, BF16).contiguous()

        def norm(name):
            self.w[name + ".g"] = sd[name + ".weight"].to(dev, torch.float32).contiguous()
            self.w[name + ".b"] = sd[name + ".bias"].to(dev, torch.float32).contiguous()

        def res(p):
            norm(p + "norm1"), conv3(p + "conv1"), norm(p + "norm2"), conv3(p + "conv2")
            if p + "nin_shortcut.weight" in sd:
                conv1(p + "nin_shortcut")

        self.z_pad = max(64, (self.z_channels + 63) // 64 * 64)
        conv3("decoder.conv_in", cin_pad=self.z_pad)
        res("decoder.mid.block_1.")
        norm("decoder.mid.attn_1.norm")
        a = "decoder.mid.attn_1."
        qkv_w = torch.cat([sd[a + n + ".weight"].to(dev, BF16).reshape(sd[a + n + ".weight"].shape[0], -1) for n in "qkv"], 0)
        self.w[a + "qkv.w"] = qkv_w.contiguous()
        self.w[a + "qkv.b"] = torch.cat([sd[a + n + ".bias"].to(dev, BF16) for n in "qkv"], 0).contiguous()
        conv1(a + "proj_out")
        res("decoder.mid.block_2.")
        for level in range(self.levels):
            for i in range(num_res_blocks + 1):
                res(f"decoder.up.{level}.block.{i}.")
            if level != 0:
                conv3(f"decoder.up.{level}.upsample.conv")
        norm("decoder.norm_out")
        conv3("decoder.conv_out", cout_pad=8)
        for k, v in self.w.items():
            if (v.data_ptr() & 15) != 0:
                raise AfbError(f"packed VAE tensor {k} is not 16-byte aligned")

    # -- blocks --------------------------------------------------------------------------------------------------
    def _gemm_rows(self, x4: torch.Tensor, w: torch.Tensor, bias, res4: Optional[torch.Tensor] = None) -> torch.Tensor:
        """1x1 convolution = GEMM over the pixel rows of an NHWC tensor (optionally + residual)."""
        n, h, wd, c = x4.shape
        out = torch.empty((n, h, wd, w.shape[0]), dtype=BF16, device=x4.device)
        rows = lambda t: t.reshape(1, n * h * wd, t.shape[-1])
        ops.gemm(rows(x4), w, rows(out), bias=bias, epilogue=_lib.AFB_EPI_BIAS_RES if res4 is not None else _lib.AFB_EPI_BIAS,
                 res=rows(res4) if res4 is not None else None)
        return out

    def _resnet(self, p: str, x: torch.Tensor) -> torch.Tensor:
        W = self.w
        h = ops.groupnorm(x, W[p + "norm1.g"], W[p + "norm1.b"], silu=True)
        h = ops.conv3x3(h, W[p + "conv1.w"], W[p + "conv1.b"])
        h = ops.groupnorm(h, W[p + "norm2.g"], W[p + "norm2.b"], silu=True, out=h)
        if p + "nin_shortcut.w" in W:
            x = self._gemm_rows(x, W[p + "nin_shortcut.w"], W[p + "nin_shortcut.b"])
        return ops.conv3x3(h, W[p + "conv2.w"], W[p + "conv2.b"], res=x)     # x + conv2(...) in the GEMM epilogue

    def _attn(self, p: str, x: torch.Tensor) -> torch.Tensor:
        W = self.w
        n, hh, ww, c = x.shape
        hw = hh * ww
        if hw % 64:
            raise AfbError(f"VAE attention block: {hh} x {ww} latent positions must be a multiple of 64 (K of the P V product)")
        h = ops.groupnorm(x, W[p + "norm.g"], W[p + "norm.b"], silu=False)
        qkv = self._gemm_rows(h, W[p + "qkv.w"], W[p + "qkv.b"]).reshape(n, hw, 3 * c)
        o = torch.empty((n, hw, c), dtype=BF16, device=x.device)
        scores = torch.empty((hw, hw), dtype=BF16, device=x.device)       # one image at a time (hw^2 bf16: 0.5 GB at 1024 px)
        scale = 1.0 / (c ** 0.5)
        for b in range(n):
            q, k, v = qkv[b, :, :c], qkv[b, :, c:2 * c], qkv[b, :, 2 * c:]
            ops.gemm(q.unsqueeze(0), k, scores.unsqueeze(0), alpha=scale)                  # S = scale * Q K^T
            ops.softmax_rows_(scores)
            ops.gemm(scores.unsqueeze(0), v, o[b].unsqueeze(0), transposed=True)           # O = P V (V read as [K, N])
        return self._gemm_rows(o.reshape(n, hh, ww, c), W[p + "proj_out.w"], W[p + "proj_out.b"], res4=x)

    @torch.no_grad()
    def decode(self, latents: torch.Tensor) -> torch.Tensor:
        if latents.dim() != 4 or latents.shape[1] != self.z_channels:
            raise AfbError(f"decode: latents must be [batch, {self.z_channels}, h, w], got {tuple(latents.shape)}")
        if not latents.is_cuda:
            raise AfbError("decode: latents must be a CUDA tensor (no CPU fallback exists)")
        W = self.w
        z = ops.vae_pre(latents.to(torch.float32).contiguous(), self.z_pad, self.scale_factor, self.shift_factor)
        h = ops.conv3x3(z, W["decoder.conv_in.w"], W["decoder.conv_in.b"])
        h = self._resnet("decoder.mid.block_1.", h)
        h = self._attn("decoder.mid.attn_1.", h)
        h = self._resnet("decoder.mid.block_2.", h)
        for level in reversed(range(self.levels)):
            for i in range(self.num_res_blocks + 1):
                h = self._resnet(f"decoder.up.{level}.block.{i}.", h)
            if level != 0:
                h = ops.upsample2x(h)
                h = ops.conv3x3(h, W[f"decoder.up.{level}.upsample.conv.w"], W[f"decoder.up.{level}.upsample.conv.b"])
        h = ops.groupnorm(h, W["decoder.norm_out.g"], W["decoder.norm_out.b"], silu=True, out=h)
        img = ops.conv3x3(h, W["decoder.conv_out.w"], W["decoder.conv_out.b"])
        return ops.vae_post(img, self.out_ch)

    __call__ = decode

    @staticmethod
    def flops(h: int, w: int, ch: int = 128, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2, z_pad: int = 64) -> float:
        """Algorithmic FLOPs of one decode of an [h, w] latent (2 per MAC; conv_in counted on the 16 real channels)."""
        c = ch * ch_mult[-1]
        px = h * w
        conv = lambda ci, co, p, k=9: 2.0 * k * ci * co * p
        total = conv(16, c, px) + 4 * conv(c, c, px) + 4 * conv(c, c, px, 1) + 4.0 * px * px * c
        cin = c
        for level in reversed(range(len(ch_mult))):
            co = ch * ch_mult[level]
            for i in range(num_res_blocks + 1):
                total += conv(cin, co, px) + conv(co, co, px) + (conv(cin, co, px, 1) if cin != co else 0)
                cin = co
            if level != 0:
                px *= 4
                total += conv(cin, cin, px)
        return total + conv(cin, 3, px)
