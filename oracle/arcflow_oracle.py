"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (arcflow_b200/, lakonlab/).

CPU restatement (plain torch, any float dtype) of the reference's denoising hot path, used ONLY by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs as the checker and
the CPU baseline.

Parity status
  * Sampler side (schedule, token<->image layout, ArcFlowPolicy, momentum_integration): PINNED — checked in
    tests/test_oracle_golden.py against tests/golden/reference_sampler.npz, which tools/make_golden.py
    produced by executing the reference's own functions from /root/reference.
  * Transformer block arithmetic: lives in diffusers==0.35.1 / peft==0.17.0 (requirements.txt:4-5 of the reference),
    which are not vendored under /root/reference and not installable offline, so the reference's own classes cannot
    pin it (PARITY UNPINNED against diffusers itself). It is restated from their published semantics as recorded in
    SURVEY.md Appendix A, anchored on the reference's own call sites (cited per function below), and — for FLUX —
    CROSS-CHECKED against an independent implementation: the original black-forest-labs FLUX model code shipped in this
    image as torchtitan.experiments.flux (tests/test_oracle_bfl.py, tests/golden/bfl_flux_tiny.npz from
    tools/make_golden_bfl.py, weights mapped with diffusers' published FLUX conversion). Embedders, double / single
    blocks, RoPE, QK-RMSNorm, modulation order and the final AdaLN agree to 2e-5. Qwen-Image has no such second source.

Every function cites the reference file:line it follows (paths relative to the upstream repo root).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# =================================================================================================
# schedule — lakonlab/pipelines/arcflux_pipeline.py:34-70, :413-431, :455-493
# =================================================================================================
def retrieve_raw_timesteps(num_inference_steps: int, total_substeps: int, timestep_ratio: float):
    base_segment_size = 1 / (num_inference_steps - 1 + timestep_ratio)
    raw_timesteps, num_inference_substeps = [], []
    _raw_t = 1.0
    for i in range(num_inference_steps):
        segment_size = base_segment_size if i < num_inference_steps - 1 else base_segment_size * timestep_ratio
        n = max(round(segment_size * total_substeps), 1)
        num_inference_substeps.append(n)
        raw_timesteps.extend(np.linspace(_raw_t, _raw_t - segment_size, n, endpoint=False).clip(min=0.0).tolist())
        _raw_t = _raw_t - segment_size
    return raw_timesteps, num_inference_substeps, sum(num_inference_substeps)


def scheduler_timesteps(raw_timesteps: Sequence[float], shift: float = 3.2) -> Tensor:
    """FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=raw) with use_dynamic_shifting=False
    (reference configures it at inference_flux.py:14): fp32 sigmas, shifted, timesteps = sigma * 1000."""
    sigmas = np.array(raw_timesteps).astype(np.float32)
    sigmas = shift * sigmas / (1 + (shift - 1) * sigmas)
    return torch.from_numpy(sigmas).to(torch.float32) * 1000.0


# =================================================================================================
# token <-> image layout — arcflux_pipeline.py:135-193
# =================================================================================================
def unpack_mp(mp: Dict[str, Tensor], grid_h: int, grid_w: int, num_gaussians: int, patch: int = 2):
    bs = mp["means"].size(0)
    k, s = num_gaussians, patch
    c = mp["means"].shape[-1] if mp["means"].dim() == 4 else mp["means"].shape[-1] // k
    out = {}
    out["means"] = mp["means"].reshape(bs, grid_h, grid_w, k, c // (s * s), s, s).permute(
        0, 3, 4, 1, 5, 2, 6).reshape(bs, k, c // (s * s), grid_h * s, grid_w * s)
    out["logweights"] = mp["logweights"].reshape(bs, grid_h, grid_w, k, 1, s, s).permute(
        0, 3, 4, 1, 5, 2, 6).reshape(bs, k, 1, grid_h * s, grid_w * s)
    out["loggammas"] = mp["loggammas"].reshape(bs, grid_h, grid_w, k - 1, 1, s, s).permute(
        0, 3, 4, 1, 5, 2, 6).reshape(bs, k - 1, 1, grid_h * s, grid_w * s)
    return out


def unpack_latents(latents: Tensor, grid_h: int, grid_w: int, patch: int = 2) -> Tensor:
    b, _, ch = latents.shape
    x = latents.view(b, grid_h, grid_w, ch // (patch * patch), patch, patch).permute(0, 3, 1, 4, 2, 5)
    return x.reshape(b, ch // (patch * patch), grid_h * patch, grid_w * patch)


def pack_latents(latents: Tensor, patch: int = 2) -> Tensor:
    b, c, hh, ww = latents.shape
    x = latents.view(b, c, hh // patch, patch, ww // patch, patch).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(b, (hh // patch) * (ww // patch), c * patch * patch)


# =================================================================================================
# policy + analytic integration — policies/arcflow.py:25-76, arcflux_pipeline.py:195-249
# =================================================================================================
def policy_velocity(mp: Dict[str, Tensor], sigma_t_src, sigma_t) -> Tensor:
    """ArcFlowPolicy.velocity (lakonlab/models/diffusions/policies/arcflow.py:52-76)."""
    means, log_gammas, logweights = mp["means"], mp["loggammas"], mp["logweights"]
    weights = torch.softmax(logweights, dim=1)
    dt_past = torch.as_tensor(sigma_t_src - sigma_t, dtype=means.dtype)
    decay = torch.exp(log_gammas * dt_past)
    decay = torch.cat([decay.new_ones((decay.shape[0], 1, *decay.shape[2:])), decay], dim=1)
    return (means * decay * weights).sum(dim=1)


def momentum_integration(mp: Dict[str, Tensor], x_t_start: Tensor, sigma_t_src: float,
                         sigma_t_start: float, sigma_t_end: float, eps: float = 1e-4) -> Tensor:
    """x_end = x_start - sum_k w_k mu_k e^{lam_k (s_src - s_start)} dt phi(lam_k dt), lam_0 = 0
    (arcflux_pipeline.py:195-249; identical math in lakonlab/models/diffusions/arcflow.py:28-79)."""
    means, log_gammas, logweights = mp["means"], mp["loggammas"], mp["logweights"]
    dt_past = torch.as_tensor(sigma_t_src, dtype=means.dtype) - torch.as_tensor(sigma_t_start, dtype=means.dtype)
    dt_step = torch.as_tensor(sigma_t_start, dtype=means.dtype) - torch.as_tensor(sigma_t_end, dtype=means.dtype)
    decay = torch.exp(log_gammas * dt_past)
    decay = torch.cat([decay.new_ones((decay.shape[0], 1, *decay.shape[2:])), decay], dim=1)
    v_at_a = means * decay
    z = log_gammas * dt_step
    sign = torch.sign(z)
    sign[sign == 0] = 1
    z = sign * torch.clamp(z.abs(), min=eps)
    step = torch.expm1(z) / z
    step = torch.cat([step.new_ones((step.shape[0], 1, *step.shape[2:])), step], dim=1)
    weights = torch.softmax(logweights, dim=1)
    return x_t_start - (weights * (v_at_a * dt_step * step)).sum(dim=1)


# =================================================================================================
# FLUX transformer — arcflux.py:134-257 over diffusers blocks
# =================================================================================================
# LoRA input dropout (peft `lora_dropout`, train mode only; configs/flux/arcflux_2nfe_k16.py:40-48 p = 0.05). torch's
# Philox stream cannot be reproduced by another implementation, so the mask is DEFINED here as a counter-based hash of
# (seed, layer id, element index) — the CUDA kernels use the same definition and are tested against this restatement.
# Set to dict(p=..., seed=..., num_double=...) to enable (tests / training oracle only); None = eval mode.
LORA_DROPOUT = None


def _lowbias32(h):
    import numpy as np
    m = np.uint64(0xFFFFFFFF)
    h = h ^ (h >> np.uint64(16))
    h = (h * np.uint64(0x7FEB352D)) & m
    h = h ^ (h >> np.uint64(15))
    h = (h * np.uint64(0x846CA68B)) & m
    return h ^ (h >> np.uint64(16))


def lora_layer_id(name: str, num_double: int) -> int:
    """Mask-stream id of a LoRA target: 4 slots per block (double blocks first), the timestep embedder past all blocks."""
    if name.endswith("timestep_embedder.linear_1"):
        return 0xFFFF0
    if name.endswith("timestep_embedder.linear_2"):
        return 0xFFFF1
    parts = name.split(".")
    i = int(parts[1])
    tail = ".".join(parts[2:])
    if parts[0] == "single_transformer_blocks":
        return 4 * (num_double + i) + {"proj_mlp": 0, "proj_out": 1}[tail]
    slot = {"ff.net.0.proj": 0, "ff.net.2": 1, "ff_context.net.0.proj": 2, "ff_context.net.2": 3,
            "img_mlp.net.0.proj": 0, "img_mlp.net.2": 1, "txt_mlp.net.0.proj": 2, "txt_mlp.net.2": 3}[tail]
    return 4 * i + slot


def lora_dropout_mask(seed: int, layer_id: int, shape, p: float) -> Tensor:
    """keep[idx] = 16 bits of hash(key(seed, layer), idx >> 1) >= round(p * 2^16) — the low half for even idx, the high half
    for odd idx; idx = row-major index into `shape` (bool tensor). One 32-bit hash serves a pair of elements."""
    import numpy as np
    m = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        inner = (np.uint64(seed >> 32) + np.uint64(layer_id) * np.uint64(0x632BE5AB) + np.uint64(1)) & m
        key = _lowbias32((np.uint64(seed & 0xFFFFFFFF) ^ _lowbias32(inner)) & m)
        n = 1
        for d in shape:
            n *= int(d)
        idx = np.arange(n, dtype=np.uint64)
        pair = idx >> np.uint64(1)
        hi = ((pair >> np.uint64(32)) * np.uint64(0x9E3779B1)) & m
        h = _lowbias32(((pair & m) ^ key ^ hi) & m)
        bits = np.where((idx & np.uint64(1)) != 0, h >> np.uint64(16), h & np.uint64(0xFFFF))
    keep = bits >= np.uint64(int(float(np.float32(p)) * 65536.0 + 0.5))
    return torch.from_numpy(keep).reshape(tuple(shape))


def _lin(sd, name: str, x: Tensor, dtype, lora_scale: float = 1.0) -> Tensor:
    """nn.Linear, plus the peft LoRA branch when `<name>.lora_A/B.weight` exist (Appendix A.6):
    result = base(x) + lora_B(lora_A(dropout(x))) * scaling, scaling = alpha / r = 1 (arcflux.py:295-301)."""
    y = F.linear(x, sd[name + ".weight"].to(dtype), sd[name + ".bias"].to(dtype) if name + ".bias" in sd else None)
    if name + ".lora_A.weight" in sd:
        a, b = sd[name + ".lora_A.weight"].to(dtype), sd[name + ".lora_B.weight"].to(dtype)
        xin = x
        if LORA_DROPOUT is not None and LORA_DROPOUT["p"] > 0:
            d = LORA_DROPOUT
            keep = lora_dropout_mask(d["seed"], lora_layer_id(name, d["num_double"]), x.shape, d["p"])
            xin = x * keep.to(x.dtype) * (1.0 / (1.0 - d["p"]))
        y = y + F.linear(F.linear(xin, a), b) * lora_scale
    return y


def timestep_proj(t: Tensor, dim: int = 256, scale: float = 1.0) -> Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0) (Appendix A.5)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t.float()[:, None] * torch.exp(exponent)[None, :] * scale
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def rms_norm(x: Tensor, weight: Tensor, eps: float = 1e-6) -> Tensor:
    """diffusers RMSNorm.forward (Appendix A.5): fp32 variance; cast to weight dtype if half; * weight."""
    var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if weight.dtype in (torch.float16, torch.bfloat16):
        y = y.to(weight.dtype)
    return y * weight


def flux_rope(txt_len: int, grid_h: int, grid_w: int, axes=(16, 56, 56), theta: float = 10000.0):
    """FluxPosEmbed on cat(txt_ids, img_ids) (arcflux.py:171-172, ids built at :360-373, :428)."""
    ids = torch.zeros(txt_len + grid_h * grid_w, 3, dtype=torch.float64)
    img = torch.zeros(grid_h, grid_w, 3, dtype=torch.float64)
    img[..., 1] += torch.arange(grid_h, dtype=torch.float64)[:, None]
    img[..., 2] += torch.arange(grid_w, dtype=torch.float64)[None, :]
    ids[txt_len:] = img.reshape(-1, 3)
    cos, sin = [], []
    for i, d in enumerate(axes):
        freqs = 1.0 / (theta ** (torch.arange(0, d, 2, dtype=torch.float64)[: d // 2] / d))
        ang = torch.outer(ids[:, i], freqs)
        cos.append(ang.cos().repeat_interleave(2, dim=1).float())
        sin.append(ang.sin().repeat_interleave(2, dim=1).float())
    return torch.cat(cos, -1), torch.cat(sin, -1)


_ROPE_CACHE: Dict = {}


def _cached_rope(key, make):
    """The tables depend on the shape only; the reference recomputes them every forward (arcflux.py:171-173) — caching them
    here only makes the baseline arms that run this oracle (bench.py) faster, never different."""
    if key not in _ROPE_CACHE:
        if len(_ROPE_CACHE) > 16:
            _ROPE_CACHE.clear()
        _ROPE_CACHE[key] = make()
    return _ROPE_CACHE[key]


def apply_rope(x: Tensor, cos: Tensor, sin: Tensor) -> Tensor:
    """diffusers apply_rotary_emb(use_real=True, unbind_dim=-1) on x [B, S, H, 128] (Appendix A.3)."""
    cos, sin = cos[None, :, None, :], sin[None, :, None, :]
    xr, xi = x.reshape(*x.shape[:-1], -1, 2).unbind(-1)
    x_rot = torch.stack([-xi, xr], dim=-1).flatten(3)
    return (x.float() * cos + x_rot.float() * sin).to(x.dtype)


def _attention(q: Tensor, k: Tensor, v: Tensor) -> Tensor:
    """[B, S, H, d] -> [B, S, H*d]; SDPA, scale 1/sqrt(d), no mask, no dropout (Appendix A.1)."""
    o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
    return o.transpose(1, 2).flatten(2)


def _ln(x: Tensor) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def flux_double_block(sd, p: str, x: Tensor, c: Tensor, temb: Tensor, rope, heads: int, dtype, ls: float):
    """diffusers FluxTransformerBlock.forward (Appendix A.1), called at arcflux.py:191-197."""
    B = x.shape[0]
    emb = _lin(sd, p + "norm1.linear", F.silu(temb), dtype)
    sh_msa, sc_msa, g_msa, sh_mlp, sc_mlp, g_mlp = emb.chunk(6, dim=1)
    n = _ln(x) * (1 + sc_msa[:, None]) + sh_msa[:, None]
    emb_c = _lin(sd, p + "norm1_context.linear", F.silu(temb), dtype)
    csh_msa, csc_msa, cg_msa, csh_mlp, csc_mlp, cg_mlp = emb_c.chunk(6, dim=1)
    nc = _ln(c) * (1 + csc_msa[:, None]) + csh_msa[:, None]

    a = p + "attn."
    q = _lin(sd, a + "to_q", n, dtype).unflatten(-1, (heads, -1))
    k = _lin(sd, a + "to_k", n, dtype).unflatten(-1, (heads, -1))
    v = _lin(sd, a + "to_v", n, dtype).unflatten(-1, (heads, -1))
    q = rms_norm(q, sd[a + "norm_q.weight"].to(dtype))
    k = rms_norm(k, sd[a + "norm_k.weight"].to(dtype))
    eq = _lin(sd, a + "add_q_proj", nc, dtype).unflatten(-1, (heads, -1))
    ek = _lin(sd, a + "add_k_proj", nc, dtype).unflatten(-1, (heads, -1))
    ev = _lin(sd, a + "add_v_proj", nc, dtype).unflatten(-1, (heads, -1))
    eq = rms_norm(eq, sd[a + "norm_added_q.weight"].to(dtype))
    ek = rms_norm(ek, sd[a + "norm_added_k.weight"].to(dtype))
    q, k, v = torch.cat([eq, q], 1), torch.cat([ek, k], 1), torch.cat([ev, v], 1)   # text first
    q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    o = _attention(q, k, v).to(q.dtype)
    St = c.shape[1]
    oc, ox = o[:, :St], o[:, St:]
    ox = _lin(sd, a + "to_out.0", ox, dtype)
    oc = _lin(sd, a + "to_add_out", oc, dtype)

    x = x + g_msa[:, None] * ox
    y = _ln(x) * (1 + sc_mlp[:, None]) + sh_mlp[:, None]
    ff = _lin(sd, p + "ff.net.2", F.gelu(_lin(sd, p + "ff.net.0.proj", y, dtype, ls), approximate="tanh"), dtype, ls)
    x = x + g_mlp[:, None] * ff
    c = c + cg_msa[:, None] * oc
    yc = _ln(c) * (1 + csc_mlp[:, None]) + csh_mlp[:, None]
    ffc = _lin(sd, p + "ff_context.net.2",
               F.gelu(_lin(sd, p + "ff_context.net.0.proj", yc, dtype, ls), approximate="tanh"), dtype, ls)
    c = c + cg_mlp[:, None] * ffc
    return c, x


def flux_single_block(sd, p: str, x: Tensor, c: Tensor, temb: Tensor, rope, heads: int, dtype, ls: float):
    """diffusers FluxSingleTransformerBlock.forward (Appendix A.2), called at arcflux.py:224-230."""
    St = c.shape[1]
    h = torch.cat([c, x], dim=1)
    res = h
    emb = _lin(sd, p + "norm.linear", F.silu(temb), dtype)
    sh, sc, gate = emb.chunk(3, dim=1)
    n = _ln(h) * (1 + sc[:, None]) + sh[:, None]
    m = F.gelu(_lin(sd, p + "proj_mlp", n, dtype, ls), approximate="tanh")
    a = p + "attn."
    q = rms_norm(_lin(sd, a + "to_q", n, dtype).unflatten(-1, (heads, -1)), sd[a + "norm_q.weight"].to(dtype))
    k = rms_norm(_lin(sd, a + "to_k", n, dtype).unflatten(-1, (heads, -1)), sd[a + "norm_k.weight"].to(dtype))
    v = _lin(sd, a + "to_v", n, dtype).unflatten(-1, (heads, -1))
    q, k = apply_rope(q, *rope), apply_rope(k, *rope)
    o = _attention(q, k, v).to(q.dtype)
    h = res + gate[:, None] * _lin(sd, p + "proj_out", torch.cat([o, m], dim=2), dtype, ls)
    return h[:, :St], h[:, St:]


def flux_trunk(sd: Dict[str, Tensor], cfg, hidden_states: Tensor, encoder_hidden_states: Tensor,
               pooled_projections: Tensor, timestep: Tensor, guidance: Optional[Tensor],
               grid_hw: Sequence[int], dtype=torch.float32, lora_scale: float = 1.0, bf16_quirks: bool = True):
    """Embedders + 19 double + 38 single blocks (arcflux.py:158-230). Returns (image hidden states, temb)."""
    heads = cfg.num_attention_heads
    x = _lin(sd, "x_embedder", hidden_states.to(dtype), dtype)
    qd = torch.bfloat16 if bf16_quirks else dtype
    t = (timestep.to(qd) * 1000)
    te = "time_text_embed."
    pooled = pooled_projections.to(dtype)

    def mlp(prefix: str, v: Tensor) -> Tensor:
        return _lin(sd, prefix + ".linear_2", F.silu(_lin(sd, prefix + ".linear_1", v, dtype, lora_scale)), dtype, lora_scale)

    temb = mlp(te + "timestep_embedder", timestep_proj(t).to(dtype))
    if guidance is not None:
        g = (guidance.to(qd) * 1000)
        temb = temb + mlp(te + "guidance_embedder", timestep_proj(g).to(dtype))
    temb = temb + mlp(te + "text_embedder", pooled)
    c = _lin(sd, "context_embedder", encoder_hidden_states.to(dtype), dtype)

    cos, sin = _cached_rope(("flux", c.shape[1], tuple(grid_hw), tuple(cfg.axes_dims_rope), str(x.device)),
                            lambda: tuple(t.to(x.device) for t in flux_rope(c.shape[1], grid_hw[0], grid_hw[1], cfg.axes_dims_rope)))
    if bf16_quirks:
        cos, sin = cos.bfloat16().float(), sin.bfloat16().float()
    rope = (cos.to(dtype) if dtype != torch.bfloat16 else cos.bfloat16(),
            sin.to(dtype) if dtype != torch.bfloat16 else sin.bfloat16())

    for i in range(cfg.num_layers):
        c, x = flux_double_block(sd, f"transformer_blocks.{i}.", x, c, temb, rope, heads, dtype, lora_scale)
    for i in range(cfg.num_single_layers):
        c, x = flux_single_block(sd, f"single_transformer_blocks.{i}.", x, c, temb, rope, heads, dtype, lora_scale)
    return x, temb


def flux_forward(sd: Dict[str, Tensor], cfg, hidden_states: Tensor, encoder_hidden_states: Tensor,
                 pooled_projections: Tensor, timestep: Tensor, guidance: Optional[Tensor],
                 grid_hw: Sequence[int], dtype=torch.float32, lora_scale: float = 1.0,
                 bf16_quirks: bool = True, return_raw: bool = False) -> Dict[str, Tensor]:
    """_ArcFluxTransformer2DModel.forward (arcflux.py:134-257). return_raw adds 'raw': the three head Linears'
    outputs concatenated [B, S_i, K*C + K*L + (K-1)*L] before the log-softmax (what the engine's head GEMM emits).

    `timestep` is sigma in [0, 1] as the pipeline passes it (arcflux_pipeline.py:472), `guidance` the
    guidance scale. `dtype` is the compute dtype (float32/float64 = oracle; bfloat16 = the reference's
    own numerics on CPU). bf16_quirks reproduces the two input roundings that exist regardless of
    compute dtype in the reference's bf16 deployment: `timestep.to(bf16) * 1000` (:160-162) and the
    RoPE tables cast to the hidden dtype (:173) — SURVEY.md Appendix A.8.
    """
    x, temb = flux_trunk(sd, cfg, hidden_states, encoder_hidden_states, pooled_projections, timestep, guidance,
                         grid_hw, dtype=dtype, lora_scale=lora_scale, bf16_quirks=bf16_quirks)
    # AdaLayerNormContinuous: scale first, then shift (Appendix A.5)
    emb = _lin(sd, "norm_out.linear", F.silu(temb).to(x.dtype), dtype)
    scale, shift = emb.chunk(2, dim=1)
    x = _ln(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    bs, seq, _ = x.shape
    K, C, L = cfg.num_gaussians, cfg.out_channels, cfg.logweights_channels
    means = _lin(sd, "proj_out_means", x, dtype)
    logits = _lin(sd, "proj_out_logweights", x, dtype)
    gam = _lin(sd, "proj_out_loggamma", x, dtype)
    out = dict(means=means.reshape(bs, seq, K, C), logweights=logits.reshape(bs, seq, K, L).log_softmax(dim=-2),
               loggammas=gam.reshape(bs, seq, K - 1, L))
    if return_raw:
        out["raw"] = torch.cat([means, logits, gam], dim=-1)
    return out


# =================================================================================================
# denoising loop — ArcFluxPipeline.__call__, arcflux_pipeline.py:453-524
# =================================================================================================
def flux_denoise(sd, cfg, latents: Tensor, prompt_embeds: Tensor, pooled: Tensor, grid_hw: Sequence[int],
                 num_inference_steps: int = 2, total_substeps: int = 128, timestep_ratio: float = 1.0,
                 shift: float = 3.2, guidance_scale: float = 3.5, dtype=torch.float32, eps: float = 1e-4,
                 net_dtype=torch.bfloat16, return_trace: bool = False):
    """latents: fp32 packed tokens [B, S_i, 64]. The sampler state stays fp32 (:407, :487); the network
    sees `latents.to(net_dtype)` (:471, net_dtype = transformer.dtype = bf16 in the reference) and its
    outputs are cast to fp32 after being emitted in `net_dtype` (here: computed in `dtype`, then rounded
    through net_dtype to mirror the bf16 emission)."""
    gh, gw = grid_hw
    raw, substeps, total = retrieve_raw_timesteps(num_inference_steps, total_substeps, timestep_ratio)
    timesteps = scheduler_timesteps(raw, shift)
    assert len(timesteps) == total
    B = latents.shape[0]
    dev = latents.device
    guidance = torch.full([B], guidance_scale, dtype=torch.float32, device=dev) if cfg.guidance_embeds else None
    tid = 0
    trace = []
    latents = latents.to(torch.float32)
    for i in range(num_inference_steps):
        t_src = timesteps[tid]
        sigma_src = t_src / 1000.0
        out = flux_forward(sd, cfg, latents.to(net_dtype), prompt_embeds, pooled,
                           (t_src.expand(B) / 1000).to(dev), guidance, grid_hw, dtype=dtype)
        out = {k: v.to(net_dtype).to(torch.float32) for k, v in out.items()}
        x_img = unpack_latents(latents, gh, gw)
        mp = unpack_mp(out, gh, gw, cfg.num_gaussians)
        tid += substeps[i]
        t_end = timesteps[tid] if tid < len(timesteps) else torch.tensor(0.0)
        x_img = momentum_integration(mp, x_img, float(sigma_src), float(sigma_src), float(t_end / 1000.0), eps)
        latents = pack_latents(x_img)
        if return_trace:
            trace.append(dict(sigma_src=float(sigma_src), sigma_end=float(t_end / 1000.0), out=out,
                              latents=latents.clone()))
    return (latents, trace) if return_trace else latents


# =================================================================================================
# Qwen-Image transformer — arcqwen.py:106-174 over diffusers blocks  [PARITY UNPINNED]
# =================================================================================================
class QwenEmbedRope:
    """diffusers QwenEmbedRope(theta=10000, axes_dim, scale_rope=True) (constructed at arcqwen.py:46)."""

    def __init__(self, theta: int = 10000, axes_dim=(16, 56, 56), scale_rope: bool = True):
        self.theta, self.axes_dim, self.scale_rope = theta, list(axes_dim), scale_rope
        pos_index = torch.arange(4096)
        neg_index = torch.arange(4096).flip(0) * -1 - 1
        self.pos_freqs = torch.cat([self.rope_params(pos_index, d, theta) for d in self.axes_dim], dim=1)
        self.neg_freqs = torch.cat([self.rope_params(neg_index, d, theta) for d in self.axes_dim], dim=1)

    @staticmethod
    def rope_params(index, dim, theta=10000):
        freqs = torch.outer(index, 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))
        return torch.polar(torch.ones_like(freqs), freqs)

    def __call__(self, frame: int, height: int, width: int, max_txt_len: int):
        split = [x // 2 for x in self.axes_dim]
        fp, fn = self.pos_freqs.split(split, dim=1), self.neg_freqs.split(split, dim=1)
        f_frame = fp[0][0:frame].view(frame, 1, 1, -1).expand(frame, height, width, -1)
        if self.scale_rope:
            f_h = torch.cat([fn[1][-(height - height // 2):], fp[1][: height // 2]], dim=0)
            f_w = torch.cat([fn[2][-(width - width // 2):], fp[2][: width // 2]], dim=0)
            max_vid_index = max(height // 2, width // 2)
        else:
            f_h, f_w = fp[1][:height], fp[2][:width]
            max_vid_index = max(height, width)
        f_h = f_h.view(1, height, 1, -1).expand(frame, height, width, -1)
        f_w = f_w.view(1, 1, width, -1).expand(frame, height, width, -1)
        vid = torch.cat([f_frame, f_h, f_w], dim=-1).reshape(frame * height * width, -1)
        txt = self.pos_freqs[max_vid_index: max_vid_index + max_txt_len]
        return vid, txt


def apply_rope_qwen(x: Tensor, freqs_cis: Tensor) -> Tensor:
    """diffusers apply_rotary_emb_qwen(use_real=False): complex multiply on adjacent pairs, x [B, S, H, 128]."""
    xc = torch.view_as_complex(x.float().reshape(*x.shape[:-1], -1, 2))
    out = torch.view_as_real(xc * freqs_cis.unsqueeze(1)).flatten(3)
    return out.type_as(x)


def qwen_block(sd, p: str, x: Tensor, c: Tensor, temb: Tensor, rope, heads: int, dtype, ls: float):
    """diffusers QwenImageTransformerBlock.forward + QwenDoubleStreamAttnProcessor2_0 (Appendix A.4),
    called at arcqwen.py:147-155. The text mask is not used inside attention (0.35.1)."""
    img_freqs, txt_freqs = rope
    img_mod = _lin(sd, p + "img_mod.1", F.silu(temb), dtype)
    txt_mod = _lin(sd, p + "txt_mod.1", F.silu(temb), dtype)
    (i_sh1, i_sc1, i_g1), (i_sh2, i_sc2, i_g2) = [m.chunk(3, dim=-1) for m in img_mod.chunk(2, dim=-1)]
    (t_sh1, t_sc1, t_g1), (t_sh2, t_sc2, t_g2) = [m.chunk(3, dim=-1) for m in txt_mod.chunk(2, dim=-1)]
    xm = _ln(x) * (1 + i_sc1[:, None]) + i_sh1[:, None]
    cm = _ln(c) * (1 + t_sc1[:, None]) + t_sh1[:, None]
    a = p + "attn."
    iq = rms_norm(_lin(sd, a + "to_q", xm, dtype).unflatten(-1, (heads, -1)), sd[a + "norm_q.weight"].to(dtype))
    ik = rms_norm(_lin(sd, a + "to_k", xm, dtype).unflatten(-1, (heads, -1)), sd[a + "norm_k.weight"].to(dtype))
    iv = _lin(sd, a + "to_v", xm, dtype).unflatten(-1, (heads, -1))
    tq = rms_norm(_lin(sd, a + "add_q_proj", cm, dtype).unflatten(-1, (heads, -1)), sd[a + "norm_added_q.weight"].to(dtype))
    tk = rms_norm(_lin(sd, a + "add_k_proj", cm, dtype).unflatten(-1, (heads, -1)), sd[a + "norm_added_k.weight"].to(dtype))
    tv = _lin(sd, a + "add_v_proj", cm, dtype).unflatten(-1, (heads, -1))
    iq, ik = apply_rope_qwen(iq, img_freqs), apply_rope_qwen(ik, img_freqs)
    tq, tk = apply_rope_qwen(tq, txt_freqs), apply_rope_qwen(tk, txt_freqs)
    o = _attention(torch.cat([tq, iq], 1), torch.cat([tk, ik], 1), torch.cat([tv, iv], 1)).to(iq.dtype)
    St = c.shape[1]
    x = x + i_g1[:, None] * _lin(sd, a + "to_out.0", o[:, St:], dtype)
    c = c + t_g1[:, None] * _lin(sd, a + "to_add_out", o[:, :St], dtype)
    xm2 = _ln(x) * (1 + i_sc2[:, None]) + i_sh2[:, None]
    x = x + i_g2[:, None] * _lin(sd, p + "img_mlp.net.2",
                                 F.gelu(_lin(sd, p + "img_mlp.net.0.proj", xm2, dtype, ls), approximate="tanh"), dtype, ls)
    cm2 = _ln(c) * (1 + t_sc2[:, None]) + t_sh2[:, None]
    c = c + t_g2[:, None] * _lin(sd, p + "txt_mlp.net.2",
                                 F.gelu(_lin(sd, p + "txt_mlp.net.0.proj", cm2, dtype, ls), approximate="tanh"), dtype, ls)
    return c, x


def qwen_trunk(sd: Dict[str, Tensor], cfg, hidden_states: Tensor, encoder_hidden_states: Tensor,
               timestep: Tensor, grid_hw: Sequence[int], dtype=torch.float32, lora_scale: float = 1.0,
               bf16_quirks: bool = True):
    """Everything of the Qwen-Image transformer up to (not including) norm_out: returns (image hidden states, temb).
    Shared by the ArcFlow student (arcqwen.py:106-157) and the stock teacher (diffusers QwenImageTransformer2DModel
    reached through lakonlab/models/architecture/diffusers/qwen.py:107-139)."""
    heads = cfg.num_attention_heads
    x = _lin(sd, "img_in", hidden_states.to(dtype), dtype)
    t = timestep.to(torch.bfloat16 if bf16_quirks else dtype)
    c = rms_norm(encoder_hidden_states.to(dtype), sd["txt_norm.weight"].to(dtype))
    c = _lin(sd, "txt_in", c, dtype)
    te = "time_text_embed.timestep_embedder"
    tproj = timestep_proj(t, scale=1000.0).to(dtype)
    temb = _lin(sd, te + ".linear_2", F.silu(_lin(sd, te + ".linear_1", tproj, dtype, lora_scale)), dtype, lora_scale)
    rope = _cached_rope(("qwen", c.shape[1], tuple(grid_hw), tuple(cfg.axes_dims_rope), str(x.device)),
                        lambda: tuple(f.to(x.device) for f in QwenEmbedRope(10000, cfg.axes_dims_rope, True)(
                            1, grid_hw[0], grid_hw[1], c.shape[1])))
    for i in range(cfg.num_layers):
        c, x = qwen_block(sd, f"transformer_blocks.{i}.", x, c, temb, rope, heads, dtype, lora_scale)
    return x, temb


def qwen_forward(sd: Dict[str, Tensor], cfg, hidden_states: Tensor, encoder_hidden_states: Tensor,
                 timestep: Tensor, grid_hw: Sequence[int], dtype=torch.float32, lora_scale: float = 1.0,
                 bf16_quirks: bool = True) -> Dict[str, Tensor]:
    """_ArcQwenImageTransformer2DModel.forward (arcqwen.py:106-174). `timestep` is sigma in [0, 1]
    (arcqwen_pipeline.py:412); it is cast to the hidden dtype (:128, bf16 in deployment) and scaled by 1000
    inside Timesteps(scale=1000) in fp32 (QwenTimestepProjEmbeddings)."""
    x, temb = qwen_trunk(sd, cfg, hidden_states, encoder_hidden_states, timestep, grid_hw, dtype, lora_scale, bf16_quirks)
    emb = _lin(sd, "norm_out.linear", F.silu(temb).to(x.dtype), dtype)
    scale, shift = emb.chunk(2, dim=1)
    x = _ln(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    bs, seq, _ = x.shape
    K, C, L = cfg.num_gaussians, cfg.out_channels, cfg.logweights_channels
    return dict(means=_lin(sd, "proj_out_means", x, dtype).reshape(bs, seq, K, C),
                logweights=_lin(sd, "proj_out_logweights", x, dtype).reshape(bs, seq, K, L).log_softmax(dim=-2),
                loggammas=_lin(sd, "proj_out_loggamma", x, dtype).reshape(bs, seq, K - 1, L))


def qwen_denoise(sd, cfg, latents: Tensor, prompt_embeds: Tensor, grid_hw: Sequence[int], num_inference_steps: int = 2,
                 total_substeps: int = 128, timestep_ratio: float = 1.0, shift: float = 3.2, dtype=torch.float32,
                 eps: float = 1e-4, net_dtype=torch.bfloat16):
    """ArcQwenImagePipeline.__call__ denoising loop (arcqwen_pipeline.py:395-463); same structure as FLUX."""
    gh, gw = grid_hw
    raw, substeps, total = retrieve_raw_timesteps(num_inference_steps, total_substeps, timestep_ratio)
    timesteps = scheduler_timesteps(raw, shift)
    B = latents.shape[0]
    tid = 0
    latents = latents.to(torch.float32)
    for i in range(num_inference_steps):
        t_src = timesteps[tid]
        sigma_src = t_src / 1000.0
        out = qwen_forward(sd, cfg, latents.to(net_dtype), prompt_embeds, (t_src.expand(B) / 1000).to(latents.device), grid_hw,
                           dtype=dtype)
        out = {k: v.to(net_dtype).to(torch.float32) for k, v in out.items()}
        mp = unpack_mp(out, gh, gw, cfg.num_gaussians)
        tid += substeps[i]
        t_end = timesteps[tid] if tid < len(timesteps) else torch.tensor(0.0)
        x_img = momentum_integration(mp, unpack_latents(latents, gh, gw), float(sigma_src), float(sigma_src),
                                     float(t_end / 1000.0), eps)
        latents = pack_latents(x_img)
    return latents
