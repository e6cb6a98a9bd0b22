"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/arcflow_oracle.py for the rules).

CPU restatement of the reference's data-free trajectory-distillation TRAIN STEP (forward + loss; gradients come
from torch autograd on this restatement), in the reference's own image-major layout:

  ArcFlowImitationDataFree.forward_initialize / forward_train   lakonlab/models/diffusions/arcflow.py:343-420
  ArcFlowImitationBase.piid_segment_momentum                    lakonlab/models/diffusions/arcflow.py:120-209
  ArcFlowImitationBase.policy_average_u_momentum                lakonlab/models/diffusions/arcflow.py:81-110
  ArcFlowImitationBase.momentum_integration (train variant)     lakonlab/models/diffusions/arcflow.py:28-79
  ArcFlowPolicy.velocity / dropout_                             lakonlab/models/diffusions/policies/arcflow.py:52-106
  ContinuousTimeStepSampler.warp_t                              lakonlab/models/diffusions/sampler.py:46-48
  DiffusionMSELoss (+ mmgen DDPMLoss / mse_loss 'flatmean')     lakonlab/models/losses/diffusion_loss.py:45-83, SURVEY App. A.9
  train_fwd_bwd (sum of step losses)                            lakonlab/models/base_diffusion.py:14-62
  GaussianFlow.pred / forward_u (no CFG for FLUX)               lakonlab/models/diffusions/gaussian_flow.py:90-107, 224-254
  teacher FluxTransformer2DModel.forward (stock FLUX velocity)  lakonlab/models/architecture/diffusers/flux.py:122-156

Every random draw of the reference (initial noise, the mixture-component dropout uniforms, the two interval
uniforms) is an INPUT here, in the order the reference draws them, so CPU and CUDA generators never have to
agree (SURVEY.md §7 "Hard parts"). Pinned against the reference's own functions in tests/test_oracle_golden.py
(roll-out with a synthetic teacher); the mmgen loss reduction (A.9) and the transformer blocks stay unpinned.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import arcflow_oracle as O

Tensor = torch.Tensor


def warp_t(t: Tensor, shift: float = 3.2) -> Tensor:
    return shift * t / (1 + (shift - 1) * t)


def dropout_mask(u: Tensor, p: float) -> Tensor:
    """ArcFlowPolicy.dropout_: mask = rand < p; a sample whose components are ALL dropped keeps all of them."""
    mask = u < p
    is_all = mask.all(dim=1, keepdim=True)
    return mask & ~is_all


def momentum_integration(mp: Dict[str, Tensor], x_t_start: Tensor, sigma_t_src: Tensor, sigma_t_start: Tensor,
                         sigma_t_end: Tensor, eps: float = 1e-4) -> Tensor:
    """Train variant: sigma_* are [B, 1, 1, 1] tensors (arcflow.py:28-79)."""
    means, log_gammas, logweights = mp["means"], mp["loggammas"], mp["logweights"]
    dt_past = (sigma_t_src - sigma_t_start).unsqueeze(1)
    dt_step = (sigma_t_start - sigma_t_end).unsqueeze(1)
    decay = torch.exp(log_gammas * dt_past)
    decay = torch.cat([decay.new_ones((decay.shape[0], 1, *decay.shape[2:])), decay], dim=1)
    z = log_gammas * dt_step
    sign = torch.sign(z)
    sign[sign == 0] = 1
    z = sign * torch.clamp(z.abs(), min=eps)
    step = torch.expm1(z) / z
    step = torch.cat([step.new_ones((step.shape[0], 1, *step.shape[2:])), step], dim=1)
    weights = torch.softmax(logweights, dim=1)
    return x_t_start - (weights * (means * decay * dt_step * step)).sum(dim=1)


def policy_velocity(mp, sigma_t_src: Tensor, sigma_t: Tensor) -> Tensor:
    weights = torch.softmax(mp["logweights"], dim=1)
    decay = torch.exp(mp["loggammas"] * (sigma_t_src - sigma_t).unsqueeze(1))
    decay = torch.cat([decay.new_ones((decay.shape[0], 1, *decay.shape[2:])), decay], dim=1)
    return (mp["means"] * decay * weights).sum(dim=1)


def policy_average_u(mp, x_t_start, sigma_t_src, sigma_t_start, raw_t_start, raw_t_end, total_substeps, shift, eps=1e-4):
    bs = x_t_start.size(0)
    is_small = torch.round((raw_t_start - raw_t_end) * total_substeps) < 2
    pred_mean = pred_local = None
    if not is_small.all():
        sigma_t_end = warp_t(raw_t_end, shift).reshape(bs, 1, 1, 1)
        x_end = momentum_integration(mp, x_t_start, sigma_t_src, sigma_t_start, sigma_t_end, eps)
        pred_mean = (x_t_start - x_end) / (sigma_t_start - sigma_t_end).clamp(min=eps)
    if is_small.any():
        pred_local = policy_velocity(mp, sigma_t_src, sigma_t_start)
    if pred_mean is None:
        return pred_local
    if pred_local is None:
        return pred_mean
    return torch.where(is_small.reshape(bs, 1, 1, 1), pred_local, pred_mean)


def mse_loss_scaled(u_pred: Tensor, u_tgt: Tensor, scale: float = 30.0) -> Tensor:
    """DiffusionMSELoss: flatmean((pred - tgt)^2) * 0.5 -> constant rescale x scale -> mean over samples."""
    per_sample = ((u_pred - u_tgt) ** 2).flatten(1).mean(1) * 0.5
    return (per_sample * scale).mean()


def piid_segment(mp, x_t_src, raw_t_src, sigma_t_src, teacher_ratio, segment_size, teacher_u: Callable,
                 rand: Dict[str, Tensor], cfg: Dict, shift: float = 3.2, loss_scale: float = 30.0):
    """piid_segment_momentum (arcflow.py:120-209). `mp` carries grad; the roll-out uses its detached, dropped copy.
    rand: drop_u [B, K], student_u [B, n], teacher_u [B, n-1] uniforms in the reference's draw order."""
    eps = cfg.get("eps", 1e-4)
    total_substeps = cfg.get("total_substeps", 128)
    n_states = cfg.get("num_intermediate_states", 2)
    window_substeps = cfg.get("window_substeps", 0)
    bs = x_t_src.size(0)
    segment_size = torch.tensor([segment_size], dtype=torch.float32)
    num_substeps = (segment_size * total_substeps).round().to(torch.long).clamp(min=1)
    substep_size = segment_size / num_substeps
    window_size = torch.minimum(window_substeps * substep_size, segment_size)
    raw_t_dst = raw_t_src - segment_size

    det = {k: v.detach() for k, v in mp.items()}
    p = cfg.get("gm_dropout", 0.0)
    if 0 < p < 1:
        mask = dropout_mask(rand["drop_u"], p).reshape(bs, -1, 1, 1, 1)
        det["logweights"] = det["logweights"].masked_fill(mask, float("-inf"))

    student_iv = rand["student_u"] * ((1 - teacher_ratio) * (segment_size - window_size).unsqueeze(-1))
    student_iv = torch.sort(student_iv, dim=-1)[0]
    student_iv = torch.diff(student_iv, dim=-1, prepend=torch.zeros((bs, 1)))
    teacher_iv = torch.sort(rand["teacher_u"], dim=-1)[0]
    teacher_iv = torch.diff(teacher_iv, dim=-1, prepend=torch.zeros((bs, 1)), append=torch.ones((bs, 1))) * (
        teacher_ratio * (segment_size - window_size).unsqueeze(-1))

    x_t, raw_t, sigma_t = x_t_src, raw_t_src, sigma_t_src
    all_pred, all_tgt = [], []
    for k in range(n_states):
        raw_t_a = (raw_t - student_iv[:, k]).clamp(min=0)
        raw_t_b = (raw_t_a - teacher_iv[:, k]).clamp(min=0)
        with torch.no_grad():
            sigma_t_a = warp_t(raw_t_a, shift).reshape(bs, 1, 1, 1)
            x_t_a = momentum_integration(det, x_t, sigma_t_src, sigma_t, sigma_t_a, eps)
            tgt_u = teacher_u(x_t_a, sigma_t_a.flatten())
        all_tgt.append(tgt_u)
        all_pred.append(policy_average_u(mp, x_t_a, sigma_t_src, sigma_t_a, raw_t_a, raw_t_b - window_size,
                                         total_substeps, shift, eps))
        sigma_t_b = warp_t(raw_t_b, shift).reshape(bs, 1, 1, 1)
        x_t = x_t_a + tgt_u * (sigma_t_b - sigma_t_a)
        raw_t, sigma_t = raw_t_b, sigma_t_b
    loss = mse_loss_scaled(torch.cat(all_pred, 0), torch.cat(all_tgt, 0), loss_scale)
    with torch.no_grad():
        x_t_dst = momentum_integration(det, x_t, sigma_t_src, sigma_t, warp_t(raw_t_dst, shift).reshape(bs, 1, 1, 1), eps)
    return loss, x_t_dst, raw_t_dst, dict(pred=all_pred, tgt=all_tgt)


class _lora_dropout:
    """Train-mode peft dropout on the student forward only (the tied teacher has no LoRA branches): enables
    O.LORA_DROPOUT with this step's mask seed for the duration of the call."""

    def __init__(self, train_cfg, rand, num_double):
        p = float(train_cfg.get("lora_dropout", 0.0) or 0.0)
        self.cfg = dict(p=p, seed=int(rand["lora_seed"]), num_double=num_double) if p > 0 else None

    def __enter__(self):
        O.LORA_DROPOUT = self.cfg

    def __exit__(self, *exc):
        O.LORA_DROPOUT = None


def teacher_state_dict(sd: Dict[str, Tensor], teacher_extra: Dict[str, Tensor]) -> Dict[str, Tensor]:
    """The teacher is the stock trunk (weights tied to the student's frozen base layers, base_diffusion.py:93-94)
    with its own `norm_out` and `proj_out`."""
    t = {k: v for k, v in sd.items() if "lora" not in k and not k.startswith(("proj_out_", "norm_out."))}
    t.update(teacher_extra)
    return t


def flux_teacher_velocity(tsd, cfg, x_tokens: Tensor, txt, pooled, sigma: Tensor, guidance: Tensor, grid_hw,
                          dtype=torch.float32, bf16_quirks: bool = True) -> Tensor:
    """Stock FLUX forward: same trunk, AdaLayerNormContinuous + proj_out (D -> 64) (diffusers/flux.py:122-156).
    Cross-checked against the original black-forest-labs FLUX model code (tests/test_oracle_bfl.py)."""
    import torch.nn.functional as F
    out = O.flux_trunk(tsd, cfg, x_tokens, txt, pooled, sigma, guidance, grid_hw, dtype=dtype, bf16_quirks=bf16_quirks)
    x, temb = out
    emb = O._lin(tsd, "norm_out.linear", F.silu(temb).to(x.dtype), dtype)
    scale, shift = emb.chunk(2, dim=1)
    x = O._ln(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    return O._lin(tsd, "proj_out", x, dtype)


def flux_train_forward(sd, teacher_extra, cfg, txt, pooled, grid_hw, noise_tokens: Tensor, rands: Sequence[Dict],
                       iteration: int, train_cfg: Dict, shift: float = 3.2, loss_scale: float = 30.0,
                       dtype=torch.float32, net_dtype=torch.bfloat16, guidance_scale: float = 3.5,
                       require_grad: Sequence[str] = ()):
    """One data-free train iteration (forward + summed loss) in image-major layout. Returns (loss, log_vars, extras).
    `require_grad`: names of sd entries to differentiate (fp32 leaf copies are made and returned in extras)."""
    gh, gw = grid_hw
    B = noise_tokens.shape[0]
    leaves = {}
    if require_grad:
        sd = dict(sd)
        for k in require_grad:
            leaves[k] = sd[k].detach().to(dtype).clone().requires_grad_(True)
            sd[k] = leaves[k]
    tsd = teacher_state_dict({k: (v.detach() if isinstance(v, Tensor) else v) for k, v in sd.items()}, teacher_extra)
    num_decay = train_cfg.get("num_decay_iters", 0)
    teacher_ratio = 1 - min(iteration, num_decay) / num_decay if num_decay > 0 else 0.0
    nfe = train_cfg["nfe"]
    eps = train_cfg.get("eps", 1e-4)
    ratio = max(train_cfg.get("timestep_ratio", 1.0), eps)
    base_seg = 1 / (nfe - 1 + ratio)
    guidance = torch.full([B], guidance_scale, dtype=torch.float32) if cfg.guidance_embeds else None

    def teacher_u(x_img, sigma):
        with torch.no_grad():
            v = flux_teacher_velocity(tsd, cfg, O.pack_latents(x_img).to(net_dtype), txt, pooled, sigma, guidance,
                                      grid_hw, dtype=dtype)
        return O.unpack_latents(v.to(net_dtype).to(torch.float32), gh, gw)

    x_t_src = O.unpack_latents(noise_tokens.to(torch.float32), gh, gw)
    raw_t_src = torch.ones(B, dtype=torch.float32)
    loss = 0
    log_vars = dict(teacher_ratio=teacher_ratio) if num_decay > 0 else {}
    trace = []
    for step_id in range(nfe):
        seg = base_seg * ratio if step_id == nfe - 1 else base_seg
        sigma_t_src = warp_t(raw_t_src, shift).reshape(B, 1, 1, 1)
        with _lora_dropout(train_cfg, rands[step_id], cfg.num_layers):
            out = O.flux_forward(sd, cfg, O.pack_latents(x_t_src).to(net_dtype), txt, pooled, sigma_t_src.flatten(), guidance,
                                 grid_hw, dtype=dtype)
        # network emits net_dtype (bf16); GaussianFlow.pred casts back to fp32. Straight-through for the grad path.
        out = {k: v + (v.to(net_dtype).to(v.dtype) - v).detach() for k, v in out.items()}
        mp = O.unpack_mp({k: v.to(torch.float32) for k, v in out.items()}, gh, gw, cfg.num_gaussians)
        step_loss, x_t_dst, raw_t_dst, tr = piid_segment(mp, x_t_src, raw_t_src, sigma_t_src, teacher_ratio, seg, teacher_u,
                                                         rands[step_id], train_cfg, shift, loss_scale)
        loss = loss + step_loss * seg
        log_vars[f"loss_diffusion_step{step_id}"] = float(step_loss.detach())
        log_vars["loss_diffusion"] = log_vars.get("loss_diffusion", 0.0) + float((step_loss * seg).detach())
        trace.append(dict(x_t_dst=x_t_dst, **tr))
        x_t_src, raw_t_src = x_t_dst, raw_t_dst
    return loss, log_vars, dict(leaves=leaves, trace=trace)


def qwen_teacher_velocity(tsd, cfg, x_tokens: Tensor, txt, sigma: Tensor, grid_hw, dtype=torch.float32) -> Tensor:
    """Stock Qwen-Image forward: same trunk, AdaLayerNormContinuous + proj_out (D -> 64)
    (lakonlab/models/architecture/diffusers/qwen.py:107-139)."""
    import torch.nn.functional as F
    x, temb = O.qwen_trunk(tsd, cfg, x_tokens, txt, sigma, grid_hw, dtype=dtype)
    emb = O._lin(tsd, "norm_out.linear", F.silu(temb).to(x.dtype), dtype)
    scale, shift = emb.chunk(2, dim=1)
    x = O._ln(x) * (1 + scale)[:, None, :] + shift[:, None, :]
    return O._lin(tsd, "proj_out", x, dtype)


def qwen_teacher_cfg_velocity(tsd, cfg, x_tokens: Tensor, txt_pos, txt_neg, sigma: Tensor, guidance_scale: float, grid_hw,
                              dtype=torch.float32, net_dtype=torch.bfloat16) -> Tensor:
    """GaussianFlow.forward_u with true classifier-free guidance (gaussian_flow.py:224-254): one batch-doubled call on
    [neg; pos] text (latent_diffusion_text_image.py:75-78), `mean_pos + (mean_pos - mean_neg) * (g - 1)` in fp32 on the
    network's bf16 outputs (pred() casts them back to the fp32 x_t dtype, :90-107)."""
    if not guidance_scale > 1.0:
        return qwen_teacher_velocity(tsd, cfg, x_tokens, txt_pos, sigma, grid_hw, dtype=dtype).to(net_dtype).to(torch.float32)
    both = qwen_teacher_velocity(tsd, cfg, torch.cat([x_tokens, x_tokens], 0), torch.cat([txt_neg, txt_pos], 0),
                                 torch.cat([sigma, sigma], 0), grid_hw, dtype=dtype).to(net_dtype).to(torch.float32)
    neg, pos = both.chunk(2, dim=0)
    return pos + (pos - neg) * (guidance_scale - 1)


def qwen_train_forward(sd, teacher_extra, cfg, txt, txt_neg, grid_hw, noise_tokens: Tensor, rands: Sequence[Dict],
                       iteration: int, train_cfg: Dict, shift: float = 3.2, loss_scale: float = 30.0,
                       dtype=torch.float32, net_dtype=torch.bfloat16, require_grad: Sequence[str] = ()):
    """flux_train_forward for the Qwen-Image student / teacher pair (configs/qwen/arcqwen_2nfe_k16.py: same roll-out,
    teacher_guidance_scale = 4.0 true CFG, no distilled-guidance embedding, no pooled text)."""
    gh, gw = grid_hw
    B = noise_tokens.shape[0]
    leaves = {}
    if require_grad:
        sd = dict(sd)
        for k in require_grad:
            leaves[k] = sd[k].detach().to(dtype).clone().requires_grad_(True)
            sd[k] = leaves[k]
    tsd = teacher_state_dict({k: (v.detach() if isinstance(v, Tensor) else v) for k, v in sd.items()}, teacher_extra)
    num_decay = train_cfg.get("num_decay_iters", 0)
    teacher_ratio = 1 - min(iteration, num_decay) / num_decay if num_decay > 0 else 0.0
    nfe = train_cfg["nfe"]
    eps = train_cfg.get("eps", 1e-4)
    ratio = max(train_cfg.get("timestep_ratio", 1.0), eps)
    base_seg = 1 / (nfe - 1 + ratio)
    g_teacher = train_cfg.get("teacher_guidance_scale", 4.0)

    def teacher_u(x_img, sigma):
        with torch.no_grad():
            v = qwen_teacher_cfg_velocity(tsd, cfg, O.pack_latents(x_img).to(net_dtype), txt, txt_neg, sigma, g_teacher,
                                          grid_hw, dtype=dtype, net_dtype=net_dtype)
        return O.unpack_latents(v, gh, gw)

    x_t_src = O.unpack_latents(noise_tokens.to(torch.float32), gh, gw)
    raw_t_src = torch.ones(B, dtype=torch.float32)
    loss = 0
    log_vars = dict(teacher_ratio=teacher_ratio) if num_decay > 0 else {}
    trace = []
    for step_id in range(nfe):
        seg = base_seg * ratio if step_id == nfe - 1 else base_seg
        sigma_t_src = warp_t(raw_t_src, shift).reshape(B, 1, 1, 1)
        with _lora_dropout(train_cfg, rands[step_id], cfg.num_layers):
            out = O.qwen_forward(sd, cfg, O.pack_latents(x_t_src).to(net_dtype), txt, sigma_t_src.flatten(), grid_hw,
                                 dtype=dtype)
        out = {k: v + (v.to(net_dtype).to(v.dtype) - v).detach() for k, v in out.items()}
        mp = O.unpack_mp({k: v.to(torch.float32) for k, v in out.items()}, gh, gw, cfg.num_gaussians)
        step_loss, x_t_dst, raw_t_dst, tr = piid_segment(mp, x_t_src, raw_t_src, sigma_t_src, teacher_ratio, seg, teacher_u,
                                                         rands[step_id], train_cfg, shift, loss_scale)
        loss = loss + step_loss * seg
        log_vars[f"loss_diffusion_step{step_id}"] = float(step_loss.detach())
        log_vars["loss_diffusion"] = log_vars.get("loss_diffusion", 0.0) + float((step_loss * seg).detach())
        trace.append(dict(x_t_dst=x_t_dst, **tr))
        x_t_src, raw_t_src = x_t_dst, raw_t_dst
    return loss, log_vars, dict(leaves=leaves, trace=trace)
