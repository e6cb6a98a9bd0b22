"""Data-free trajectory-distillation train step on the native kernels — FORWARD + LOSS (round 1).

Restates, in packed-token layout and with every random draw injected,
  ArcFlowImitationDataFree.forward_initialize / forward_train  lakonlab/models/diffusions/arcflow.py:343-420
  ArcFlowImitationBase.piid_segment_momentum                   lakonlab/models/diffusions/arcflow.py:120-209
  train_fwd_bwd (sum of per-step losses)                       lakonlab/models/base_diffusion.py:14-62
  DiffusionMSELoss, constant rescale 30                        lakonlab/models/losses/diffusion_loss.py:45-83
Per student step: 1 student forward (afb_engine_forward), 4 x {afb_policy_eval INTEGRATE with the detached, dropped
policy -> teacher velocity (afb_engine_forward on the tied teacher) -> afb_policy_eval AVERAGE_U -> afb_mse_rows ->
afb_axpy_rows (teacher Euler step)}, then one INTEGRATE to the segment end. Host code only does the O(batch)
schedule arithmetic the reference also does on tiny tensors.

Backward, round-1 state: `backward_heads()` returns the EXACT gradients of the adapter tensors that sit after the
trunk — proj_out_means / proj_out_logweights / proj_out_loggamma (weight, bias) and norm_out.linear (weight, bias) —
through afb_policy_backward (loss + average-velocity + softmax/log-softmax + integral term), afb_gemm_tn (dW = dY^T X on
tcgen05), afb_gemm (dX through the head weights), afb_ln_mod_param_grad and afb_rowlinear_param_grad.
NOT built yet (next round): the LoRA gradients, which need the backward through the frozen trunk
(attention backward, dX GEMMs with transposed weights, LN/GELU/RMSNorm/RoPE backward, per-block recompute), LoRA dropout
(p = 0.05; this path is the p = 0 forward), optimizer / EMA / DDP all-reduce. `backward()` raises for those.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import _lib, ops
from ._lib import AfbError

DEFAULT_TRAIN_CFG = dict(  # configs/flux/arcflux_2nfe_k16.py:89-99
    num_decay_iters=2000, window_substeps=3, gm_dropout=0.1, num_intermediate_states=4,
    distilled_guidance_scale=3.5, teacher_distilled_guidance_scale=3.5, nfe=2, timestep_ratio=1.0,
    total_substeps=128, eps=1e-4, lora_dropout=0.05)   # lora_dropout: configs/flux/arcflux_2nfe_k16.py:40-48
QWEN_TRAIN_CFG = dict(  # configs/qwen/arcqwen_2nfe_k16.py:96-106 — true CFG on the teacher, no guidance embedding
    num_decay_iters=2000, window_substeps=3, gm_dropout=0.1, num_intermediate_states=4, teacher_guidance_scale=4.0,
    nfe=2, timestep_ratio=1.0, total_substeps=128, eps=1e-4, lora_dropout=0.05)


def warp_t(t: torch.Tensor, shift: float) -> torch.Tensor:
    """ContinuousTimeStepSampler.warp_t with a fixed shift (lakonlab/models/diffusions/sampler.py:46-48)."""
    return shift * t / (1 + (shift - 1) * t)


def draw_rollout_randoms(batch: int, num_states: int, num_gaussians: int, generator: Optional[torch.Generator] = None):
    """The three uniform draws of one student step, in the reference's order (policies/arcflow.py:100-101,
    arcflow.py:148-155)."""
    r = lambda *s: torch.rand(s, generator=generator)
    out = dict(drop_u=r(batch, num_gaussians), student_u=r(batch, num_states), teacher_u=r(batch, num_states - 1))
    # seed of the student forward's LoRA-dropout mask stream (drawn last so the three reference draws keep their order)
    out["lora_seed"] = int(torch.randint(0, 2 ** 62, (1,), generator=generator).item())
    return out


class ArcFlowDistillStep:
    def __init__(self, student, teacher, train_cfg: Optional[Dict] = None, shift: float = 3.2, loss_scale: float = 30.0):
        self.student, self.teacher = student, teacher
        self.qwen = getattr(student, "arch", "flux") == "qwen"
        self.cfg = dict(QWEN_TRAIN_CFG if self.qwen else DEFAULT_TRAIN_CFG)
        if train_cfg:
            self.cfg.update(train_cfg)
        self.shift, self.loss_scale = shift, loss_scale

    def teacher_ratio(self, iteration: int) -> float:
        n = self.cfg.get("num_decay_iters", 0)
        return 1 - min(iteration, n) / n if n > 0 else 0.0

    @torch.no_grad()
    def forward(self, txt: torch.Tensor, pooled: torch.Tensor, grid_hw: Sequence[int], noise: torch.Tensor,
                rands: Sequence[Dict[str, torch.Tensor]], iteration: int = 0, save_for_backward: bool = False,
                step_hook=None, neg_txt: Optional[torch.Tensor] = None):
        """One train iteration, forward only. noise: fp32 packed tokens [B, S_i, 64] (the data-free x_t_src);
        rands: one dict of uniforms per student step (see draw_rollout_randoms). Returns (loss, log_vars, extras).
        FLUX: pooled = pooled text projections. Qwen: pooled is None and neg_txt carries the negative-prompt embeds of
        the teacher's true CFG (latent_diffusion_text_image.py:63-78).
        step_hook(saved): called after each student step's roll-out while that step's trunk checkpoints are still live
        (forward_backward() uses it to run the step's backward before the next student forward overwrites them)."""
        cfg, st, te = self.cfg, self.student, self.teacher
        if noise.dtype != torch.float32 or not noise.is_cuda:
            raise AfbError("noise must be an fp32 CUDA tensor [batch, tokens, 64]")
        B = noise.shape[0]
        nfe, eps = cfg["nfe"], cfg.get("eps", 1e-4)
        if len(rands) != nfe:
            raise AfbError(f"need {nfe} sets of roll-out uniforms, got {len(rands)}")
        ratio_t = max(cfg.get("timestep_ratio", 1.0), eps)
        base_seg = 1.0 / (nfe - 1 + ratio_t)
        n_states, total_sub = cfg["num_intermediate_states"], cfg["total_substeps"]
        K = st.num_gaussians
        if self.qwen:
            g_true = cfg.get("teacher_guidance_scale") or 1.0
            student_heads = lambda x_, sig_, train_: st.forward_heads(x_, txt, sig_, grid_hw, train=train_)
            teacher_u = lambda x_, sig_: te.velocity(x_, txt, neg_txt, sig_, g_true, grid_hw)
        else:
            g_student, g_teacher = cfg["distilled_guidance_scale"], cfg["teacher_distilled_guidance_scale"]
            student_heads = lambda x_, sig_, train_: st.forward_heads(x_, txt, pooled, sig_, g_student, grid_hw, train=train_)
            teacher_u = lambda x_, sig_: te.velocity(x_, txt, pooled, sig_, g_teacher, grid_hw)
        teacher_ratio = self.teacher_ratio(iteration)
        log_vars = dict(teacher_ratio=teacher_ratio) if cfg.get("num_decay_iters", 0) > 0 else {}

        x_src = noise.contiguous()
        raw_t_src = torch.ones(B, dtype=torch.float32)
        loss_total = 0.0
        extras = dict(steps=[])
        for step_id in range(nfe):
            seg_f = base_seg * ratio_t if step_id == nfe - 1 else base_seg
            seg = torch.tensor([seg_f], dtype=torch.float32)
            num_sub = (seg * total_sub).round().to(torch.long).clamp(min=1)
            window = torch.minimum(cfg["window_substeps"] * (seg / num_sub), seg)
            raw_t_dst = raw_t_src - seg
            sigma_src = warp_t(raw_t_src, self.shift)

            p_lora = float(cfg.get("lora_dropout", 0.0) or 0.0)
            if p_lora > 0 and "lora_seed" not in rands[step_id]:
                raise AfbError("lora_dropout > 0 needs rands[step]['lora_seed'] (see draw_rollout_randoms)")
            st.set_lora_dropout(p_lora, rands[step_id].get("lora_seed", 0))
            # the dropped forward IS the train forward (peft drops only in train mode), with or without a backward hook
            head = student_heads(x_src, sigma_src, step_hook is not None or p_lora > 0)
            head2 = head.reshape(-1, head.shape[-1])
            saved = None
            if save_for_backward:
                St, Si = txt.shape[1], x_src.shape[1]
                saved = dict(head=head2, seg=seg_f, sigma_src=sigma_src.clone(), states=[],
                             hidden=st.export_activation("hidden", B, St, Si),
                             head_in=st.export_activation("head_in", B, St, Si),
                             temb=st.export_activation("temb", B, St, Si))

            rd = rands[step_id]
            p = cfg.get("gm_dropout", 0.0)
            drop = None
            if 0 < p < 1:
                m = rd["drop_u"].cpu() < p
                drop = m & ~m.all(dim=1, keepdim=True)
            s_iv = rd["student_u"].cpu() * ((1 - teacher_ratio) * (seg - window).unsqueeze(-1))
            s_iv = torch.diff(torch.sort(s_iv, dim=-1)[0], dim=-1, prepend=torch.zeros((B, 1)))
            t_iv = torch.diff(torch.sort(rd["teacher_u"].cpu(), dim=-1)[0], dim=-1, prepend=torch.zeros((B, 1)),
                              append=torch.ones((B, 1))) * (teacher_ratio * (seg - window).unsqueeze(-1))

            x_t, raw_t, sigma_t = x_src, raw_t_src, sigma_src
            mse_sum = torch.zeros(B, dtype=torch.float32, device=noise.device)
            for k in range(n_states):
                raw_t_a = (raw_t - s_iv[:, k]).clamp(min=0)
                raw_t_b = (raw_t_a - t_iv[:, k]).clamp(min=0)
                sigma_a, sigma_b = warp_t(raw_t_a, self.shift), warp_t(raw_t_b, self.shift)
                x_a, x_a_bf = ops.policy_eval(head2, _lib.AFB_POLICY_INTEGRATE, sigma_src, sigma_t, sigma_a, x=x_t,
                                              batch=B, drop_mask=drop, num_gaussians=K, eps=eps, want_bf16=True)
                tgt_u = teacher_u(x_a_bf, sigma_a)
                raw_end = raw_t_b - window
                small = torch.round((raw_t_a - raw_end) * total_sub) < 2
                pred_u = ops.policy_eval(head2, _lib.AFB_POLICY_AVERAGE_U, sigma_src, sigma_a, warp_t(raw_end, self.shift),
                                         batch=B, small=small, num_gaussians=K, eps=eps)
                mse_sum += ops.mse_rows(pred_u, tgt_u)
                if saved is not None:
                    saved["states"].append(dict(tgt_u=tgt_u, sigma_a=sigma_a.clone(), sigma_end=warp_t(raw_end, self.shift),
                                                small=small.clone()))
                x_t = ops.axpy_rows(x_a, tgt_u, sigma_b - sigma_a)
                raw_t, sigma_t = raw_t_b, sigma_b
            # DiffusionMSELoss: 0.5 * flatmean -> x loss_scale -> mean over the 4B stacked samples
            step_loss = float(mse_sum.sum().item()) * 0.5 * self.loss_scale / (n_states * B)
            x_dst = ops.policy_eval(head2, _lib.AFB_POLICY_INTEGRATE, sigma_src, sigma_t, warp_t(raw_t_dst, self.shift),
                                    x=x_t, batch=B, drop_mask=drop, num_gaussians=K, eps=eps)
            loss_total += step_loss * seg_f
            log_vars[f"loss_diffusion_step{step_id}"] = step_loss
            log_vars["loss_diffusion"] = log_vars.get("loss_diffusion", 0.0) + step_loss * seg_f
            if step_hook is not None:
                step_hook(saved)
            extras["steps"].append(dict(x_t_dst=x_dst, head=head, saved=saved))
            x_src, raw_t_src = x_dst, raw_t_dst
        return loss_total, log_vars, extras

    def _head_grad_buffers(self):
        st = self.student
        D, head_n, dev = st.cfg.inner_dim, st.weights.head_n, st.device
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        return dict(g_w=z(head_n, D), g_b=z(head_n), g_nw=z(2 * D, D), g_nb=z(2 * D))

    def _split_head_grads(self, acc) -> Dict[str, torch.Tensor]:
        nm, nw, ng = self.student.cfg.head_dims
        g_w, g_b = acc["g_w"], acc["g_b"]
        return {
            "proj_out_means.weight": g_w[:nm], "proj_out_means.bias": g_b[:nm],
            "proj_out_logweights.weight": g_w[nm:nm + nw], "proj_out_logweights.bias": g_b[nm:nm + nw],
            "proj_out_loggamma.weight": g_w[nm + nw:nm + nw + ng], "proj_out_loggamma.bias": g_b[nm + nw:nm + nw + ng],
            "norm_out.linear.weight": acc["g_nw"], "norm_out.linear.bias": acc["g_nb"],
        }

    @torch.no_grad()
    def _backward_step_heads(self, sv, acc, d_mod: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One student step: loss -> d(raw heads) -> head / norm_out gradients (accumulated into acc). Returns dy, the
        gradient w.r.t. the norm_out output (bf16 [B, tokens, D]) that the trunk backward starts from. d_mod (fp32
        [B, mod_total]) receives norm_out's (scale, shift) gradient in its slot."""
        cfg, st = self.cfg, self.student
        n_states, eps = cfg["num_intermediate_states"], cfg.get("eps", 1e-4)
        D, head_n, dev = st.cfg.inner_dim, st.weights.head_n, st.device
        head2 = sv["head"]
        B = sv["temb"].shape[0]
        tokens = head2.shape[0] // B
        # d(loss)/d pred = seg * loss_scale * 0.5 * 2 (pred - tgt) / (elements per sample * n_states * B)
        coef = sv["seg"] * self.loss_scale / (tokens * 64 * n_states * B)
        dhead = None
        for s in sv["states"]:
            dhead = ops.policy_backward(head2, s["tgt_u"], sv["sigma_src"], s["sigma_a"], s["sigma_end"], coef,
                                        dhead=dhead, small=s["small"], num_gaussians=st.num_gaussians, eps=eps)
        return self.backward_from_dhead(sv, dhead, acc, d_mod)

    @torch.no_grad()
    def backward_from_dhead(self, sv, dhead: torch.Tensor, acc, d_mod: Optional[torch.Tensor] = None) -> torch.Tensor:
        """d(raw heads) (fp32 [B * tokens, head_n]) -> head / norm_out gradients (accumulated into acc) and dy, the gradient
        w.r.t. the norm_out output. sv needs 'hidden', 'head_in', 'temb' of the forward (export_activation)."""
        st = self.student
        D, head_n, dev = st.cfg.inner_dim, st.weights.head_n, st.device
        B = sv["temb"].shape[0]
        tokens = dhead.shape[0] // B
        ops.colsum_f32(dhead, acc["g_b"])
        dhead_bf = dhead.to(torch.bfloat16)      # plumbing cast; the GEMM operands are bf16
        x_in = sv["head_in"].reshape(-1, D)
        ops.gemm_tn(dhead_bf, x_in, acc["g_w"])                                # dW_heads += dHead^T y
        dy = torch.empty(B, tokens, D, dtype=torch.bfloat16, device=dev)
        # dy = dHead W_heads: the fused [head_n, D] head weight read transposed in place
        ops.gemm(dhead_bf.reshape(B, tokens, head_n), st.head_weight(), dy, transposed=True)
        dscale, dshift = ops.ln_mod_param_grad(sv["hidden"], dy)
        demb = torch.cat([dscale, dshift], dim=1).contiguous()             # AdaLayerNormContinuous: (scale, shift)
        ops.rowlinear_param_grad(demb, sv["temb"], acc["g_nw"], acc["g_nb"], silu_in=True)
        if d_mod is not None:
            off = st.norm_out_mod_off
            d_mod[:, off:off + 2 * D] += demb
        return dy

    @torch.no_grad()
    def backward_heads(self, extras, head_weight_t: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Exact fp32 gradients of the summed train loss w.r.t. the post-trunk adapter tensors (proj_out_* and
        norm_out.linear), from a forward run with save_for_backward=True. (head_weight_t is no longer needed: the dX
        GEMM reads the head weight transposed in place.)"""
        acc = self._head_grad_buffers()
        for step in extras["steps"]:
            if step["saved"] is None:
                raise AfbError("backward_heads: run forward(save_for_backward=True) first")
            self._backward_step_heads(step["saved"], acc)
        return self._split_head_grads(acc)

    @torch.no_grad()
    def forward_backward(self, txt: torch.Tensor, pooled: torch.Tensor, grid_hw: Sequence[int], noise: torch.Tensor,
                         rands: Sequence[Dict[str, torch.Tensor]], iteration: int = 0,
                         grads: Optional[Dict[str, torch.Tensor]] = None, neg_txt: Optional[torch.Tensor] = None):
        """train_fwd_bwd (lakonlab/models/base_diffusion.py:14-62): forward + loss + the gradients of every adapter
        tensor: heads, norm_out.linear, the trunk's LoRA pairs, and — through the gradients of every AdaLN shift / scale /
        gate vector — the timestep embedder's LoRA pairs.
        grads: fp32 CUDA tensors keyed by state-dict name, accumulated into (+=); created zero-filled when None.
        Each student step's backward runs right after its roll-out: heads / norm_out (backward_heads path), then the
        frozen trunk in reverse with per-block recompute (afb_engine_backward), then afb_engine_backward_embed."""
        st = self.student
        if grads is None:
            grads = {}
        shapes = dict(st.trunk_lora_shapes())
        shapes.update(st.embed_lora_shapes())
        for n, shp in shapes.items():
            if n not in grads:
                grads[n] = torch.zeros(shp, dtype=torch.float32, device=st.device)
        acc = self._head_grad_buffers()

        def hook(saved):
            d_mod = torch.zeros(noise.shape[0], st.mod_total, dtype=torch.float32, device=st.device)
            dy = self._backward_step_heads(saved, acc, d_mod)
            st.backward_trunk(dy, grads, d_mod)
            st.backward_embed(d_mod, grads)

        loss, log_vars, extras = self.forward(txt, pooled, grid_hw, noise, rands, iteration, save_for_backward=True,
                                              step_hook=hook, neg_txt=neg_txt)
        for n, g in self._split_head_grads(acc).items():
            if n in grads:
                grads[n] += g
            else:
                grads[n] = g.clone()
        return loss, log_vars, grads

    def backward(self, *a, **k):
        raise NotImplementedError("use forward_backward(): each student step's backward must run before the next student "
                                  "forward overwrites the trunk checkpoints")


class ArcFlowTrainer:
    """One optimisation iteration of the data-free distillation: forward_backward -> (DDP all-reduce) -> grad clip ->
    AdamW -> Karras EMA -> write the updated bf16 adapter back into the engine's packed weights.
    Reference: the mmcv iter-based runner driving BaseModel.train_step (lakonlab/models/base.py:76-103,
    lakonlab/models/base_diffusion.py:14-62) with configs/flux/_ddp_train.py's optimizer and the EMA hook
    (lakonlab/runner/hooks/ema_hook.py:86-121). Every adapter tensor (LoRA pairs, proj_out_*, norm_out.linear) lives in ONE
    flat fp32 arena (arcflow_b200/optim.py), its gradient view is what the native backward accumulates into."""

    def __init__(self, student, teacher, train_cfg: Optional[Dict] = None, shift: float = 3.2, loss_scale: float = 30.0,
                 **optim_kwargs):
        from .optim import FlatAdamW
        self.student = student
        self.distill = ArcFlowDistillStep(student, teacher, train_cfg, shift, loss_scale)
        views = student.weights.adapter_views
        if not views:
            raise AfbError("ArcFlowTrainer: the student carries no adapter tensors")
        self.opt = FlatAdamW({n: tuple(v.shape) for n, v in views.items()}, student.device, **optim_kwargs)
        self.opt.load_params(views)
        self.grads = {n: self.opt.grad(n) for n in views}
        self.iteration = 0

    @torch.no_grad()
    def write_back(self, use_ema: bool = False):
        """bf16 shadow of the arena (or of the EMA weights, for evaluation / export) -> the engine's packed buffers."""
        src = self.opt.ema.to(torch.bfloat16) if use_ema else self.opt.shadow
        for n, dst in self.student.weights.adapter_views.items():
            dst.copy_(self.opt.view(src, n))

    def adapter_state_dict(self, use_ema: bool = True) -> Dict[str, torch.Tensor]:
        """The adapter tensors under the reference's on-disk names (export_arcflow_to_diffusers.py:104-127)."""
        buf = self.opt.ema if use_ema else self.opt.params
        return {n: self.opt.view(buf, n).to(torch.bfloat16).clone() for n in self.student.weights.adapter_views}

    @torch.no_grad()
    def train_step(self, txt, pooled, grid_hw, noise, rands, iteration: Optional[int] = None, neg_txt=None):
        it = self.iteration if iteration is None else iteration
        self.opt.grads.zero_()
        loss, log_vars, _ = self.distill.forward_backward(txt, pooled, grid_hw, noise, rands, it, grads=self.grads,
                                                          neg_txt=neg_txt)
        log_vars.update(self.opt.step(it))
        self.write_back()
        self.iteration = it + 1
        return loss, log_vars
