"""`ArcFlowImitationDataFree` — the reference's distillation model surface (lakonlab/models/diffusions/arcflow.py:339-426)
on the native train step (arcflow_b200.train.ArcFlowDistillStep). The `forward_initialize` / `train_forward` step-state
protocol of train_fwd_bwd (lakonlab/models/base_diffusion.py:14-62) is kept for callers that want the loss only;
`forward_backward` is the whole multi-step forward + adapter-only backward (each student step's backward runs right after
its roll-out, so the trunk checkpoints are never stale). The runner-facing object is
lakonlab.models.LatentDiffusionTextImage."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from arcflow_b200.train import ArcFlowDistillStep, draw_rollout_randoms


class ArcFlowImitationDataFree:
    is_multistep = True

    def __init__(self, denoising, teacher, train_cfg: Optional[Dict] = None, shift: float = 3.2, loss_scale: float = 30.0,
                 policy_type: str = "ArcFlow"):
        assert policy_type == "ArcFlow", f"Invalid policy: {policy_type}. Supported policies are ['ArcFlow']."
        self.denoising, self.teacher = denoising, teacher
        self.step = ArcFlowDistillStep(denoising, teacher, train_cfg, shift, loss_scale)
        self.train_cfg = self.step.cfg

    def forward_initialize(self, x_0: torch.Tensor, running_status=None, generator=None, **kwargs):
        """x_0 is a dummy (data-free): x_t_src = randn_like(x_0), raw_t_src = 1 (reference :343-367)."""
        it = (running_status or {}).get("iteration", 0)
        ratio = self.step.teacher_ratio(it)
        log_vars = dict(teacher_ratio=ratio) if self.train_cfg.get("num_decay_iters", 0) > 0 else {}
        noise = torch.randn(x_0.shape, generator=generator, device=x_0.device, dtype=torch.float32)
        return dict(step_id=0, terminate=False, detachable=True, teacher_ratio=ratio, x_t_src=noise, iteration=it), log_vars

    def train_forward(self, prompt_embeds, pooled_prompt_embeds, grid_hw, noise, running_status=None, rands=None,
                      generator=None):
        """Whole multi-step forward (what train_fwd_bwd accumulates before its single backward)."""
        it = (running_status or {}).get("iteration", 0)
        B, n, K = noise.shape[0], self.train_cfg["num_intermediate_states"], self.denoising.num_gaussians
        if rands is None:
            rands = [draw_rollout_randoms(B, n, K, generator) for _ in range(self.train_cfg["nfe"])]
        return self.step.forward(prompt_embeds, pooled_prompt_embeds, grid_hw, noise, rands, iteration=it)

    def forward_backward(self, prompt_embeds, pooled_prompt_embeds, grid_hw, noise, running_status=None, rands=None,
                         generator=None, grads=None, neg_prompt_embeds=None):
        """loss, log_vars, grads (fp32 tensors keyed by adapter state-dict name, accumulated into when given)."""
        it = (running_status or {}).get("iteration", 0)
        B, n, K = noise.shape[0], self.train_cfg["num_intermediate_states"], self.denoising.num_gaussians
        if rands is None:
            rands = [draw_rollout_randoms(B, n, K, generator) for _ in range(self.train_cfg["nfe"])]
        return self.step.forward_backward(prompt_embeds, pooled_prompt_embeds, grid_hw, noise, rands, iteration=it,
                                          grads=grads, neg_txt=neg_prompt_embeds)
