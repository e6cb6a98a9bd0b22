"""`LatentDiffusionTextImage` — the model object the runner drives (lakonlab/models/latent_diffusion_text_image.py:13-106
over BaseDiffusion.train_fwd_bwd, lakonlab/models/base_diffusion.py:14-62, and BaseModel.train_step / step_optimizer,
lakonlab/models/base.py:76-103,162-189), on the native distillation step.

Kept: `train_step(data, optimizer, loss_scaler=None, running_status=None) -> dict(log_vars=..., num_samples=...)`; `data`
carries `prompt_embed_kwargs` (`encoder_hidden_states`, `pooled_projections`, and for Qwen `negative_prompt_embed_kwargs`)
and the dummy `latents` whose shape sets the resolution (data-free training). `optimizer` is the dict `{'diffusion': ...}`
of configs/*/_ddp_train.py; here its value is the `FlatAdamW` arena that `build_optimizers` made — clip, AdamW, EMA and the
bf16 write-back are one fused pass (the reference's EMA hook becomes a parameter of it).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from arcflow_b200.train import ArcFlowTrainer, draw_rollout_randoms


class LatentDiffusionTextImage:
    def __init__(self, student, teacher, train_cfg: Optional[Dict] = None, test_cfg: Optional[Dict] = None,
                 shift: float = 3.2, loss_scale: float = 30.0, policy_type: str = "ArcFlow", use_ema: bool = True):
        if policy_type != "ArcFlow":
            raise ValueError(f"Invalid policy: {policy_type}. Supported policies are ['ArcFlow'].")
        self.diffusion, self.teacher = student, teacher
        self.train_cfg, self.test_cfg = dict(train_cfg or {}), dict(test_cfg or {})
        self.shift, self.loss_scale, self.use_ema = shift, loss_scale, use_ema
        self.trainer: Optional[ArcFlowTrainer] = None
        self.generator = torch.Generator().manual_seed(0)

    # -- optimizer -------------------------------------------------------------------------------------------------
    def build_trainer(self, optimizer_cfg: Dict, lr_config: Optional[Dict] = None, ema_cfg: Optional[Dict] = None):
        """optimizer_cfg: the `'diffusion'` entry of configs/*/_ddp_train.py:18-26; lr_config :27-31; ema_cfg: the
        ExponentialMovingAverageHookMod entry of custom_hooks (arcflux_2nfe_k16.py:141-151)."""
        if self.teacher is None:
            raise ValueError("training needs a teacher")
        o, lr, ema = dict(optimizer_cfg or {}), dict(lr_config or {}), dict(ema_cfg or {})
        keys = (o.get("paramwise_cfg") or {}).get("custom_keys") or {}
        mult_key = next(iter(keys), "proj_out_loggamma")
        kw = dict(lr=o.get("lr", 1e-4), betas=tuple(o.get("betas", (0.9, 0.95))), eps=o.get("eps", 1e-8),
                  weight_decay=o.get("weight_decay", 0.0), lr_mult_key=mult_key,
                  lr_mult=(keys.get(mult_key) or {}).get("lr_mult", 1.0) if keys else 0.1,
                  max_norm=self.train_cfg.get("diffusion_grad_clip", 0.0),
                  clip_begin_iter=self.train_cfg.get("diffusion_grad_clip_begin_iter", 0),
                  clip_skip_ratio=self.train_cfg.get("diffusion_grad_clip_skip_ratio", 0.0),
                  warmup_iters=lr.get("warmup_iters", 0) if lr.get("warmup") else 0,
                  warmup_ratio=lr.get("warmup_ratio", 1.0),
                  ema_gamma=(ema.get("momentum_cfg") or {}).get("gamma", 7.0), ema_start_iter=ema.get("start_iter", 0),
                  # bitsandbytes' AdamW8bit keeps block-wise 8-bit moments; `optim_bits=32` is its own switch back to fp32 state
                  state_bits=int(o.get("optim_bits", 8 if o.get("type", "AdamW8bit") == "AdamW8bit" else 32)))
        step_cfg = {k: v for k, v in self.train_cfg.items() if not k.startswith("diffusion_grad_clip")}
        self.trainer = ArcFlowTrainer(self.diffusion, self.teacher, step_cfg, self.shift, self.loss_scale, **kw)
        return {"diffusion": self.trainer.opt}

    # -- one iteration ---------------------------------------------------------------------------------------------
    def train_step(self, data: Dict, optimizer=None, loss_scaler=None, running_status: Optional[Dict] = None) -> Dict:
        if self.trainer is None:
            raise RuntimeError("call build_trainer() (lakonlab.apis.train_model does) before train_step")
        if optimizer is not None and optimizer.get("diffusion") is not self.trainer.opt:
            raise ValueError("train_step: `optimizer['diffusion']` is not this model's arena optimizer")
        st = self.diffusion
        pe = data["prompt_embed_kwargs"]
        txt = pe["encoder_hidden_states"].to(st.device, torch.bfloat16)
        bs = txt.shape[0]
        latents = data["latents"]
        if latents.dim() != 4 or latents.shape[0] != bs:
            raise ValueError("`latents` must be (batch, 16, h, w) with the prompt batch size")
        grid = (latents.shape[2] // 2, latents.shape[3] // 2)
        noise = torch.randn((bs, grid[0] * grid[1], st.cfg.in_channels), generator=self.generator).to(st.device)
        K, n = st.num_gaussians, self.trainer.distill.cfg["num_intermediate_states"]
        rands = [draw_rollout_randoms(bs, n, K, self.generator) for _ in range(self.trainer.distill.cfg["nfe"])]
        it = int((running_status or {}).get("iteration", self.trainer.iteration))
        if getattr(st, "arch", "flux") == "qwen":
            neg = data.get("negative_prompt_embed_kwargs")
            if neg is None:
                raise ValueError("Either `negative_prompt_embed_kwargs` or `negative_prompt_kwargs` should be provided "
                                 "in the input data for classifier-free guidance.")
            loss, log_vars = self.trainer.train_step(txt, None, grid, noise, rands, iteration=it,
                                                     neg_txt=neg["encoder_hidden_states"].to(st.device, torch.bfloat16))
        else:
            pooled = pe["pooled_projections"].to(st.device, torch.bfloat16)
            loss, log_vars = self.trainer.train_step(txt, pooled, grid, noise, rands, iteration=it)
        log_vars = {k: (float(v) if isinstance(v, (int, float, bool)) else v) for k, v in log_vars.items()}
        log_vars.setdefault("loss", float(loss))
        return dict(log_vars=log_vars, num_samples=bs)

    # -- state -----------------------------------------------------------------------------------------------------
    def set_seed(self, seed: int):
        self.generator.manual_seed(seed)

    def state_dict(self, trainable_only: bool = True) -> Dict[str, torch.Tensor]:
        """Trainable tensors only (`ckpt_trainable_only=True`, configs/*/_ddp_train.py:34), under the reference's
        checkpoint prefixes `diffusion.denoising.` / `diffusion_ema.denoising.`."""
        if not trainable_only:
            raise NotImplementedError("the frozen base is never re-saved; checkpoints hold the adapter only")
        out = {}
        for k, v in self.trainer.adapter_state_dict(use_ema=False).items():
            out["diffusion.denoising." + k] = v
        if self.use_ema:
            for k, v in self.trainer.adapter_state_dict(use_ema=True).items():
                out["diffusion_ema.denoising." + k] = v
        return out
