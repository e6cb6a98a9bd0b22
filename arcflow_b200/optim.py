"""Step glue after the backward: DDP gradient all-reduce, grad clip, AdamW, Karras EMA — over ONE flat fp32 arena.

Reference:
  BaseModel.step_optimizer                     lakonlab/models/base.py:76-103       clip 50.0 from iteration 100, skip on NaN/Inf
  optimizer / lr config                        configs/flux/_ddp_train.py:13-31     AdamW8bit lr 1e-4 betas (.9,.95) wd 0,
                                                                                    proj_out_loggamma lr x 0.1, linear warm-up 100 it from 1e-3
  ExponentialMovingAverageHookMod (Karras)     lakonlab/runner/hooks/ema_hook.py:86-121, configs/flux/arcflux_2nfe_k16.py:135-145
  DDP all-reduce of the trainable submodule    lakonlab/parallel/ddp_wrapper.py:7-26
All trainable tensors live back to back in one buffer, so the all-reduce is ONE NCCL call on the gradient arena and the
update is two kernel launches (afb_grad_norm_sq, afb_adamw_ema_step) regardless of how many tensors there are.
bitsandbytes' 8-bit optimizer state is not reproduced (fp32 state, torch.optim.AdamW arithmetic).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import AfbError


def warmup_lr(base_lr: float, iteration: int, warmup_iters: int = 100, warmup_ratio: float = 0.001) -> float:
    """mmcv 'linear' warm-up of a fixed-policy LR (SURVEY.md Appendix A.9 iv)."""
    if iteration >= warmup_iters:
        return base_lr
    k = (1 - iteration / warmup_iters) * (1 - warmup_ratio)
    return base_lr * (1 - k)


def karras_momentum(iteration: int, start_iter: int = 100, gamma: float = 7.0, max_momentum: float = 1.0) -> float:
    t = max(iteration + 1 - start_iter, 1)
    return min((1 - 1 / t) ** (gamma + 1), max_momentum)


class FlatAdamW:
    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device, lr: float = 1e-4, betas=(0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.0, lr_mult_key: str = "proj_out_loggamma", lr_mult: float = 0.1,
                 max_norm: float = 50.0, clip_begin_iter: int = 100, clip_skip_ratio: float = 0.0, warmup_iters: int = 100, warmup_ratio: float = 0.001,
                 ema_gamma: float = 7.0, ema_start_iter: int = 100):
        self.lib = _lib.load()
        self.device = torch.device(device)
        # tensors matching lr_mult_key are laid out contiguously so one [begin, end) range carries the multiplier
        names = sorted(shapes, key=lambda n: (lr_mult_key not in n, n))
        self.views, off = {}, 0
        self.lo = [0, 0]
        for n in names:
            size = 1
            for d in shapes[n]:
                size *= d
            if lr_mult_key in n:
                self.lo[1] = off + size
            self.views[n] = (off, tuple(shapes[n]))
            off += (size + 3) // 4 * 4   # keep every tensor 16-byte aligned
        self.n = off
        z = lambda dt=torch.float32: torch.zeros(self.n, dtype=dt, device=self.device)
        self.params, self.grads, self.exp_avg, self.exp_avg_sq, self.ema = z(), z(), z(), z(), z()
        self.shadow = z(torch.bfloat16)
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.skipped = torch.zeros(1, dtype=torch.int32, device=self.device)
        # partials + ticket of the deterministic norm reduction: one per optimizer instance (never shared across streams)
        self.norm_scratch = torch.zeros(max(int(self.lib.afb_grad_norm_scratch_floats()), 1), dtype=torch.float32,
                                        device=self.device) if self.device.type == "cuda" else None
        self.hp = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, lr_mult=lr_mult, max_norm=max_norm,
                       clip_begin_iter=clip_begin_iter, clip_skip_ratio=clip_skip_ratio, warmup_iters=warmup_iters, warmup_ratio=warmup_ratio,
                       ema_gamma=ema_gamma, ema_start_iter=ema_start_iter)
        self.steps_taken = 0

    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self.views[name]
        size = 1
        for d in shape:
            size *= d
        return buf[off:off + size].view(shape)

    def param(self, name):
        return self.view(self.params, name)

    def grad(self, name):
        return self.view(self.grads, name)

    def load_params(self, tensors: Dict[str, torch.Tensor]):
        for n, t in tensors.items():
            self.param(n).copy_(t.to(torch.float32))
        self.ema.copy_(self.params)
        self.shadow.copy_(self.params)

    def state_dict(self) -> Dict:
        """fp32 arenas + layout, for bit-exact resume (lakonlab/runner/checkpoint.py)."""
        return dict(params=self.params.cpu(), exp_avg=self.exp_avg.cpu(), exp_avg_sq=self.exp_avg_sq.cpu(), ema=self.ema.cpu(),
                    steps_taken=self.steps_taken, layout={n: (o, list(s)) for n, (o, s) in self.views.items()}, hp=dict(self.hp))

    def load_state_dict(self, sd: Dict):
        layout = {n: (o, tuple(s)) for n, (o, s) in sd["layout"].items()}
        if layout != self.views:
            raise AfbError("optimizer checkpoint was written for a different set of adapter tensors")
        for name in ("params", "exp_avg", "exp_avg_sq", "ema"):
            getattr(self, name).copy_(sd[name].to(self.device))
        self.shadow.copy_(self.params)
        self.steps_taken = int(sd["steps_taken"])

    def all_reduce_grads(self):
        """DDP semantics: gradients averaged over ranks; one collective on the whole arena (NCCL on GPU, gloo in CPU tests)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grads)
            self.grads.div_(dist.get_world_size())

    @torch.no_grad()
    def step(self, iteration: int) -> Dict[str, float]:
        hp = self.hp
        if self.device.type != "cuda":
            raise AfbError("FlatAdamW.step needs CUDA tensors (no CPU fallback exists)")
        stream = torch.cuda.current_stream().cuda_stream
        self.all_reduce_grads()
        clip = hp["max_norm"] > 0 and iteration >= hp["clip_begin_iter"]
        _lib.check(self.lib.afb_grad_norm_sq_ws(self.grads.data_ptr(), self.n, self.norm_sq.data_ptr(),
                                                self.norm_scratch.data_ptr(), self.norm_scratch.numel(), stream), "afb_grad_norm_sq_ws")
        a = _lib.AdamwArgs()
        a.params, a.grads = self.params.data_ptr(), self.grads.data_ptr()
        a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        a.ema, a.bf16_shadow, a.n = self.ema.data_ptr(), self.shadow.data_ptr(), self.n
        a.lr = warmup_lr(hp["lr"], iteration, hp["warmup_iters"], hp["warmup_ratio"])
        a.beta1, a.beta2, a.eps, a.weight_decay = hp["betas"][0], hp["betas"][1], hp["eps"], hp["weight_decay"]
        a.step = self.steps_taken + 1
        a.max_norm = hp["max_norm"] if clip else 0.0
        ratio = hp.get("clip_skip_ratio", 0.0)
        a.skip_norm = hp["max_norm"] * ratio if (clip and ratio > 0) else 0.0
        a.grad_norm_sq, a.skipped = self.norm_sq.data_ptr(), self.skipped.data_ptr()
        a.ema_copy = int(iteration < hp["ema_start_iter"])
        a.ema_momentum = karras_momentum(iteration, hp["ema_start_iter"], hp["ema_gamma"])
        a.lr_mult_begin, a.lr_mult_end, a.lr_mult = self.lo[0], self.lo[1], hp["lr_mult"]
        _lib.check(self.lib.afb_adamw_ema_step(C.byref(a), stream), "afb_adamw_ema_step")
        skipped = bool(self.skipped.item()) if clip else False
        if not skipped:
            self.steps_taken += 1
        norm = float(self.norm_sq.sqrt().item())
        return {"diffusion_grad_norm": float("nan") if skipped else norm, "skipped": skipped, "lr": a.lr}
