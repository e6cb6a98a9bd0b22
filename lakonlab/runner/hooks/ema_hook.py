"""`ExponentialMovingAverageHookMod` (lakonlab/runner/hooks/ema_hook.py:32-121): Karras-schedule EMA of the trainable
tensors, momentum beta_t = (1 - 1/t)^(gamma + 1), plain copy before `start_iter`.

In this build the EMA update is fused into the optimizer pass (`afb_adamw_ema_step` updates params, moments, EMA and the
bf16 shadow in one sweep over the arena), so the hook only carries the schedule into `LatentDiffusionTextImage.build_trainer`
and reports the momentum it implies; `after_train_iter` has nothing left to do."""
from __future__ import annotations

from arcflow_b200.optim import karras_momentum


class ExponentialMovingAverageHookMod:
    def __init__(self, module_keys=("diffusion_ema",), interp_mode="lerp", interval=1, start_iter=0,
                 momentum_policy="karras", momentum_cfg=None, priority="VERY_HIGH", **_unused):
        if interp_mode != "lerp":    # the fused pass computes ema = lerp(param, ema, momentum) — mmgen's default interpolation
            raise NotImplementedError(f"interp_mode='{interp_mode}': only 'lerp' is fused into the optimizer pass")
        if momentum_policy != "karras":
            raise NotImplementedError("only momentum_policy='karras' is fused into the optimizer pass")
        if interval != 1:
            raise NotImplementedError("the fused EMA runs every iteration (interval=1)")
        self.module_keys, self.start_iter = tuple(module_keys), start_iter
        self.momentum_cfg = dict(momentum_cfg or {})
        self.priority = priority

    def to_trainer_cfg(self):
        return dict(start_iter=self.start_iter, momentum_cfg=self.momentum_cfg)

    def karras(self, runner, gamma=7.0, max_momentum=1.0):
        return karras_momentum(runner.iter, self.start_iter, gamma, max_momentum)

    def before_run(self, runner):
        pass

    def after_train_iter(self, runner):
        pass
