"""`DynamicIterBasedRunnerMod` — the iteration loop that drives `model.train_step`
(lakonlab/runner/dynamic_iter_based_runner.py:44-219 over mmgen's DynamicIterBasedRunner).

Kept: `run(data_loaders, workflow)` with `('train', n)` flows until `max_iters`; `pass_training_status` (train_step receives
`running_status=dict(iteration=, epoch=)`); `ckpt_trainable_only / ckpt_fp16 / ckpt_fp16_ema` (always on here: only the bf16
adapter + fp32 optimizer arenas exist to be saved); `gc_interval`; `save_checkpoint(out_dir, filename_tmpl='iter_{}.pth')`
with the `latest.pth` link, `resume(checkpoint)`; hook call points `before_run / after_train_iter / after_run`.
"""
from __future__ import annotations

import gc
import logging
import os
import time
from typing import Dict, List, Optional

import torch

from .checkpoint import get_checkpoint, load_checkpoint, write_checkpoint_to_file


class IterLoader:
    def __init__(self, dataloader):
        self._dataloader, self._epoch = dataloader, 0
        self._iter = iter(dataloader)

    @property
    def epoch(self):
        return self._epoch

    def __next__(self):
        try:
            return next(self._iter)
        except StopIteration:
            self._epoch += 1
            if hasattr(self._dataloader.sampler, "set_epoch"):
                self._dataloader.sampler.set_epoch(self._epoch)
            self._iter = iter(self._dataloader)
            return next(self._iter)


class DynamicIterBasedRunnerMod:
    def __init__(self, model, optimizer: Optional[Dict] = None, work_dir: Optional[str] = None,
                 logger: Optional[logging.Logger] = None, meta: Optional[Dict] = None, max_iters: Optional[int] = None,
                 pass_training_status: bool = False, ckpt_trainable_only: bool = True, ckpt_fp16: bool = True,
                 ckpt_fp16_ema: bool = True, ckpt_bf16_optim: bool = False, gc_interval: int = -1, **_unused):
        self.model, self.optimizer, self.work_dir, self.meta = model, optimizer, work_dir, meta
        self.logger = logger or logging.getLogger("lakonlab")
        self._max_iters, self._iter, self._epoch, self._inner_iter = max_iters, 0, 0, 0
        self.pass_training_status = pass_training_status
        if not ckpt_trainable_only:
            raise NotImplementedError("ckpt_trainable_only=False: the frozen base is not re-saved by this build")
        self.gc_interval = gc_interval
        self.manual_gc = isinstance(gc_interval, int) and gc_interval > 0
        self.hooks: List = []
        self.outputs: Dict = {}
        self.log_buffer: List[Dict] = []
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0

    iter = property(lambda self: self._iter)
    epoch = property(lambda self: self._epoch)
    max_iters = property(lambda self: self._max_iters)

    def register_hook(self, hook):
        self.hooks.append(hook)

    def call_hook(self, name: str):
        for h in self.hooks:
            fn = getattr(h, name, None)
            if fn is not None:
                fn(self)

    def train(self, data_loader: IterLoader, **kwargs):
        data_batch = next(data_loader)
        self._epoch = data_loader.epoch
        if self.pass_training_status:
            kwargs["running_status"] = dict(iteration=self._iter, epoch=self._epoch)
        outputs = self.model.train_step(data_batch, self.optimizer, **kwargs)
        if not isinstance(outputs, dict) or "log_vars" not in outputs:
            raise TypeError("model.train_step() must return a dict with 'log_vars' and 'num_samples'")
        self.outputs = outputs
        self.log_buffer.append(outputs["log_vars"])
        self.call_hook("after_train_iter")
        self._inner_iter += 1
        self._iter += 1

    def run(self, data_loaders, workflow, max_iters: Optional[int] = None, **kwargs):
        assert isinstance(data_loaders, list) and len(data_loaders) == len(workflow)
        if max_iters is not None:
            self._max_iters = max_iters
        assert self._max_iters is not None, "max_iters must be specified during instantiation"
        self.logger.info("Start running, work_dir: %s", self.work_dir if self.work_dir is not None else "NONE")
        self.logger.info("workflow: %s, max: %d iters", workflow, self._max_iters)
        self.call_hook("before_run")
        iter_loaders = [IterLoader(x) for x in data_loaders]
        if self.manual_gc:
            gc.disable()
        try:
            while self._iter < self._max_iters:
                for i, (mode, iters) in enumerate(workflow):
                    if not isinstance(mode, str) or not hasattr(self, mode):
                        raise ValueError(f'runner has no method named "{mode}" to run a workflow')
                    self._inner_iter = 0
                    for _ in range(iters):
                        if mode == "train" and self._iter >= self._max_iters:
                            break
                        if self.manual_gc and self._inner_iter % self.gc_interval == 0:
                            gc.collect()
                        getattr(self, mode)(iter_loaders[i], **kwargs)
        finally:
            if self.manual_gc:
                gc.enable()
        self.call_hook("after_run")

    def save_checkpoint(self, out_dir: str, filename_tmpl: str = "iter_{}.pth", meta: Optional[Dict] = None,
                        save_optimizer: bool = True, create_symlink: bool = True):
        meta = dict(meta or {})
        meta.update(iter=self._iter, epoch=self._epoch)
        if self.meta is not None:
            meta.update({k: v for k, v in self.meta.items() if isinstance(v, (str, int, float, bool))})
        ckpt = get_checkpoint(self.model, self.optimizer if save_optimizer else None, meta)
        if self.rank == 0:
            path = os.path.join(out_dir, filename_tmpl.format(self._iter))
            write_checkpoint_to_file(ckpt, path, create_symlink)
            return path

    def resume(self, checkpoint: str, resume_optimizer: bool = True, map_location="cpu"):
        ckpt = load_checkpoint(checkpoint, map_location)
        self._iter = self._inner_iter = int(ckpt["meta"]["iter"])
        self._epoch = int(ckpt["meta"].get("epoch", 0))
        if resume_optimizer and "optimizer" in ckpt and self.optimizer is not None:
            for k, opt in self.optimizer.items():
                opt.load_state_dict(ckpt["optimizer"][k])
            self.model.trainer.iteration = self._iter
            self.model.trainer.write_back()
        else:
            from .checkpoint import adapter_from_checkpoint
            self.model.trainer.opt.load_params({k: v.to(self.model.diffusion.device)
                                                for k, v in adapter_from_checkpoint(ckpt, use_ema=False).items()})
            self.model.trainer.write_back()
        if resume_optimizer and "rng_state" in ckpt and hasattr(self.model, "generator"):
            from .checkpoint import restore_rng_state
            restore_rng_state(self.model.generator, ckpt["rng_state"], self.rank, self._iter)
        self.logger.info("resumed from epoch: %d, iter %d", self._epoch, self._iter)
        return ckpt


class TextLoggerHook:
    def __init__(self, interval: int = 1, **_unused):
        self.interval, self.t0 = interval, None

    def before_run(self, runner):
        self.t0 = time.time()

    def after_train_iter(self, runner):
        if runner.rank != 0 or (runner.iter + 1) % self.interval:
            return
        lv = runner.outputs["log_vars"]
        txt = ", ".join(f"{k}: {v:.5g}" if isinstance(v, float) else f"{k}: {v}" for k, v in lv.items())
        runner.logger.info("Iter [%d/%d] %s, time: %.1fs", runner.iter + 1, runner.max_iters, txt, time.time() - self.t0)


class CheckpointHook:
    def __init__(self, interval: int = -1, out_dir: Optional[str] = None, max_keep_ckpts: int = -1,
                 must_save_interval: Optional[int] = None, by_epoch: bool = False, **_unused):
        self.interval, self.out_dir, self.max_keep, self.must_save = interval, out_dir, max_keep_ckpts, must_save_interval
        self.saved: List[str] = []

    def after_train_iter(self, runner):
        it = runner.iter + 1
        if self.interval <= 0 or it % self.interval:
            return
        runner._iter += 1   # checkpoints are named by completed iterations (the hook runs before the counter advances)
        try:
            path = runner.save_checkpoint(self.out_dir or runner.work_dir)
        finally:
            runner._iter -= 1
        if path and self.max_keep > 0:
            keep_forever = self.must_save and it % self.must_save == 0
            if not keep_forever:
                self.saved.append(path)
            while len(self.saved) > self.max_keep:
                old = self.saved.pop(0)
                if os.path.isfile(old):
                    os.remove(old)
