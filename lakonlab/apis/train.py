"""`train_model(model, dataset, cfg, distributed=, validate=, timestamp=, meta=)` — lakonlab/apis/train.py:20-166: data
loaders from `cfg.data`, the optimizer dict from `cfg.optimizer`, the runner from `cfg.runner`, lr / checkpoint / logger /
custom hooks, resume-or-load, then `runner.run(data_loaders, cfg.workflow, cfg.total_iters)`.

Data parallelism: every rank holds the full model (the module_wrapper='ddp' configuration); the gradient exchange is the
single all-reduce of the flat gradient arena inside the optimizer step, so there is no wrapper object."""
from __future__ import annotations

import logging
import os

from lakonlab.datasets import build_dataloader
from lakonlab.runner import CheckpointHook, DynamicIterBasedRunnerMod, TextLoggerHook, exists_ckpt
from lakonlab.runner.hooks import ExponentialMovingAverageHookMod


def train_model(model, dataset, cfg, distributed=False, validate=False, timestamp=None, meta=None):
    logger = logging.getLogger("lakonlab")
    dataset = dataset if isinstance(dataset, (list, tuple)) else [dataset]
    loader_cfg = {k: v for k, v in cfg.data.items()
                  if k not in ("train", "val", "test", "train_dataloader", "val_dataloader", "test_dataloader")}
    loader_cfg = {**loader_cfg, **cfg.data.get("train_dataloader", {})}
    loader_cfg.update(distributed=distributed, seed=cfg.get("seed", 0) or 0)
    data_loaders = [build_dataloader(ds, **loader_cfg) for ds in dataset]

    if distributed and cfg.get("module_wrapper", "ddp") != "ddp":
        raise NotImplementedError("module_wrapper must be 'ddp' (FSDP wrappers are out of scope: the adapter fits one GPU)")

    ema_hooks = [ExponentialMovingAverageHookMod(**{k: v for k, v in h.items() if k != "type"})
                 for h in cfg.get("custom_hooks", []) if h.get("type") == "ExponentialMovingAverageHookMod"]
    opt_cfg = dict(cfg.optimizer["diffusion"])
    if opt_cfg.get("type") not in ("AdamW8bit", "AdamW"):
        raise ValueError(f"unsupported optimizer type '{opt_cfg.get('type')}'")
    optimizer = model.build_trainer(opt_cfg, cfg.get("lr_config"), ema_hooks[0].to_trainer_cfg() if ema_hooks else None)

    runner_cfg = dict(cfg.runner)
    if runner_cfg.pop("type", "DynamicIterBasedRunnerMod") != "DynamicIterBasedRunnerMod":
        raise ValueError("runner type must be 'DynamicIterBasedRunnerMod'")
    runner = DynamicIterBasedRunnerMod(model, optimizer=optimizer, work_dir=cfg.work_dir, logger=logger, meta=meta,
                                       max_iters=cfg.total_iters, **runner_cfg)
    runner.timestamp = timestamp
    for h in ema_hooks:
        runner.register_hook(h)
    for h in (cfg.get("log_config") or {}).get("hooks", []):
        if h.get("type") == "TextLoggerHook":
            runner.register_hook(TextLoggerHook(interval=cfg.log_config.get("interval", 1)))
    if cfg.get("checkpoint_config"):
        ck = dict(cfg.checkpoint_config)
        out_dir = ck.pop("out_dir", None)
        if out_dir:   # the reference nests checkpoints under out_dir/<config name>
            out_dir = os.path.join(out_dir, cfg.get("name", os.path.basename(cfg.work_dir)))
        runner.register_hook(CheckpointHook(out_dir=out_dir or cfg.work_dir, **ck))

    if cfg.get("resume_from") and exists_ckpt(cfg.resume_from):
        runner.resume(cfg.resume_from)
    elif cfg.get("load_from"):
        runner.resume(cfg.load_from, resume_optimizer=False)
        runner._iter = runner._inner_iter = 0
    runner.run(data_loaders, [tuple(w) for w in cfg.workflow], cfg.total_iters)
    return runner
