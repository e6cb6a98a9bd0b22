"""Training checkpoints (`iter_N.pth` + `latest.pth`), the trainable-only / half-precision variant the reference's runner
writes (`ckpt_trainable_only`, `ckpt_fp16`, `ckpt_fp16_ema`: configs/*/_ddp_train.py:32-38;
lakonlab/runner/checkpoint.py:491-534, dynamic_iter_based_runner.py:110-160).

Layout: `dict(meta=dict(iter=, epoch=, ...), state_dict={'diffusion.denoising.<adapter key>': bf16, 'diffusion_ema.…': bf16},
optimizer={'diffusion': dict(params, exp_avg, exp_avg_sq, ema : fp32 flat arenas, layout, steps_taken)})`.
The frozen base is never saved. The flat fp32 arenas make resume bit-exact.
"""
from __future__ import annotations

import os
import shutil
from typing import Dict, Optional

import torch


def exists_ckpt(path: Optional[str]) -> bool:
    return bool(path) and os.path.isfile(path)


def get_checkpoint(model, optimizer=None, meta: Optional[Dict] = None) -> Dict:
    ckpt = dict(meta=dict(meta or {}), state_dict={k: v.cpu() for k, v in model.state_dict(trainable_only=True).items()})
    if optimizer is not None:
        ckpt["optimizer"] = {k: opt.state_dict() for k, opt in optimizer.items()}
    if hasattr(model, "generator"):   # noise / roll-out draws continue where they stopped (the reference re-seeds)
        ckpt["rng_state"] = model.generator.get_state()
    return ckpt


def write_checkpoint_to_file(checkpoint: Dict, filepath: str, create_symlink: bool = True):
    os.makedirs(os.path.dirname(os.path.abspath(filepath)), exist_ok=True)
    tmp = filepath + ".tmp"
    torch.save(checkpoint, tmp)
    os.replace(tmp, filepath)
    if create_symlink:
        latest = os.path.join(os.path.dirname(filepath), "latest.pth")
        try:
            if os.path.lexists(latest):
                os.remove(latest)
            os.symlink(os.path.basename(filepath), latest)
        except OSError:
            shutil.copy(filepath, latest)


def load_checkpoint(filename: str, map_location="cpu") -> Dict:
    return torch.load(filename, map_location=map_location, weights_only=False)


def adapter_from_checkpoint(checkpoint: Dict, use_ema: bool = True) -> Dict[str, torch.Tensor]:
    """The adapter tensors of a training checkpoint under their on-disk names — what
    export_arcflow_to_diffusers.py:58-103 extracts (EMA weights preferred)."""
    sd = checkpoint.get("state_dict", checkpoint)
    for prefix in (("diffusion_ema.denoising.", "diffusion.denoising.") if use_ema else ("diffusion.denoising.",)):
        out = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
        if out:
            return out
    raise KeyError("checkpoint holds no 'diffusion(.ema).denoising.*' tensors")
