"""Training checkpoints (`iter_N.pth` + `latest.pth`), the trainable-only / half-precision variant the reference's runner
writes (`ckpt_trainable_only`, `ckpt_fp16`, `ckpt_fp16_ema`: configs/*/_ddp_train.py:32-38;
lakonlab/runner/checkpoint.py:491-534, dynamic_iter_based_runner.py:110-160).

Layout: `dict(meta=dict(iter=, epoch=, ...), state_dict={'diffusion.denoising.<adapter key>': bf16, 'diffusion_ema.…': bf16},
optimizer={'diffusion': dict(params, exp_avg, exp_avg_sq, ema : fp32 flat arenas, layout, steps_taken; with AdamW8bit also
state1, state2 : uint8 code arenas and absmax1, absmax2 : fp32 per 256-element block)})`.
The frozen base is never saved. The flat arenas make resume bit-exact.
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
        ckpt["rng_state"] = gather_rng_states(model.generator)
    return ckpt


def gather_rng_states(generator) -> Dict[int, torch.Tensor]:
    """{rank: generator state} of EVERY rank (collective: all ranks call get_checkpoint, rank 0 writes the file).
    train.py --diff_seed gives each rank its own noise / roll-out stream (train.py:74-77,222-225 of the reference);
    restoring rank 0's state everywhere would make all ranks draw identical x_T after a resume."""
    import torch.distributed as dist
    state = generator.get_state()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return {0: state}
    states = [None] * dist.get_world_size()
    dist.all_gather_object(states, state)
    return {r: s for r, s in enumerate(states)}


def restore_rng_state(generator, saved, rank: int, iteration: int = 0) -> str:
    """Restore THIS rank's stream. `saved` is the {rank: state} dict written by get_checkpoint (a bare tensor from an
    older checkpoint counts as rank 0's). A rank the checkpoint does not know (the world size grew) is re-seeded from
    (rank 0's seed, rank, iteration) instead of cloning another rank's stream. Returns what was done."""
    if isinstance(saved, torch.Tensor):
        saved = {0: saved}
    if rank in saved:
        generator.set_state(saved[rank])
        return "restored"
    probe = torch.Generator()
    probe.set_state(saved[min(saved)])
    generator.manual_seed((probe.initial_seed() + 1000003 * rank + 7919 * int(iteration)) % (1 << 63))
    return "reseeded"


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
