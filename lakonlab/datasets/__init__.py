"""Prompt-embedding datasets for the data-free distillation (the reference's `ImagePrompts`,
lakonlab/datasets/image_prompts.py:93-391, reduced to what the train step consumes).

A sample is `dict(prompt_embed_kwargs=dict(encoder_hidden_states=[S_t, C], pooled_projections=[P]) ,
latents=[16, h, w] dummy zeros)` (+ `negative_prompt_embed_kwargs` when a negative embedding file is configured — the
Qwen teacher's true CFG). Cached embeddings are read from `.pt` / `.safetensors` files holding `prompt_embed_kwargs` (or
the legacy flat keys `prompt_embeds` / `pooled_prompt_embeds`, image_prompts.py:86-91); the reference's zstd-pickled
cache cannot be read here (no zstandard offline). `SyntheticPrompts` draws seeded embeddings of the configured shape.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, Optional, Tuple

import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler

PROMPT_KEY_MAPS = {"prompt_embeds": "encoder_hidden_states", "pooled_prompt_embeds": "pooled_projections",
                   "prompt_embeds_mask": "encoder_hidden_states_mask"}


def parse_prompt_embeds(data: Dict, pad_seq_len: Optional[int] = None) -> Dict[str, torch.Tensor]:
    pe = dict(data.get("prompt_embed_kwargs", {}))
    for old, new in PROMPT_KEY_MAPS.items():
        if old in data and new not in pe:
            pe[new] = data[old]
    scale = pe.pop("encoder_hidden_states_scale", None)
    if "encoder_hidden_states" in pe:
        x = pe["encoder_hidden_states"].float()
        if scale is not None:
            x = x * scale
        if pad_seq_len is not None:
            x = x[:pad_seq_len] if x.size(0) >= pad_seq_len else torch.cat(
                [x, x.new_zeros((pad_seq_len - x.size(0),) + x.shape[1:])], 0)
        pe["encoder_hidden_states"] = x
    if "pooled_projections" in pe:
        pe["pooled_projections"] = pe["pooled_projections"].float()
    return pe


def _load_any(path: str) -> Dict:
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


class ImagePrompts(Dataset):
    def __init__(self, cache_dir: str, negative_prompt_embeds_path: Optional[str] = None, pad_seq_len: Optional[int] = None,
                 latent_size: Tuple[int, int, int] = (16, 128, 128), repeat: int = 1, **_unused):
        self.files = sorted(glob.glob(os.path.join(cache_dir, "*.pt")) + glob.glob(os.path.join(cache_dir, "*.safetensors")))
        if not self.files:
            raise FileNotFoundError(f"no cached prompt embeddings (*.pt / *.safetensors) under '{cache_dir}'")
        self.pad_seq_len, self.latent_size, self.repeat = pad_seq_len, tuple(latent_size), repeat
        self.negative = (parse_prompt_embeds(_load_any(negative_prompt_embeds_path), pad_seq_len)
                         if negative_prompt_embeds_path else None)

    def __len__(self):
        return len(self.files) * self.repeat

    def __getitem__(self, i):
        out = dict(prompt_embed_kwargs=parse_prompt_embeds(_load_any(self.files[i % len(self.files)]), self.pad_seq_len),
                   latents=torch.zeros(self.latent_size))
        if self.negative is not None:
            out["negative_prompt_embed_kwargs"] = self.negative
        return out


class SyntheticPrompts(Dataset):
    """Seeded N(0, 1)*0.1 text embeddings / N(0, 1) pooled embeddings — for benchmarks and tests."""

    def __init__(self, joint_attention_dim: int, pooled_projection_dim: Optional[int] = 768, seq_len: int = 512,
                 latent_size: Tuple[int, int, int] = (16, 128, 128), length: int = 1 << 20, negative: bool = False,
                 seed: int = 0):
        self.c, self.p, self.s, self.latent_size, self.length = joint_attention_dim, pooled_projection_dim, seq_len, tuple(latent_size), length
        self.negative, self.seed = negative, seed

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1_000_003 + i)
        pe = dict(encoder_hidden_states=torch.randn(self.s, self.c, generator=g) * 0.1)
        if self.p:
            pe["pooled_projections"] = torch.randn(self.p, generator=g)
        out = dict(prompt_embed_kwargs=pe, latents=torch.zeros(self.latent_size))
        if self.negative:
            gn = torch.Generator().manual_seed(self.seed * 1_000_003 - 1)
            out["negative_prompt_embed_kwargs"] = dict(encoder_hidden_states=torch.randn(self.s, self.c, generator=gn) * 0.1)
        return out


DATASETS = {"ImagePrompts": ImagePrompts, "SyntheticPrompts": SyntheticPrompts}


def build_dataset(cfg: Dict):
    cfg = dict(cfg)
    t = cfg.pop("type")
    if t not in DATASETS:
        raise ValueError(f"unsupported dataset type '{t}' (supported: {sorted(DATASETS)})")
    return DATASETS[t](**cfg)


def build_dataloader(dataset, samples_per_gpu: int = 1, workers_per_gpu: int = 0, distributed: bool = False, seed: int = 0,
                     shuffle: bool = True, **_unused):
    sampler = None
    if distributed:
        sampler = DistributedSampler(dataset, shuffle=shuffle, seed=seed)
        shuffle = False
    g = torch.Generator().manual_seed(seed)
    return DataLoader(dataset, batch_size=samples_per_gpu, shuffle=shuffle, sampler=sampler, num_workers=workers_per_gpu,
                      drop_last=True, generator=g)
