"""Prompt-embedding datasets for the data-free distillation (the reference's `ImagePrompts`,
lakonlab/datasets/image_prompts.py:93-391, reduced to what the train step consumes).

A sample is `dict(prompt_embed_kwargs=dict(encoder_hidden_states=[S_t, C], pooled_projections=[P]) ,
latents=[16, h, w] dummy zeros)` (+ `negative_prompt_embed_kwargs` when a negative embedding file is configured — the
Qwen teacher's true CFG). Cached embeddings are read from the reference's zstd-pickled cache
(`<cache_dir>/<name>.zst`, image_prompts.py:357-391 — coded through libzstd, see zstd_cache.py) or from `.pt` /
`.safetensors` files holding `prompt_embed_kwargs` (or the legacy flat keys `prompt_embeds` / `prompt_embeds_scale` /
`pooled_prompt_embeds` / `prompt_embeds_mask`, image_prompts.py:86-91). `SyntheticPrompts` draws seeded embeddings of the configured shape.
"""
from __future__ import annotations

import glob
import os
from typing import Dict, Optional, Tuple

import torch
from torch.utils.data import DataLoader, Dataset, DistributedSampler

PROMPT_KEY_MAPS = {"prompt_embeds": "encoder_hidden_states", "prompt_embeds_scale": "encoder_hidden_states_scale",
                   "pooled_prompt_embeds": "pooled_projections", "prompt_embeds_mask": "encoder_hidden_states_mask"}


def pad_prompt_embeds(x: torch.Tensor, pad_seq_len: Optional[int]) -> torch.Tensor:
    """Pad with zeros / truncate along the sequence dimension (image_prompts.py:272-279)."""
    if pad_seq_len is None:
        return x
    if x.size(0) > pad_seq_len:
        return x[:pad_seq_len]
    return torch.cat([x, x.new_zeros((pad_seq_len - x.size(0),) + x.shape[1:])], 0)


def parse_prompt_embeds(data: Dict, pad_seq_len: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """image_prompts.py:281-309: legacy flat keys are mapped unless the new key exists, `encoder_hidden_states` is
    returned as fp32 x its stored scale (caches keep fp8/bf16 payloads + a scale), padded to `pad_seq_len` like the mask."""
    pe = dict(data.get("prompt_embed_kwargs", {}))
    for old, new in PROMPT_KEY_MAPS.items():
        if old in data and new not in pe:
            pe[new] = data[old]
    scale = pe.pop("encoder_hidden_states_scale", None)
    if "encoder_hidden_states" in pe:
        x = pe["encoder_hidden_states"].float()
        if scale is not None:
            x = x * scale
        pe["encoder_hidden_states"] = pad_prompt_embeds(x, pad_seq_len)
    if "pooled_projections" in pe:
        pe["pooled_projections"] = pe["pooled_projections"].float()
    if "encoder_hidden_states_mask" in pe:
        pe["encoder_hidden_states_mask"] = pad_prompt_embeds(pe["encoder_hidden_states_mask"], pad_seq_len)
    return pe


def _load_any(path: str) -> Dict:
    if path.endswith(".zst"):
        from .zstd_cache import loads_record
        with open(path, "rb") as f:
            return loads_record(f.read())
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


class ImagePrompts(Dataset):
    """Cache mode of the reference's `ImagePrompt` dataset (image_prompts.py:28-391): `<data_root>/<cache_dir>/<name>.zst`
    zstd-pickled records listed by `cache_datalist_path` (or the directory listing); `.pt` / `.safetensors` files with the
    same keys are accepted too. A sample carries `prompt_embed_kwargs`, `name`, `ids` and `latents` (cached latents x
    `latents_scale`, else an uninitialised buffer of the record's `latent_size` / the default) or, in `test_mode`,
    per-index seeded `noise`. The prompt-dataset / image modes (HF datasets, image folders) are outside this build."""

    def __init__(self, cache_dir: str, data_root: Optional[str] = None, cache_datalist_path: Optional[str] = None,
                 ignore_cached_latents: bool = False, negative_prompt_embeds_path: Optional[str] = None,
                 pad_seq_len: Optional[int] = None, latent_size: Tuple[int, int, int] = (16, 128, 128), repeat: int = 1,
                 start_ind: Optional[int] = None, end_ind: Optional[int] = None, bucketize: bool = False,
                 test_mode: bool = False, **_unused):
        from .zstd_cache import parse_datalist
        self.dir = os.path.join(data_root, cache_dir) if data_root else cache_dir
        if not os.path.isdir(self.dir):
            raise FileNotFoundError(f"cache directory '{self.dir}' does not exist")
        names, bucket_ids = parse_datalist(self.dir, cache_datalist_path, bucketize=bucketize)
        present = {}
        for f in os.listdir(self.dir):
            stem, ext = os.path.splitext(f)
            if ext in (".zst", ".pt", ".safetensors"):
                present.setdefault(stem, ext)
        if cache_datalist_path is None or not os.path.isfile(cache_datalist_path):
            names = [n for n in names if n in present]
        self.names = names
        self.ext = present
        if not self.names:
            raise FileNotFoundError(f"no cached prompt embeddings (*.zst / *.pt / *.safetensors) under '{self.dir}'")
        n = len(self.names)
        start_ind = max(min(start_ind, n - 1), -n) % n if start_ind is not None else 0
        end_ind = max(min(end_ind - 1, n - 1), -n) % n + 1 if end_ind is not None else n
        if not start_ind < end_ind:
            raise ValueError("Invalid start_ind and end_ind.")
        self.start_ind, self.end_ind = start_ind, end_ind
        self.pad_seq_len, self.latent_size, self.repeat = pad_seq_len, tuple(latent_size), repeat
        self.ignore_cached_latents, self.test_mode, self.bucketize = ignore_cached_latents, test_mode, bucketize
        if bucketize:
            self.bucket_ids = [bucket_ids[self._map_idx(i)] for i in range(len(self))]
        self.negative = (parse_prompt_embeds(_load_any(negative_prompt_embeds_path), pad_seq_len)
                         if negative_prompt_embeds_path else None)

    def _map_idx(self, idx: int) -> int:
        return self.start_ind + (idx // self.repeat)

    def __len__(self):
        return self.repeat * (self.end_ind - self.start_ind)

    def __getitem__(self, idx):
        name = self.names[self._map_idx(idx)]
        raw = _load_any(os.path.join(self.dir, name + self.ext.get(name, ".zst")))
        out = dict(ids=idx, name=raw.get("prompt", name), prompt_embed_kwargs=parse_prompt_embeds(raw, self.pad_seq_len))
        if not self.ignore_cached_latents and "latents" in raw:
            if self.test_mode:
                out["noise"] = torch.randn(raw["latents"].size(), dtype=torch.float32,
                                           generator=torch.Generator().manual_seed(idx))
            else:
                lat = raw["latents"].float()
                if raw.get("latents_scale") is not None:
                    lat = lat * raw["latents_scale"]
                out["latents"] = lat
        else:
            size = tuple(raw.get("latent_size", self.latent_size)) if not self.ignore_cached_latents else self.latent_size
            if self.test_mode:
                out["noise"] = torch.randn(size, dtype=torch.float32, generator=torch.Generator().manual_seed(idx))
            else:   # data-free training only reads the shape (the reference hands out torch.empty)
                out["latents"] = torch.zeros(size, dtype=torch.float32)
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
