"""Model hyper-parameters of the ArcFlow student / teacher transformers.

Field names follow the reference constructors (`_ArcFluxTransformer2DModel.__init__`,
lakonlab/models/architecture/arcflow/arcflux.py:28-40; `_ArcQwenImageTransformer2DModel.__init__`,
arcqwen.py:26-38) and configs/flux/arcflux_2nfe_k16.py:27-48 / configs/qwen/arcqwen_2nfe_k16.py:35-56.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict, field
from typing import Tuple

FLUX_LORA_TARGETS = (
    "proj_mlp", "proj_out", "ff.net.0.proj", "ff.net.2", "ff_context.net.0.proj", "ff_context.net.2",
    "timestep_embedder.linear_1", "timestep_embedder.linear_2")


@dataclass
class ArcFluxConfig:
    num_gaussians: int = 16
    logweights_channels: int = 4
    in_channels: int = 64
    out_channels: int = 64
    num_layers: int = 19
    num_single_layers: int = 38
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 4096
    pooled_projection_dim: int = 768
    guidance_embeds: bool = True
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    lora_rank: int = 256          # 0: no adapter
    mlp_ratio: int = 4

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def mlp_dim(self) -> int:
        return self.mlp_ratio * self.inner_dim

    @property
    def head_dims(self):
        k = self.num_gaussians
        return k * self.out_channels, k * self.logweights_channels, (k - 1) * self.logweights_channels

    def to_dict(self):
        d = asdict(self)
        d["axes_dims_rope"] = list(self.axes_dims_rope)
        return d


def flux_dev() -> ArcFluxConfig:
    """ArcFlow-FLUX (FLUX.1-dev trunk): 19 double + 38 single blocks, D 3072, rank-256 LoRA."""
    return ArcFluxConfig()


def flux_tiny(num_layers: int = 2, num_single_layers: int = 2, heads: int = 2) -> ArcFluxConfig:
    """Depth/width-reduced FLUX of the same structure, for CPU-oracle parity tests."""
    return ArcFluxConfig(num_layers=num_layers, num_single_layers=num_single_layers,
                         num_attention_heads=heads, joint_attention_dim=256, pooled_projection_dim=256)
