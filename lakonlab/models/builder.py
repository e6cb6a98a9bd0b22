"""`build_model(cfg.model, train_cfg=, test_cfg=)` — what train.py:229-230 obtains from mmgen's registry, for the one model
family on the path: `LatentDiffusionTextImage` wrapping an `ArcFlowImitationDataFree` student (ArcFlux / ArcQwenImage
transformer) and a tied `GaussianFlow` teacher (configs/flux/arcflux_2nfe_k16.py:5-86, configs/qwen/arcqwen_2nfe_k16.py).

Weights: `pretrained=` accepts a local `.safetensors` file, a folder of shards or a `*.safetensors.index.json` (the
reference's `huggingface://` URIs cannot be fetched offline), or `synthetic://<seed>` for seeded random weights of the
configured shape. `pretrained_adapter=` (a folder in the on-disk adapter format) overrides the fresh adapter init.
"""
from __future__ import annotations

import glob
import json
import os
from typing import Dict, Optional

import torch

from arcflow_b200.adapter_init import flux_lora_target_paths, init_arcflow_adapter
from arcflow_b200.config import ArcFluxConfig
from arcflow_b200.qwen import ArcQwenConfig, qwen_lora_targets

STUDENT_TYPES = {"ArcFluxTransformer2DModel": "flux", "ArcQwenImageTransformer2DModel": "qwen"}
TEACHER_TYPES = {"FluxTransformer2DModel": "flux", "QwenImageTransformer2DModel": "qwen"}


def load_transformer_weights(pretrained: str) -> Dict[str, torch.Tensor]:
    from safetensors.torch import load_file
    if pretrained.startswith("huggingface://"):
        raise ValueError(f"'{pretrained}': Hub URIs cannot be resolved offline; point `pretrained` at a local "
                         f".safetensors file / shard folder, or use 'synthetic://<seed>'")
    if pretrained.endswith(".index.json"):
        with open(pretrained) as f:
            files = sorted(set(json.load(f)["weight_map"].values()))
        files = [os.path.join(os.path.dirname(pretrained), x) for x in files]
    elif os.path.isdir(pretrained):
        files = sorted(glob.glob(os.path.join(pretrained, "*.safetensors")))
    else:
        files = [pretrained]
    if not files:
        raise FileNotFoundError(f"no .safetensors under '{pretrained}'")
    sd: Dict[str, torch.Tensor] = {}
    for f in files:
        sd.update(load_file(f))
    return sd


def synthetic_base_state_dict(arch: str, cfg, seed: int, device) -> Dict[str, torch.Tensor]:
    """Seeded stock transformer (diffusers names, incl. its own `norm_out.linear` / `proj_out`) of the configured shape."""
    if arch == "flux":
        from arcflow_b200.synthetic import make_flux_state_dict, make_flux_teacher_extras
        base, extras = make_flux_state_dict(cfg, seed, device), make_flux_teacher_extras(cfg, seed + 1, device)
    else:
        from arcflow_b200.qwen import make_qwen_state_dict, make_qwen_teacher_extras
        base, extras = make_qwen_state_dict(cfg, seed, device), make_qwen_teacher_extras(cfg, seed + 1, device)
    base = {k: v for k, v in base.items() if "lora" not in k and not k.startswith(("proj_out_", "norm_out."))}
    base.update(extras)
    return base


def student_config(denoising: dict):
    arch = STUDENT_TYPES.get(denoising.get("type"))
    if arch is None:
        raise ValueError(f"unsupported denoising type '{denoising.get('type')}' (supported: {sorted(STUDENT_TYPES)})")
    cls = ArcFluxConfig if arch == "flux" else ArcQwenConfig
    fields = set(cls.__dataclass_fields__)
    kw = {k: (tuple(v) if k == "axes_dims_rope" else v) for k, v in denoising.items() if k in fields}
    if not denoising.get("use_lora", False):
        kw["lora_rank"] = 0
    return arch, cls(**kw)


def build_student_state_dict(arch: str, cfg, denoising: dict, device, seed: int):
    """Returns (student state dict = frozen base + adapter, teacher extras = the base's own norm_out / proj_out)."""
    pretrained = denoising.get("pretrained") or "synthetic://1234"
    g = torch.Generator().manual_seed(seed)
    if pretrained.startswith("synthetic://"):
        base = synthetic_base_state_dict(arch, cfg, int(pretrained[len("synthetic://"):] or 1234), device)
    else:
        base = load_transformer_weights(pretrained)
    extras = {k: base[k] for k in ("norm_out.linear.weight", "norm_out.linear.bias", "proj_out.weight", "proj_out.bias")}
    if arch == "flux":
        targets = flux_lora_target_paths(cfg, denoising.get("lora_target_modules") or ())
    else:
        targets = qwen_lora_targets(cfg)
    adapter = init_arcflow_adapter(base, cfg, targets, generator=g)
    if denoising.get("pretrained_adapter"):
        from lakonlab.pipelines.arcflow_loader import read_adapter_folder
        _, loaded = read_adapter_folder(denoising["pretrained_adapter"])
        adapter.update({k.replace(".default.weight", ".weight"): v for k, v in loaded.items()})
    sd = {k: v for k, v in base.items() if k not in ("proj_out.weight", "proj_out.bias")}
    sd.update({k: v.to(device) for k, v in adapter.items()})
    return sd, extras


def build_model(model_cfg: dict, train_cfg: Optional[dict] = None, test_cfg: Optional[dict] = None, device=None,
                seed: int = 0):
    from .latent_diffusion_text_image import LatentDiffusionTextImage
    if model_cfg.get("type") != "LatentDiffusionTextImage":
        raise ValueError(f"unsupported model type '{model_cfg.get('type')}'")
    diffusion = model_cfg["diffusion"]
    if diffusion.get("type") != "ArcFlowImitationDataFree":
        raise ValueError(f"unsupported diffusion type '{diffusion.get('type')}'")
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    arch, cfg = student_config(diffusion["denoising"])
    teacher_cfg = model_cfg.get("teacher")
    if teacher_cfg is not None:
        if TEACHER_TYPES.get(teacher_cfg["denoising"].get("type")) != arch:
            raise ValueError("teacher and student must be the same architecture (weights are tied)")
        if not model_cfg.get("tie_teacher", False):
            raise NotImplementedError("only tie_teacher=True is supported: the teacher borrows the student's frozen trunk")
    sd, extras = build_student_state_dict(arch, cfg, diffusion["denoising"], device, seed)
    if arch == "flux":
        from arcflow_b200.model import ArcFluxEngineModel, FluxTeacherEngine
        student = ArcFluxEngineModel(sd, cfg, device, consume_state_dict=True)
        teacher = FluxTeacherEngine(student, extras) if teacher_cfg is not None else None
    else:
        from arcflow_b200.qwen import ArcQwenEngineModel, QwenTeacherEngine
        student = ArcQwenEngineModel(sd, cfg, device, consume_state_dict=True)
        teacher = QwenTeacherEngine(student, extras) if teacher_cfg is not None else None
    del sd
    shift = (diffusion.get("timestep_sampler") or {}).get("shift", 3.2)
    loss_scale = ((diffusion.get("flow_loss") or {}).get("rescale_cfg") or {}).get("scale", 30.0)
    merged_train_cfg = dict(train_cfg or {})
    if "lora_dropout" in diffusion["denoising"]:
        merged_train_cfg.setdefault("lora_dropout", diffusion["denoising"]["lora_dropout"])
    return LatentDiffusionTextImage(student, teacher, train_cfg=merged_train_cfg, test_cfg=dict(test_cfg or {}),
                                    shift=shift, loss_scale=loss_scale, policy_type=diffusion.get("policy_type", "ArcFlow"),
                                    use_ema=model_cfg.get("diffusion_use_ema", True))
