"""`pipe.load_arcflow_adapter(...)` — the drop-in boundary of the inference path.

Contract kept from the reference's ArcFlowLoaderMixin.load_arcflow_adapter
(lakonlab/pipelines/arcflow_loader.py:45-51, :242-275): reads `config.json` (`_class_name` in
{ArcFluxTransformer2DModel, ...} + constructor args) and `diffusion_pytorch_model.safetensors` from a
local folder (`subfolder=` honoured; Hub ids cannot be fetched offline), overlays the non-LoRA adapter
tensors (3 heads, norm_out) on the base transformer's state dict, attaches the LoRA pairs, swaps
`getattr(pipe, target_module_name)` for the ArcFlow module, and returns `f"{target}_arcflow"` — or
`None` with a warning when the folder holds no `lora` keys.  What differs: the swapped-in module is an
`ArcFluxEngineModel` (packed weights + C-ABI engine handle), not a torch nn.Module.
"""
from __future__ import annotations

import json
import os
import warnings
from typing import Dict, Optional, Union

import torch

from arcflow_b200.config import ArcFluxConfig
from arcflow_b200.model import ArcFluxEngineModel
from arcflow_b200.qwen import ArcQwenConfig, ArcQwenEngineModel

# same keys as the reference's LOCAL_CLASS_MAPPING (arcflow_loader.py:30-33)
LOCAL_CLASS_MAPPING = {"ArcFluxTransformer2DModel": (ArcFluxEngineModel, ArcFluxConfig),
                       "ArcQwenImageTransformer2DModel": (ArcQwenEngineModel, ArcQwenConfig)}


def read_adapter_folder(path: Union[str, os.PathLike], subfolder: Optional[str] = None):
    folder = os.path.join(path, subfolder) if subfolder else str(path)
    cfg_file = os.path.join(folder, "config.json")
    w_file = os.path.join(folder, "diffusion_pytorch_model.safetensors")
    if not os.path.isfile(cfg_file) or not os.path.isfile(w_file):
        raise FileNotFoundError(
            f"'{folder}' must contain config.json and diffusion_pytorch_model.safetensors "
            f"(Hub ids cannot be resolved: this build runs offline)")
    with open(cfg_file) as f:
        config = json.load(f)
    from safetensors.torch import load_file
    return config, load_file(w_file)


def write_adapter_folder(path: Union[str, os.PathLike], cfg, adapter_sd: Dict[str, torch.Tensor]):
    """Writes the on-disk format export_arcflow_to_diffusers.py:104-127 produces (used by tests/tools)."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    cls_name = "ArcQwenImageTransformer2DModel" if isinstance(cfg, ArcQwenConfig) else "ArcFluxTransformer2DModel"
    config = dict(cfg.to_dict(), _class_name=cls_name)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(config, f, indent=2)
    save_file({k: v.contiguous() for k, v in adapter_sd.items()},
              os.path.join(path, "diffusion_pytorch_model.safetensors"),
              metadata={"policy_config": json.dumps({"type": "ArcFlow"})})


def split_adapter_keys(sd: Dict[str, torch.Tensor]):
    """The reference splits adapter tensors by `"lora" in key` (arcflow_loader.py:246-251)."""
    lora = {k: v for k, v in sd.items() if "lora" in k}
    other = {k: v for k, v in sd.items() if "lora" not in k}
    return other, lora


class ArcFlowLoaderMixin:

    def load_arcflow_adapter(self, pretrained_model_name_or_path: Union[str, os.PathLike],
                             target_module_name: str = "transformer", adapter_name: Optional[str] = None,
                             **kwargs) -> Optional[str]:
        subfolder = kwargs.pop("subfolder", None)
        config, adapter_sd = read_adapter_folder(pretrained_model_name_or_path, subfolder)
        cls_name = config.get("_class_name")
        if cls_name not in LOCAL_CLASS_MAPPING:
            raise ValueError(f"Unsupported ArcFlow adapter class '{cls_name}' "
                             f"(supported: {sorted(LOCAL_CLASS_MAPPING)})")
        base = getattr(self, target_module_name, None)
        if base is None or not hasattr(base, "state_dict"):
            raise ValueError(f"pipeline has no module '{target_module_name}' with a state_dict() to adapt")
        base_sd = dict(base.state_dict())
        # keys saved under the target module's name (`transformer.<key>`) are accepted as the reference does
        # (arcflow_loader.py:246-250 strips f"{target_module_name}." before splitting LoRA from non-LoRA tensors)
        adapter_sd = {k.removeprefix(target_module_name + "."): v for k, v in adapter_sd.items()}
        other, lora = split_adapter_keys(adapter_sd)
        if not lora:
            warnings.warn(f"No LoRA weights found in '{pretrained_model_name_or_path}'; adapter not loaded.")
            return None
        # stock `proj_out` is replaced by the three ArcFlow heads (arcflux.py:86-88)
        base_sd.pop("proj_out.weight", None)
        base_sd.pop("proj_out.bias", None)
        base_sd.update(other)
        # accept both the exported (`lora_A.weight`) and the peft-internal (`lora_A.default.weight`) names
        for k, v in lora.items():
            base_sd[k.replace(".default.weight", ".weight")] = v
        model_cls, cfg_cls = LOCAL_CLASS_MAPPING[cls_name]
        fields = set(cfg_cls.__dataclass_fields__)
        cfg = cfg_cls(**{k: (tuple(v) if k == "axes_dims_rope" else v) for k, v in config.items() if k in fields})
        rank = next(iter(v.shape[0] for k, v in lora.items() if "lora_A" in k))
        cfg.lora_rank = int(rank)
        device = getattr(base, "device", torch.device("cuda"))
        module = model_cls(base_sd, cfg, device=device, consume_state_dict=True)
        setattr(self, target_module_name, module)
        return adapter_name or f"{target_module_name}_arcflow"
