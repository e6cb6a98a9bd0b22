"""Host-side timestep schedule of the ArcFlow sampler.

Follows `retrieve_raw_timesteps` (lakonlab/pipelines/arcflux_pipeline.py:34-70), the fixed-shift warp
`sigma = s*r / (1 + (s-1)*r)` (ContinuousTimeStepSampler.warp_t, lakonlab/models/diffusions/sampler.py:46-48
== FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=...) with use_dynamic_shifting=False, which the
reference configures at inference_flux.py:14) and the indexing of the denoising loop
(arcflux_pipeline.py:455-493): step i reads timesteps[sum(substeps[:i])], the last step integrates to 0.
"""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def retrieve_raw_timesteps(num_inference_steps: int, total_substeps: int, timestep_ratio: float):
    seg = 1.0 / (num_inference_steps - 1 + timestep_ratio)
    raw: List[float] = []
    counts: List[int] = []
    t = 1.0
    for i in range(num_inference_steps):
        size = seg if i < num_inference_steps - 1 else seg * timestep_ratio
        n = max(round(size * total_substeps), 1)
        counts.append(n)
        raw.extend(np.linspace(t, t - size, n, endpoint=False).clip(min=0.0).tolist())
        t -= size
    return raw, counts, sum(counts)


def warp_sigma(raw, shift: float = 3.2):
    raw = np.asarray(raw, dtype=np.float64)
    return shift * raw / (1.0 + (shift - 1.0) * raw)


def denoise_sigmas(nfe: int, total_substeps: int = 128, timestep_ratio: float = 1.0,
                   shift: float = 3.2) -> List[float]:
    """sigma at each of the `nfe` network calls followed by the final sigma (0): nfe + 1 fp32 values.

    The scheduler keeps sigmas/timesteps as fp32 tensors (timesteps = sigma * 1000 in fp32, then
    sigma_t_src = t / 1000, arcflux_pipeline.py:461-462), reproduced here.
    """
    raw, counts, total = retrieve_raw_timesteps(nfe, total_substeps, timestep_ratio)
    sig = torch.tensor(np.asarray(raw, dtype=np.float32))           # scheduler: np.float32 sigmas
    sig = shift * sig / (1 + (shift - 1) * sig)
    timesteps = sig * 1000.0
    out = []
    idx = 0
    for i in range(nfe):
        out.append(float((timesteps[idx] / 1000.0).item()))
        idx += counts[i]
    out.append(0.0)
    assert idx == total
    return out


def flux_time_inputs(sigma: float, guidance_scale: float) -> Tuple[float, float]:
    """What diffusers' Timesteps layer sees for FLUX: `timestep.to(bf16) * 1000` (arcflux.py:160-162).

    The pipeline passes t/1000 in fp32 (arcflux_pipeline.py:472); the transformer casts it to bf16 and
    multiplies by 1000 in bf16 — e.g. 0.761905 -> 0.76171875 -> 760.0; guidance 3.5 -> 3504.
    """
    def q(v: float) -> float:
        t = torch.tensor(v, dtype=torch.float32).to(torch.bfloat16) * 1000
        return float(t.float().item())
    return q(sigma), q(guidance_scale)
