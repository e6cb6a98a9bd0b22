"""Generates tests/golden/reference_sampler.npz by EXECUTING the reference's own code.

Run in the build container only (needs /root/reference). The GPU box never reads the reference; it
reads the committed .npz. What is executed, unmodified, from /root/reference:
  * lakonlab/models/diffusions/policies/arcflow.py      ArcFlowPolicy (imports standalone)
  * lakonlab/pipelines/arcflux_pipeline.py              retrieve_raw_timesteps, ArcFluxPipeline._unpack_mp,
                                                        _pack_latents, _unpack_latents, momentum_integration
The reference's third-party imports that are absent here (diffusers, mmcv, ...) are replaced by empty
stub modules — none of the functions above touches them; only the `class X(FluxPipeline, ...)`
statement needs the names to exist.
"""
from __future__ import annotations

import importlib
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "reference_sampler.npz"


class _Anything(type):
    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub


class _Stub(metaclass=_Anything):
    def __init__(self, *a, **k):
        pass


def _stub_module(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _ga(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Stub

    m.__getattr__ = _ga  # type: ignore[attr-defined]
    sys.modules[name] = m
    return m


def _pkg(name: str, path: Path):
    m = types.ModuleType(name)
    m.__path__ = [str(path)]
    sys.modules[name] = m
    return m


def load_reference():
    for name in ["diffusers", "diffusers.utils", "diffusers.image_processor", "diffusers.models",
                 "diffusers.pipelines", "diffusers.pipelines.flux", "diffusers.pipelines.flux.pipeline_flux",
                 "diffusers.schedulers"]:
        _stub_module(name)
    sys.modules["diffusers.utils"].is_torch_xla_available = lambda: False

    class FluxPipeline:  # the only base-class behaviour the exercised methods rely on
        pass

    sys.modules["diffusers.pipelines.flux.pipeline_flux"].FluxPipeline = FluxPipeline
    _pkg("lakonlab", REF / "lakonlab")
    _pkg("lakonlab.models", REF / "lakonlab/models")
    _pkg("lakonlab.models.diffusions", REF / "lakonlab/models/diffusions")
    _pkg("lakonlab.pipelines", REF / "lakonlab/pipelines")
    loader = _stub_module("lakonlab.pipelines.arcflow_loader")

    class ArcFlowLoaderMixin:
        pass

    loader.ArcFlowLoaderMixin = ArcFlowLoaderMixin
    policies = importlib.import_module("lakonlab.models.diffusions.policies")
    pipe_mod = importlib.import_module("lakonlab.pipelines.arcflux_pipeline")
    return policies, pipe_mod


def main():
    policies, pipe_mod = load_reference()
    ArcFlowPolicy = policies.ArcFlowPolicy
    Pipe = pipe_mod.ArcFluxPipeline
    out = {}

    # ---- schedules ------------------------------------------------------------------------------
    for nfe in (1, 2, 4, 8):
        for ratio in (1.0, 0.5):
            raw, sub, tot = pipe_mod.retrieve_raw_timesteps(nfe, 128, ratio)
            key = f"sched_nfe{nfe}_r{int(ratio * 10)}"
            out[key + "_raw"] = np.asarray(raw, dtype=np.float64)
            out[key + "_sub"] = np.asarray(sub, dtype=np.int64)
            out[key + "_tot"] = np.asarray([tot], dtype=np.int64)

    # ---- a fake pipeline object carrying exactly the attributes the methods read ------------------
    class Fake:
        pass

    fake = Fake()
    fake.vae_scale_factor = 8
    fake.transformer = Fake()
    fake.transformer.num_gaussians = 16
    fake.scheduler = Fake()
    fake.scheduler.config = Fake()
    fake.scheduler.config.num_train_timesteps = 1000
    fake.num_timesteps = 128

    g = torch.Generator().manual_seed(20251017)
    B, K, C, px = 2, 16, 16, 64   # 64 px image -> 8x8 latent -> 4x4 tokens
    h = w = px // 16
    S = h * w
    # network outputs are emitted in bf16 by the reference (transformer.dtype) and cast to fp32 at
    # arcflux_pipeline.py:487, so every input here is bf16-representable; the logits are log-softmaxed
    # over K in bf16 exactly as arcflux.py:246-247 does.
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    means_tok = bf(torch.randn(B, S, K, 64, generator=g))
    logits_tok = bf(torch.randn(B, S, K, 4, generator=g) * 2.0)
    logw_tok = logits_tok.to(torch.bfloat16).log_softmax(dim=-2).to(torch.float32)
    gam_tok = bf(torch.randn(B, S, K - 1, 4, generator=g))
    # exercise the |z| < eps and z == 0 branches of the integral term
    gam_tok[0, 0, 0, :] = 0.0
    gam_tok[0, 1, 1, :] = bf(torch.tensor(1e-5))
    gam_tok[0, 2, 2, :] = bf(torch.tensor(-1e-5))
    x_tok = torch.randn(B, S, 64, generator=g)
    out["in_means_tok"] = means_tok.numpy()
    out["in_logits_tok"] = logits_tok.numpy()
    out["in_logw_tok"] = logw_tok.numpy()
    out["in_gam_tok"] = gam_tok.numpy()
    out["in_x_tok"] = x_tok.numpy()

    mp = dict(means=means_tok.clone(), logweights=logw_tok.clone(), loggammas=gam_tok.clone())
    mp = Pipe._unpack_mp(fake, mp, px, px, 16, gm_patch_size=1)
    x_img = Pipe._unpack_latents(x_tok, px, px, 8, target_patch_size=1)
    out["unpacked_means"] = mp["means"].numpy()
    out["unpacked_logw"] = mp["logweights"].numpy()
    out["unpacked_gam"] = mp["loggammas"].numpy()
    out["unpacked_x"] = x_img.numpy()
    repacked = Pipe._pack_latents(x_img, B, 16, 2 * h, 2 * w, patch_size=1)
    out["repacked_x"] = repacked.numpy()

    cases = [(1.0, 1.0, 761.9047619), (0.7619047619, 0.7619047619, 0.0), (1.0, 0.9, 500.0),
             (0.9056603774, 0.85, 300.0)]
    for i, (s_src, s_start, t_end) in enumerate(cases):
        sigma_src = torch.tensor(s_src)
        policy = ArcFlowPolicy({k: v.clone() for k, v in mp.items()}, x_img, sigma_src)
        x_end, sig_end, _ = Pipe.momentum_integration(
            fake, sigma_src, x_img, torch.tensor(s_start), torch.tensor(t_end), policy, eps=1e-4)
        out[f"mi{i}_args"] = np.asarray([s_src, s_start, t_end / 1000.0], dtype=np.float64)
        out[f"mi{i}_x_end"] = x_end.numpy()
        out[f"mi{i}_x_end_tok"] = Pipe._pack_latents(x_end, B, 16, 2 * h, 2 * w, patch_size=1).numpy()
        vel = policy.velocity(sigma_src.reshape(1, 1, 1, 1), torch.tensor(s_start).reshape(1, 1, 1, 1))
        out[f"mi{i}_velocity"] = vel.numpy()
        out[f"mi{i}_means_x0"] = policy.denoising_output_x_0["means"].numpy()

    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.1f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
