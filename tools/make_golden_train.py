"""Generates tests/golden/reference_train_rollout.npz by EXECUTING the reference's own training roll-out:

  lakonlab/models/diffusions/arcflow.py     ArcFlowImitationBase.piid_segment_momentum, policy_average_u_momentum,
                                            momentum_integration (train variant)
  lakonlab/models/diffusions/policies/      ArcFlowPolicy (detach / dropout_ / velocity)
  lakonlab/models/diffusions/sampler.py     ContinuousTimeStepSampler.warp_t

Run in the build container only. Absent third-party imports (mmcv, mmgen) are stubbed; what that removes from the
executed path: `GaussianFlow` (base class; none of its methods is reached) and the loss module — `flow_loss` is
replaced by the restatement of DiffusionMSELoss + mmgen reduction (15 * mean_samples mean_chw
(pred - tgt)^2), which therefore stays UNPINNED. The teacher is a synthetic closed-form velocity field.
Random draws are reproduced by seeding torch and drawing in the reference's order (dropout uniforms [B,K,1,1,1],
student uniforms [B,n], teacher uniforms [B,n-1]); they are stored as inputs.
"""
from __future__ import annotations

import contextlib
import importlib
import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import REF, _pkg, _stub_module  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "reference_train_rollout.npz"


def load_reference():
    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    _stub_module("mmcv")
    for name in ["mmgen", "mmgen.models", "mmgen.models.architectures", "mmgen.models.architectures.common",
                 "mmgen.models.builder"]:
        _stub_module(name)
    sys.modules["mmgen.models.builder"].MODULES = _Registry()
    sys.modules["mmgen.models.architectures.common"].get_module_device = lambda m: torch.device("cpu")
    _pkg("lakonlab", REF / "lakonlab")
    _pkg("lakonlab.models", REF / "lakonlab/models")
    diff = _pkg("lakonlab.models.diffusions", REF / "lakonlab/models/diffusions")

    class GaussianFlow:   # base class placeholder: no method of it is reached by the exercised functions
        pass

    diff.GaussianFlow = GaussianFlow
    utils = _stub_module("lakonlab.utils")
    utils.module_eval = lambda m: contextlib.nullcontext()
    arc = importlib.import_module("lakonlab.models.diffusions.arcflow")
    sampler = importlib.import_module("lakonlab.models.diffusions.sampler")
    policies = importlib.import_module("lakonlab.models.diffusions.policies")
    return arc, sampler, policies


def main():
    arc, sampler, policies = load_reference()
    out = {}
    B, K, C, H, W = 3, 16, 16, 8, 8
    bf = lambda t: t.to(torch.bfloat16).to(torch.float32)
    g = torch.Generator().manual_seed(777)
    means = bf(torch.randn(B, K, C, H, W, generator=g))
    logw = torch.randn(B, K, 1, H, W, generator=g).mul(2).to(torch.bfloat16).log_softmax(dim=1).to(torch.float32)
    gam = bf(torch.randn(B, K - 1, 1, H, W, generator=g))
    x_src = torch.randn(B, C, H, W, generator=g)
    out.update(in_means=means.numpy(), in_logw=logw.numpy(), in_gam=gam.numpy(), in_x=x_src.numpy())

    class Teacher:
        """closed-form velocity field standing in for the frozen FLUX teacher"""
        def __call__(self, return_u=True, x_t=None, t=None, **kw):
            return torch.tanh(x_t * 0.7) * (0.5 + t.reshape(-1, 1, 1, 1)) - 0.3 * x_t.flip(1)

    def flow_loss(d):   # restatement of DiffusionMSELoss + mmgen 'flatmean' / constant rescale 30 / mean (A.9)
        per = ((d["u_t_pred"] - d["u_t"]) ** 2).flatten(1).mean(1) * 0.5
        return (per * 30.0).mean()

    cases = [  # (raw_t_src, segment_size, teacher_ratio, gm_dropout)
        (1.0, 0.5, 1.0, 0.1), (1.0, 0.5, 0.35, 0.1), (0.5, 0.5, 0.0, 0.3), (1.0, 0.25, 0.6, 0.0),
    ]
    for ci, (raw0, seg, ratio, p) in enumerate(cases):
        obj = arc.ArcFlowImitationBase.__new__(arc.ArcFlowImitationBase)
        obj.train_cfg = dict(eps=1e-4, total_substeps=128, num_intermediate_states=4, window_substeps=3, gm_dropout=p)
        obj.timestep_sampler = sampler.ContinuousTimeStepSampler(num_timesteps=1, shift=3.2, logit_normal_enable=False)
        obj.num_timesteps = 1
        obj.flow_loss = flow_loss
        raw_t_src = torch.full((B,), raw0)
        sigma_t_src = obj.timestep_sampler.warp_t(raw_t_src).reshape(B, 1, 1, 1)
        m = means.clone().requires_grad_(True)
        lw = logw.clone().requires_grad_(True)
        gm = gam.clone().requires_grad_(True)
        policy = policies.ArcFlowPolicy(dict(means=m, logweights=lw, loggammas=gm), x_src, sigma_t_src)
        seed = 1000 + ci
        # pre-draw in the reference's order with the same seed
        torch.manual_seed(seed)
        drop_u = torch.rand((B, K, 1, 1, 1)) if 0 < p < 1 else torch.ones((B, K, 1, 1, 1))
        student_u = torch.rand((B, 4))
        teacher_u = torch.rand((B, 3))
        torch.manual_seed(seed)
        loss, x_dst, raw_dst = obj.piid_segment_momentum(Teacher(), policy, x_src, raw_t_src, sigma_t_src, ratio, seg,
                                                         dict(), get_x_t_dst=True)
        loss.backward()
        pre = f"c{ci}_"
        out[pre + "args"] = np.asarray([raw0, seg, ratio, p], dtype=np.float64)
        out[pre + "drop_u"] = drop_u.reshape(B, K).numpy()
        out[pre + "student_u"] = student_u.numpy()
        out[pre + "teacher_u"] = teacher_u.numpy()
        out[pre + "loss"] = np.asarray([float(loss)], dtype=np.float64)
        out[pre + "x_dst"] = x_dst.numpy()
        out[pre + "raw_dst"] = raw_dst.numpy()
        out[pre + "grad_means"] = m.grad.numpy()
        out[pre + "grad_logw"] = lw.grad.numpy()
        out[pre + "grad_gam"] = gm.grad.numpy()
    OUT.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1024:.1f} KiB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
