"""ArcFlowPolicy in packed-token layout, backed by the native policy kernel.

Mirror of lakonlab/models/diffusions/policies/arcflow.py:9-114 (velocity / detach / dropout). The policy wraps the RAW
head tensor of one student call ([B*tokens, head_ld] bf16: means | logits | loggamma) instead of three unpacked
image-major tensors; the reference's unused `means_x_0 = x_t - sigma * means` pass (:45) is not computed.
"""
from __future__ import annotations

from typing import Optional

import torch

from arcflow_b200 import _lib, ops


class ArcFlowPolicy:
    def __init__(self, head: torch.Tensor, x_t_src: torch.Tensor, sigma_t_src, num_gaussians: int = 16, eps: float = 1e-4,
                 drop_mask: Optional[torch.Tensor] = None):
        self.head = head.reshape(-1, head.shape[-1])
        self.x_t_src = x_t_src
        self.batch = x_t_src.shape[0]
        self.sigma_t_src = sigma_t_src
        self.num_gaussians = num_gaussians
        self.eps = eps
        self.drop_mask = drop_mask

    def velocity(self, sigma_t_src, sigma_t) -> torch.Tensor:
        """u = sum_k softmax(w)_k mu_k exp(lambda_k (sigma_src - sigma_t))  (reference :52-76)."""
        return ops.policy_eval(self.head, _lib.AFB_POLICY_VELOCITY, sigma_t_src, sigma_t, batch=self.batch,
                               drop_mask=self.drop_mask, num_gaussians=self.num_gaussians, eps=self.eps)

    def integrate(self, x_t_start, sigma_t_start, sigma_t_end) -> torch.Tensor:
        """momentum_integration from sigma_t_start to sigma_t_end (arcflux_pipeline.py:195-249 / arcflow.py:28-79)."""
        return ops.policy_eval(self.head, _lib.AFB_POLICY_INTEGRATE, self.sigma_t_src, sigma_t_start, sigma_t_end,
                               x=x_t_start, batch=self.batch, drop_mask=self.drop_mask,
                               num_gaussians=self.num_gaussians, eps=self.eps)

    def copy(self):
        return ArcFlowPolicy(self.head, self.x_t_src, self.sigma_t_src, self.num_gaussians, self.eps, self.drop_mask)

    def detach(self):
        return self.copy()          # the head tensor carries no autograd graph on this path

    def dropout_(self, p: float, uniforms: Optional[torch.Tensor] = None, generator=None):
        """mask = rand < p, never dropping every component of a sample (reference :96-106)."""
        if p <= 0 or p >= 1:
            return self
        u = uniforms if uniforms is not None else torch.rand((self.batch, self.num_gaussians), generator=generator)
        m = u.cpu() < p
        self.drop_mask = m & ~m.all(dim=1, keepdim=True)
        return self

    def dropout(self, p: float, **kw):
        return self.copy().dropout_(p, **kw)
