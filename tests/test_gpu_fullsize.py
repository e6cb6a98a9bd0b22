"""Full-size checks (BASELINE.json configs[1]: ArcFlow-FLUX, 19 + 38 blocks, D = 3072, 1024 x 1024, batch 8, 2 NFE).

The CPU oracle cannot run the whole network at this size in test time, so parity at full size is carried by
  * a FULL-WIDTH, depth-reduced forward (1 double + 1 single block, D = 3072, 24 heads, S = 4608) against the oracle —
    the real tile shapes of every kernel (CTA-pair GEMM, multi-wave attention, LoRA K-extension at r = 256), and
  * size-independent properties of the full-depth loop: run-to-run determinism, independence of an image from its
    batch-mates (images are the sharded unit), NFE consistency of the schedule, and the closed form of the
    sampler step (lambda = 0 and equal weights reduce it to an Euler step with the mixture mean).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import arcflow_oracle as O  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def test_full_width_forward_parity(lib, parity):
    from arcflow_b200.config import ArcFluxConfig
    from arcflow_b200.model import ArcFluxEngineModel
    from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict
    cfg = ArcFluxConfig(num_layers=1, num_single_layers=1)
    sd = make_flux_state_dict(cfg, seed=1234, device="cpu")
    x, txt, pooled = make_flux_inputs(cfg, 1, 1024, 1024, txt_len=512, seed=42, device="cpu")
    model = ArcFluxEngineModel(sd, cfg, device=DEV)
    sigma, gs, grid = 0.7619047761, 3.5, (64, 64)
    ours = model.split_heads(model.forward_heads(x.to(DEV), txt.to(DEV), pooled.to(DEV), sigma, gs, grid))
    args = (x.bfloat16(), txt, pooled, torch.full([1], sigma), torch.full([1], gs), grid)
    ref = O.flux_forward(sd, cfg, *args, dtype=torch.float32)
    ref_bf16 = O.flux_forward(sd, cfg, *args, dtype=torch.bfloat16)
    for key in ("means", "logweights", "loggammas"):
        e_ours, e_bf16 = rel(ours[key], ref[key]), rel(ref_bf16[key], ref[key])
        parity(f"flux_fullwidth_fwd.{key}", e_ours, max(2e-2, 1.5 * e_bf16))


@pytest.fixture(scope="module")
def full_model(lib):
    from arcflow_b200.config import flux_dev
    from arcflow_b200.model import ArcFluxEngineModel
    from arcflow_b200.synthetic import make_flux_inputs, make_flux_state_dict
    cfg = flux_dev()
    sd = make_flux_state_dict(cfg, seed=1234, device=DEV)
    model = ArcFluxEngineModel(sd, cfg, device=DEV, consume_state_dict=True)
    del sd
    x, txt, pooled = make_flux_inputs(cfg, 8, 1024, 1024, txt_len=512, seed=42, device=DEV)
    yield model, x, txt, pooled
    del model
    torch.cuda.empty_cache()


def test_full_size_denoise_properties(full_model):
    model, x, txt, pooled = full_model
    grid = (64, 64)
    a = model.denoise(x.clone(), txt, pooled, grid, num_inference_steps=2, timestep_ratio=1.0)
    b = model.denoise(x.clone(), txt, pooled, grid, num_inference_steps=2, timestep_ratio=1.0)
    assert a.shape == (8, 4096, 64) and a.dtype == torch.float32 and torch.isfinite(a).all()
    assert torch.equal(a, b), "the 2-NFE loop must be run-to-run deterministic"
    # an image does not depend on its batch-mates, whatever the batch size the kernels tile over
    sub = [5, 2]
    c = model.denoise(x[sub].clone(), txt[sub], pooled[sub], grid, num_inference_steps=2, timestep_ratio=1.0)
    assert torch.equal(c, a[sub])
    # one network evaluation moves the latents by the whole first segment: the 1-NFE result differs from 2-NFE,
    # and both stay at the scale of a unit-variance latent
    d = model.denoise(x.clone(), txt, pooled, grid, num_inference_steps=1, timestep_ratio=1.0)
    assert torch.isfinite(d).all() and not torch.equal(d, a)
    assert 0.05 < a.std().item() < 20 and 0.05 < d.std().item() < 20


def test_full_size_sampler_step_closed_form(lib):
    """lambda_k = 0 (phi = 1) and equal log-weights: x_end = x - (s_start - s_end) * mean_k(mu_k), any sizes."""
    from arcflow_b200 import ops
    g = torch.Generator().manual_seed(3)
    n_tok, K, C = 8 * 4096, 16, 64
    means = torch.randn(n_tok, K, C, generator=g).bfloat16()
    head = torch.zeros(n_tok, 1152, dtype=torch.bfloat16)
    head[:, :K * C] = means.reshape(n_tok, K * C)
    x = torch.randn(8, 4096, C, generator=g)
    out = ops.sampler_step(head.to(DEV), x.to(DEV), 1.0, 1.0, 0.7619047761, num_gaussians=K)
    want = x - (1.0 - 0.7619047761) * means.float().mean(1).reshape(8, 4096, C)
    assert (out.cpu() - want).abs().max().item() < 2e-5
