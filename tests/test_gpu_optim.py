"""GPU: fused grad-clip + AdamW + Karras-EMA arena step vs torch.optim.AdamW / the reference's formulas on CPU."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture()
def opt(lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from arcflow_b200.optim import FlatAdamW
    shapes = {"proj_out_means.weight": (8, 33), "proj_out_loggamma.weight": (5, 7), "proj_out_loggamma.bias": (5,),
              "blocks.0.lora_A.weight": (16, 10)}
    return FlatAdamW(shapes, "cuda"), shapes


def test_adamw_matches_torch_and_applies_lr_mult_clip_warmup(opt):
    from arcflow_b200.optim import warmup_lr
    o, shapes = opt
    g = torch.Generator().manual_seed(0)
    init = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    o.load_params(init)
    ref_p = {n: t.clone().requires_grad_(True) for n, t in init.items()}
    groups = [dict(params=[p], lr=1e-4 * (0.1 if "loggamma" in n else 1.0)) for n, p in ref_p.items()]
    base_lrs = [gr["lr"] for gr in groups]
    ref = torch.optim.AdamW(groups, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    for it in (98, 99, 100, 101, 250):            # crosses warm-up end (100) and clip begin (100)
        grads = {n: torch.randn(s, generator=g) * (30.0 if it == 101 else 1.0) for n, s in shapes.items()}
        for n in shapes:
            o.grad(n).copy_(grads[n])
            ref_p[n].grad = grads[n].clone()
        if it >= 100:
            torch.nn.utils.clip_grad_norm_(list(ref_p.values()), 50.0)
        for gr, b in zip(groups, base_lrs):
            gr["lr"] = warmup_lr(b, it)
        ref.step()
        info = o.step(it)
        total = math.sqrt(sum((grads[n] ** 2).sum().item() for n in shapes))
        assert info["diffusion_grad_norm"] == pytest.approx(total, rel=1e-5) and not info["skipped"]
        for n in shapes:
            assert torch.allclose(o.param(n).cpu(), ref_p[n].detach(), rtol=2e-5, atol=1e-7), (it, n)
    assert torch.equal(o.shadow.cpu(), o.params.cpu().bfloat16())


def test_nan_gradient_skips_the_step_but_not_the_ema(opt):
    o, shapes = opt
    o.load_params({n: torch.ones(s) for n, s in shapes.items()})
    before = o.params.clone()
    o.grads.fill_(1.0)
    o.grad("proj_out_means.weight")[0, 0] = float("nan")
    info = o.step(500)
    assert info["skipped"] and math.isnan(info["diffusion_grad_norm"])
    assert torch.equal(o.params, before) and o.steps_taken == 0


def test_grad_clip_skip_ratio_skips_large_norms(lib):
    """`<k>_grad_clip_skip_ratio`: a norm above ratio x grad_clip skips the step like a non-finite one
    (reference lakonlab/models/base.py:81,91-95)."""
    from arcflow_b200.optim import FlatAdamW
    shapes = {"proj_out_means.weight": (8, 33), "proj_out_loggamma.bias": (5,)}
    o = FlatAdamW(shapes, "cuda", max_norm=1.0, clip_begin_iter=0, clip_skip_ratio=4.0)
    o.load_params({n: torch.ones(s) for n, s in shapes.items()})
    before = o.params.clone()
    o.grads.fill_(1.0)                               # norm = sqrt(272) = 16.5 > 4 x 1
    info = o.step(10)
    assert info["skipped"] and math.isnan(info["diffusion_grad_norm"]) and torch.equal(o.params, before)
    o.grads.fill_(0.2)                               # norm = 3.3: clipped to 1, not skipped
    info = o.step(11)
    assert not info["skipped"] and info["diffusion_grad_norm"] == pytest.approx(0.2 * math.sqrt(o.n), rel=1e-5)
    assert not torch.equal(o.params, before) and o.steps_taken == 1


def test_two_optimizers_on_two_streams_do_not_share_norm_scratch(lib):
    from arcflow_b200.optim import FlatAdamW
    a = FlatAdamW({"w.lora_A.weight": (512, 1024)}, "cuda", max_norm=0.0)
    b = FlatAdamW({"w.lora_A.weight": (640, 1024)}, "cuda", max_norm=0.0)
    assert a.norm_scratch.data_ptr() != b.norm_scratch.data_ptr()
    a.grads.fill_(1.0), b.grads.fill_(2.0)
    torch.cuda.synchronize()
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(20):
        with torch.cuda.stream(sa):
            ia = a.step(0)
        with torch.cuda.stream(sb):
            ib = b.step(0)
        assert ia["diffusion_grad_norm"] == pytest.approx(math.sqrt(a.n), rel=1e-6)
        assert ib["diffusion_grad_norm"] == pytest.approx(2.0 * math.sqrt(b.n), rel=1e-6)


def test_karras_ema(opt):
    from arcflow_b200.optim import karras_momentum
    o, shapes = opt
    o.load_params({n: torch.zeros(s) for n, s in shapes.items()})
    ema_ref = torch.zeros(o.n)
    for it in (50, 100, 101, 150):
        o.grads.fill_(1.0)
        o.step(it)
        p = o.params.cpu()
        if it < 100:
            ema_ref = p.clone()                      # straight copy before start_iter
        else:
            m = karras_momentum(it, 100, 7.0)
            ema_ref = p + (ema_ref - p) * m          # lerp(net, ema, m)
        assert torch.allclose(o.ema.cpu(), ema_ref, rtol=1e-5, atol=1e-8)
    assert karras_momentum(100) == 0.0 and karras_momentum(101) == pytest.approx(0.5 ** 8)
