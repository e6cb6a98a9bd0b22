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


# ---- block-wise 8-bit moment state (bitsandbytes AdamW8bit restated: oracle/adamw8bit_oracle.py) ------------------------

SHAPES8 = {"blocks.0.lora_A.weight": (16, 300), "blocks.0.lora_B.weight": (300, 16), "proj_out_means.weight": (64, 100),
           "proj_out_means.bias": (64,), "proj_out_loggamma.weight": (60, 100), "proj_out_loggamma.bias": (60,),
           "norm_out.linear.bias": (200,)}


def _oracle_state(o):
    import numpy as np
    return dict(p=o.params.cpu().numpy().copy(), c1=o.state1.cpu().numpy().copy(), c2=o.state2.cpu().numpy().copy(),
                a1=o.absmax1.cpu().numpy().copy(), a2=o.absmax2.cpu().numpy().copy(),
                m=o.exp_avg.cpu().numpy().copy(), v=o.exp_avg_sq.cpu().numpy().copy())


def test_adamw8bit_kernel_matches_the_blockwise_oracle_over_steps(lib):
    """Five steps (warm-up, clip, lr multiplier range, 8-bit and fp32-state tensors) against the numpy restatement run on
    the same gradients: codes equal except where fp32 rounding lands within an ulp of a code mid-point (then off by one),
    absmax and parameters equal to fp32 rounding."""
    import numpy as np
    from oracle import adamw8bit_oracle as O
    from arcflow_b200.optim import FlatAdamW, warmup_lr
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    o = FlatAdamW(SHAPES8, "cuda", state_bits=8)
    g = torch.Generator().manual_seed(0)
    o.load_params({n: torch.randn(s, generator=g) for n, s in SHAPES8.items()})
    q1, q2 = o.qmap1.cpu().numpy(), o.qmap2.cpu().numpy()
    st = _oracle_state(o)
    lo_mask = np.zeros(o.n, bool)
    lo_mask[o.lo[0]:o.lo[1]] = True
    n8 = o.n8
    mismatched = 0
    for k, it in enumerate((98, 99, 100, 101, 250)):
        scale = 30.0 if it == 101 else 1.0
        for n, s in SHAPES8.items():
            o.grad(n).copy_(torch.randn(s, generator=g) * scale * (1e-3 if "lora" in n else 1.0))
        grads = o.grads.cpu().numpy().copy()
        info = o.step(it)
        assert not info["skipped"]
        norm = np.float32(np.sqrt(np.float32((grads.astype(np.float64) ** 2).sum())))
        clip = np.float32(min(1.0, 50.0 / (norm + np.float32(1e-6)))) if it >= 100 else np.float32(1.0)
        gc = (grads * clip).astype(np.float32)
        lr = np.where(lo_mask, np.float32(warmup_lr(1e-4, it) * 0.1), np.float32(warmup_lr(1e-4, it))).astype(np.float32)
        st["p"][:n8], st["c1"], st["c2"], st["a1"], st["a2"] = O.adamw8bit_step(
            st["p"][:n8], gc[:n8], st["c1"], st["c2"], st["a1"], st["a2"], q1, q2, k + 1, lr[:n8])
        st["p"][n8:], st["m"], st["v"] = O.adamw32_step(st["p"][n8:], gc[n8:], st["m"], st["v"], k + 1, lr[n8:])
        got = _oracle_state(o)
        d1 = np.abs(got["c1"].astype(int) - st["c1"].astype(int))
        d2 = np.abs(got["c2"].astype(int) - st["c2"].astype(int))
        assert d1.max() <= 1 and d2.max() <= 1, (it, d1.max(), d2.max())
        mismatched += int((d1 > 0).sum() + (d2 > 0).sum())
        np.testing.assert_allclose(got["a1"], st["a1"], rtol=2e-6, atol=0, err_msg=str(it))
        np.testing.assert_allclose(got["a2"], st["a2"], rtol=2e-6, atol=0, err_msg=str(it))
        np.testing.assert_allclose(got["p"], st["p"], rtol=2e-5, atol=2e-7, err_msg=str(it))
        # continue from the kernel's own state so one-code differences do not compound in the comparison
        st = got
    assert mismatched <= 5e-4 * 2 * n8 * 5, mismatched
    assert torch.equal(o.shadow.cpu(), o.params.cpu().bfloat16())
    assert o.steps_taken == 5


def test_adamw8bit_tracks_fp32_adamw_and_resumes_bit_exactly(lib):
    """60 steps on noisy gradients (a pulled-towards-target signal + unit noise, the regime the block-wise code book is made
    for): the kernel's trajectory equals the oracle's run on the CPU, stays within a few % of fp32 AdamW's, and a checkpoint
    taken mid-way resumes onto the same bytes."""
    import numpy as np
    from oracle import adamw8bit_oracle as O
    from arcflow_b200.optim import FlatAdamW
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    shapes = {"a.lora_A.weight": (64, 512), "a.lora_B.weight": (512, 64), "proj_out_loggamma.weight": (60, 128)}
    kw = dict(lr=1e-3, warmup_iters=0, max_norm=0.0)
    a = FlatAdamW(shapes, "cuda", state_bits=8, **kw)
    b = FlatAdamW(shapes, "cuda", state_bits=32, **kw)
    g = torch.Generator().manual_seed(3)
    init = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    target = {n: torch.randn(s, generator=g) for n, s in shapes.items()}
    noise = [{n: torch.randn(s, generator=g) for n, s in shapes.items()} for _ in range(60)]
    a.load_params(init)
    b.load_params(init)
    assert a.state1.dtype == torch.uint8 and a.state1.numel() == a.n8 == a.n    # every tensor here is 8-bit
    # the oracle's state for the same arena
    q1, q2 = a.qmap1.cpu().numpy(), a.qmap2.cpu().numpy()
    op = a.params.cpu().numpy().copy()
    oc1, oc2 = np.zeros(a.n, np.uint8), np.zeros(a.n, np.uint8)
    oa1, oa2 = np.zeros(a.n // 256, np.float32), np.zeros(a.n // 256, np.float32)
    lr = np.full(a.n, 1e-3, np.float32)
    lr[a.lo[0]:a.lo[1]] = 1e-4
    tflat = torch.zeros(a.n)
    for n in shapes:
        a.view(tflat, n).copy_(target[n])
    for k in range(60):
        it = 200 + k
        zflat = torch.zeros(a.n)
        for n in shapes:
            a.view(zflat, n).copy_(noise[k][n])
        for o in (a, b):
            for n in shapes:
                o.grad(n).copy_(o.param(n) - target[n].cuda() + noise[k][n].cuda())
            o.step(it)
        og = (op - tflat.numpy() + zflat.numpy()).astype(np.float32)
        op, oc1, oc2, oa1, oa2 = O.adamw8bit_step(op, og, oc1, oc2, oa1, oa2, q1, q2, k + 1, lr)
        if it == 229:
            saved = a.state_dict()
        if it == 239:
            at_240 = {k_: getattr(a, k_).clone() for k_ in ("params", "state1", "state2", "absmax1", "absmax2", "ema")}
    for n in shapes:
        travel = (b.param(n) - init[n].cuda()).norm()
        vs_oracle = (a.param(n).cpu() - a.view(torch.from_numpy(op), n)).norm() / travel.cpu()
        vs_fp32 = (a.param(n) - b.param(n)).norm() / travel
        assert vs_oracle < 5e-3, (n, float(vs_oracle))            # same algorithm: one-code flips at mid-points only
        assert vs_fp32 < 0.04, (n, float(vs_fp32))                # oracle on the same seeds: 1.6e-2 (8-bit moment noise)
    # resume from the snapshot taken after iteration 229: ten more steps land on the same bytes
    c = FlatAdamW(shapes, "cuda", state_bits=8, **kw)
    c.load_state_dict(saved)
    for k in range(30, 40):
        for n in shapes:
            c.grad(n).copy_(c.param(n) - target[n].cuda() + noise[k][n].cuda())
        c.step(200 + k)
    for k_, v in at_240.items():
        assert torch.equal(getattr(c, k_), v), k_


def test_adamw8bit_skip_leaves_codes_and_params_untouched(lib):
    from arcflow_b200.optim import FlatAdamW
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    o = FlatAdamW(SHAPES8, "cuda", state_bits=8)
    o.load_params({n: torch.ones(s) for n, s in SHAPES8.items()})
    o.grads.fill_(0.5)
    o.step(500)
    before = {k: getattr(o, k).clone() for k in ("params", "state1", "state2", "absmax1", "absmax2", "exp_avg")}
    o.grads.fill_(1.0)
    o.grad("blocks.0.lora_A.weight")[0, 0] = float("inf")
    info = o.step(501)
    assert info["skipped"] and o.steps_taken == 1
    for k, v in before.items():
        assert torch.equal(getattr(o, k), v), k
