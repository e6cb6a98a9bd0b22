"""Developer smoke for the individual kernels against torch on the same GPU (not the parity suite —
that lives in tests/ and checks against oracle/). Usage: python tools/dev_check_ops.py [gemm|attn|elem]"""
import math
import sys
import time

import torch

sys.path.insert(0, ".")
from arcflow_b200 import ops, _lib  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def rel(a, b):
    a = a.float(); b = b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item(), (a - b).abs().max().item()


def bench(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def check_gemm():
    for (B, R, N, Ks, epi) in [
        (1, 128, 256, [64], 0),
        (1, 256, 512, [128], 0),
        (1, 1000, 1152, [3072], 0),
        (2, 384, 3072, [3072, 256], 1),
        (2, 200, 3072, [3072, 12288, 256], 2),
        (1, 4096, 9216, [3072], 0),
    ]:
        a = [torch.randn(B, R, k, device=dev).mul(0.5).bfloat16() for k in Ks]
        w = torch.randn(N, sum(Ks), device=dev).mul(0.05).bfloat16()
        bias = torch.randn(N, device=dev).bfloat16()
        out = torch.zeros(B, R, N, device=dev, dtype=torch.bfloat16)
        gate = torch.randn(B, N, device=dev).bfloat16() if epi == 2 else None
        res = torch.randn(B, R, N, device=dev).bfloat16() if epi == 2 else None
        ops.gemm(a, w, out, bias=bias, epilogue=epi, gate=gate, res=res)
        torch.cuda.synchronize()
        ref = torch.cat(a, -1).float() @ w.float().t() + bias.float()
        if epi == 1:
            ref = torch.nn.functional.gelu(ref, approximate="tanh")
        if epi == 2:
            ref = res.float() + gate.float()[:, None, :] * ref
        r, mx = rel(out, ref)
        print(f"gemm B{B} R{R} N{N} K{Ks} epi{epi}: rel {r:.3e} max {mx:.3e}", flush=True)
    # perf
    M, N, K = 36864, 12288, 3072
    a = torch.randn(1, M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).mul(0.02).bfloat16()
    out = torch.empty(1, M, N, device=dev, dtype=torch.bfloat16)
    ms = bench(lambda: ops.gemm(a, w, out))
    print(f"gemm {M}x{N}x{K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
    ms = bench(lambda: torch.matmul(a[0], w.t()))
    print(f"torch.matmul same: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


def check_attn():
    for (B, S, H) in [(1, 256, 1), (1, 512, 2), (2, 768, 2), (1, 300, 1), (1, 1000, 3), (1, 4608, 4)]:
        qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
        q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
        o = ops.attention(q, k, v)
        torch.cuda.synchronize()
        qh, kh, vh = [t.reshape(B, S, H, 128).transpose(1, 2).float() for t in (q, k, v)]
        ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(B, S, H * 128)
        r, mx = rel(o, ref)
        print(f"attn B{B} S{S} H{H}: rel {r:.3e} max {mx:.3e}", flush=True)
    B, S, H = 8, 4608, 24
    qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
    q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
    o = torch.empty(B, S, H * 128, device=dev, dtype=torch.bfloat16)
    ms = bench(lambda: ops.attention(q, k, v, out=o))
    fl = 4 * B * H * S * S * 128
    print(f"attn B{B} S{S} H{H}: {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    qh, kh, vh = [t.reshape(B, S, H, 128).transpose(1, 2) for t in (q, k, v)]
    ms = bench(lambda: torch.nn.functional.scaled_dot_product_attention(qh, kh, vh))
    print(f"torch sdpa same: {ms:.3f} ms {fl/ms/1e9:.1f} TFLOP/s", flush=True)


def check_elem():
    B, R, D = 2, 300, 3072
    x = torch.randn(B, R, D, device=dev).bfloat16()
    mod = torch.randn(B, 2 * D, device=dev).mul(0.3).bfloat16()
    y = ops.ln_modulate(x, mod[:, :D], mod[:, D:])
    ref = torch.nn.functional.layer_norm(x.float(), (D,), eps=1e-6) * (1 + mod[:, None, :D].float()) + mod[:, None, D:].float()
    print("ln_modulate", rel(y, ref), flush=True)
    # rmsnorm_rope
    H, S, St = 3, 200, 40
    qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
    orig = qkv.clone()
    ws = [torch.randn(128, device=dev).mul(0.1).add(1).bfloat16() for _ in range(4)]
    ang = torch.rand(S, 64, device=dev) * 6.28
    cos = torch.cos(ang).repeat_interleave(2, 1).contiguous(); sin = torch.sin(ang).repeat_interleave(2, 1).contiguous()
    ops.rmsnorm_rope(qkv, 0, H * 128, H, St, ws[0], ws[1], cos, sin, wq_txt=ws[2], wk_txt=ws[3])
    def ref_rr(t, w_img, w_txt):
        t = t.float().reshape(B, S, H, 128)
        t = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6)
        w = torch.where((torch.arange(S, device=dev) < St)[None, :, None, None], w_txt.float(), w_img.float())
        t = t.bfloat16().float() * w
        t = t.bfloat16().float()
        tr = torch.stack([-t[..., 1::2], t[..., 0::2]], -1).flatten(3)
        return (t * cos[None, :, None, :] + tr * sin[None, :, None, :]).reshape(B, S, H * 128)
    print("rmsnorm_rope q", rel(qkv[..., :H * 128], ref_rr(orig[..., :H * 128], ws[0], ws[2])), flush=True)
    print("rmsnorm_rope k", rel(qkv[..., H * 128:2 * H * 128], ref_rr(orig[..., H * 128:2 * H * 128], ws[1], ws[3])), flush=True)
    print("rmsnorm_rope v untouched", torch.equal(qkv[..., 2 * H * 128:], orig[..., 2 * H * 128:]), flush=True)
    # small linear
    for m, n, k in [(8, 1000, 3072), (1, 3072, 256), (3, 77, 768)]:
        xs = torch.randn(m, k, device=dev).bfloat16(); w = torch.randn(n, k, device=dev).mul(0.05).bfloat16(); bb = torch.randn(n, device=dev).bfloat16()
        y = ops.small_linear(xs, w, bb, silu_in=True)
        ref = torch.nn.functional.silu(xs.float()).bfloat16().float() @ w.float().t() + bb.float()
        print("small_linear", (m, n, k), rel(y, ref), flush=True)
        y2 = ops.small_linear(xs, w, None, out=y.clone(), accumulate=True)
        ref2 = y.float() + xs.float() @ w.float().t()
        print("small_linear acc", rel(y2, ref2), flush=True)
    t = torch.tensor([1000.0, 760.0, 3504.0], device=dev)
    e = ops.timestep_embed(t)
    f = torch.exp(-math.log(10000) * torch.arange(128, device=dev) / 128)
    ref = torch.cat([torch.cos(t[:, None] * f), torch.sin(t[:, None] * f)], -1)
    print("timestep_embed", rel(e, ref), flush=True)
    # sampler
    T, K = 777, 16
    head = torch.randn(T, 1152, device=dev).bfloat16()
    xin = torch.randn(T, 64, device=dev)
    ssrc, sst, send = 1.0, 0.9, 0.7619
    xo = ops.sampler_step(head, xin, ssrc, sst, send)
    means = head[:, :1024].float().reshape(T, K, 16, 4)
    lw = torch.log_softmax(head[:, 1024:1088].float().reshape(T, K, 1, 4), 1).bfloat16().float()
    lg = head[:, 1088:1148].float().reshape(T, K - 1, 1, 4)
    w = torch.softmax(lw, 1)
    dtp, dts = ssrc - sst, sst - send
    decay = torch.cat([torch.ones(T, 1, 1, 4, device=dev), torch.exp(lg * dtp)], 1)
    z = lg * dts
    sg = torch.sign(z); sg[sg == 0] = 1
    zs = sg * z.abs().clamp(min=1e-4)
    phi = torch.cat([torch.ones(T, 1, 1, 4, device=dev), torch.expm1(zs) / zs], 1)
    disp = (w * means * decay * dts * phi).sum(1)
    ref = xin - disp.reshape(T, 64)
    print("sampler_step", rel(xo, ref), flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    t0 = time.time()
    if which in ("gemm", "all"):
        check_gemm()
    if which in ("attn", "all"):
        check_attn()
    if which in ("elem", "all"):
        check_elem()
    print(f"done in {time.time()-t0:.1f}s launches={_lib.load().afb_launch_count()}")
