"""GPU: activation-gradient kernels of the adapter-only backward vs torch autograd (fp32 on the same bf16 inputs).
Tolerance: gradients are emitted in bf16 -> rel-L2 <= 6e-3 per op (attention: 1e-2, P and dS are bf16 MMA operands)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import arcflow_oracle as O  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


@pytest.fixture(scope="module")
def ops(lib):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from arcflow_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("B,S,H", [(1, 128, 1), (1, 256, 2), (2, 384, 2), (1, 300, 1), (1, 1000, 2), (1, 77, 1)])
def test_attention_backward_parity(ops, B, S, H):
    g = torch.Generator(device=DEV).manual_seed(S)
    qkv = torch.randn(B, S, 3 * H * 128, device=DEV, generator=g).bfloat16()
    q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
    d_o = torch.randn(B, S, H * 128, device=DEV, generator=g).bfloat16()
    lse = torch.empty(B, H, S, device=DEV, dtype=torch.float32)
    o = ops.attention(q, k, v, lse=lse)
    dq, dk, dv = ops.attention_backward(q, k, v, o, d_o, lse)
    qf, kf, vf = [t.detach().float().reshape(B, S, H, 128).requires_grad_(True) for t in (q, k, v)]
    ref = O._attention(qf, kf, vf)
    # lse check (log2 domain of the scaled scores)
    sc = torch.einsum("bshd,bthd->bhst", qf, kf) / 128 ** 0.5
    assert torch.allclose(lse, torch.logsumexp(sc, -1) * 1.4426950408889634, atol=2e-3)
    ref.backward(d_o.float())
    for name, got, r in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        assert torch.isfinite(got.float()).all()
        assert rel(got, r.reshape(B, S, H * 128)) < 1e-2, (name, rel(got, r.reshape(B, S, H * 128)))


def test_ln_modulate_bwd_parity(ops):
    g = torch.Generator().manual_seed(1)
    B, R, D = 2, 50, 3072
    x = torch.randn(B, R, D, generator=g).mul(2).add(0.5).bfloat16()
    dy = torch.randn(B, R, D, generator=g).bfloat16()
    sc = torch.randn(B, D, generator=g).mul(0.3).bfloat16()
    xf = x.float().requires_grad_(True)
    (O._ln(xf) * (1 + sc.float()[:, None])).backward(dy.float())
    got = ops.ln_modulate_bwd(x.to(DEV), dy.to(DEV), sc.to(DEV))
    assert rel(got, xf.grad) < 6e-3
    base = torch.randn(B, R, D, generator=g).bfloat16()
    got2 = ops.ln_modulate_bwd(x.to(DEV), dy.to(DEV), sc.to(DEV), dh=base.clone().to(DEV))
    assert rel(got2, xf.grad + base.float()) < 6e-3


def test_rowscale_and_gelu_bwd_parity(ops):
    g = torch.Generator().manual_seed(2)
    B, R, C = 2, 40, 1024
    x = torch.randn(B, R, C, generator=g).bfloat16()
    vec = torch.randn(B, C, generator=g).bfloat16()
    assert rel(ops.rowscale(x.to(DEV), vec.to(DEV)), x.float() * vec.float()[:, None]) < 6e-3
    buf = torch.randn(B * R, C + 256, generator=g).bfloat16().to(DEV)       # strided view inside a wider buffer
    pre = torch.randn(B * R, C, generator=g).mul(2).bfloat16()
    pf = pre.float().requires_grad_(True)
    d0 = buf[:, 128:128 + C].clone()
    torch.nn.functional.gelu(pf, approximate="tanh").backward(d0.float().cpu())
    ops.gelu_bwd(buf[:, 128:128 + C], pre.to(DEV))
    assert rel(buf[:, 128:128 + C], pf.grad) < 6e-3


def test_rmsnorm_rope_bwd_parity(ops):
    g = torch.Generator().manual_seed(3)
    B, H, St, gh, gw = 2, 2, 10, 4, 5
    S = St + gh * gw
    raw = torch.randn(B, S, 3 * H * 128, generator=g).bfloat16()
    dout = torch.randn(B, S, 3 * H * 128, generator=g).bfloat16()
    ws = [torch.randn(128, generator=g).mul(0.1).add(1).bfloat16() for _ in range(4)]
    cos, sin = O.flux_rope(St, gh, gw)
    cos, sin = cos.bfloat16().float(), sin.bfloat16().float()
    rf = raw.float().requires_grad_(True)

    def fwd(t, w_img, w_txt):
        t = t.reshape(B, S, H, 128)
        y = t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True) + 1e-6)
        w = torch.where((torch.arange(S) < St)[None, :, None, None], w_txt.float(), w_img.float())
        return O.apply_rope(y * w, cos, sin).reshape(B, S, H * 128)

    out = torch.cat([fwd(rf[..., :H * 128], ws[0], ws[2]), fwd(rf[..., H * 128:2 * H * 128], ws[1], ws[3]),
                     rf[..., 2 * H * 128:]], -1)
    out.backward(dout.float())
    got = ops.rmsnorm_rope_bwd(dout.clone().to(DEV), raw.to(DEV), 0, H * 128, H, St, ws[0].to(DEV), ws[1].to(DEV),
                               cos.to(DEV), sin.to(DEV), wq_txt=ws[2].to(DEV), wk_txt=ws[3].to(DEV))
    assert rel(got[..., :2 * H * 128], rf.grad[..., :2 * H * 128]) < 6e-3
    assert torch.equal(got[..., 2 * H * 128:].cpu(), dout[..., 2 * H * 128:])      # v columns untouched


@pytest.mark.parametrize("B,R,K,N,K2", [(1, 256, 256, 256, 0), (2, 300, 512, 384, 0), (1, 700, 1024, 3072, 256),
                                        (2, 130, 3072, 264, 0), (1, 512, 64, 1024, 64)])
def test_gemm_transposed_weight(ops, B, R, K, N, K2):
    """dX = dY W (+ dT A) with W in the forward's [out, in] layout, no transposed copy; bf16 result of an fp32 sum."""
    g = torch.Generator(device=DEV).manual_seed(K + N)
    dy = torch.randn(B, R, K, device=DEV, generator=g).bfloat16()
    wbuf = torch.randn(K, N + 64, device=DEV, generator=g).mul(0.05).bfloat16()
    w = wbuf[:, 32:32 + N] if N % 8 == 0 and False else wbuf[:, :N]          # strided rows (ld = N + 64)
    res = torch.randn(B, R, N, device=DEV, generator=g).bfloat16()
    ref = dy.float() @ w.float()
    segs, w2 = [dy], None
    if K2:
        dt = torch.randn(B, R, K2, device=DEV, generator=g).bfloat16()
        w2 = torch.randn(K2, N, device=DEV, generator=g).mul(0.05).bfloat16()
        ref = ref + dt.float() @ w2.float()
        segs.append(dt)
    out = torch.empty(B, R, N, device=DEV, dtype=torch.bfloat16)
    ops.gemm(segs, w, out, transposed=True, w2=w2)
    assert rel(out, ref) < 4e-3
    from arcflow_b200 import _lib
    ops.gemm(segs, w, out, transposed=True, w2=w2, epilogue=_lib.AFB_EPI_BIAS_RES, res=res)
    assert rel(out, ref + res.float()) < 4e-3


def test_lora_dropout_mask_matches_oracle(ops):
    """The counter-based keep mask of afb_dropout_rows is the oracle's `lora_dropout_mask`, including column slices of one
    logical tensor (FLUX single-block proj_out input = [attn | mlp]) and the accumulate form used by the backward."""
    g = torch.Generator().manual_seed(5)
    B, R, C0, C1 = 2, 37, 128, 256
    seed, layer, p = 0x1234_5678_9ABC_DEF, 11, 0.3
    a = torch.randn(B, R, C0, generator=g).bfloat16()
    b = torch.randn(B, R, C1, generator=g).bfloat16()
    keep = O.lora_dropout_mask(seed, layer, (B, R, C0 + C1), p)
    assert 0.25 < 1 - keep.float().mean().item() < 0.35
    out = torch.zeros(B, R, C0 + C1, dtype=torch.bfloat16, device=DEV)
    ops.dropout_rows(a.to(DEV), seed, layer, p, out=out[..., :C0], logical_cols=C0 + C1, col0=0)
    ops.dropout_rows(b.to(DEV), seed, layer, p, out=out[..., C0:], logical_cols=C0 + C1, col0=C0)
    ref = (torch.cat([a, b], -1).float() * keep / (1 - p)).bfloat16()
    assert torch.equal(out.cpu(), ref)
    base = torch.randn(B, R, C0, generator=g).bfloat16()
    acc = ops.dropout_rows(a.to(DEV), seed, layer, p, out=base.clone().to(DEV), logical_cols=C0 + C1, col0=0, accumulate=True)
    assert torch.equal(acc.cpu(), (base.float() + a.float() * keep[..., :C0] / (1 - p)).bfloat16())
    assert not torch.equal(O.lora_dropout_mask(seed, layer + 1, (B, R, C0 + C1), p), keep)
