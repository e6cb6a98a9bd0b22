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
