"""Developer check: afb_attention vs torch SDPA (fp32 math on the same bf16 inputs) on ragged shapes, plus timing."""
import os, sys, torch
sys.path.insert(0, ".")
from arcflow_b200 import ops
torch.manual_seed(0)
dev = "cuda"
worst = 0.0
for (B, S, H) in [(1, 64, 1), (1, 65, 2), (2, 200, 2), (1, 777, 3), (2, 1280, 4), (1, 4608, 2)]:
    qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
    q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
    lse = torch.empty(B, H, S, device=dev, dtype=torch.float32)
    o = ops.attention(q, k, v, lse=lse)
    f = lambda t: t.float().view(B, S, H, 128).transpose(1, 2)
    ref = torch.nn.functional.scaled_dot_product_attention(f(q), f(k), f(v)).transpose(1, 2).reshape(B, S, H * 128)
    sc = (f(q) @ f(k).transpose(-1, -2)) / 128 ** 0.5
    lse_ref = torch.logsumexp(sc, -1) * 1.4426950408889634
    rel = ((o.float() - ref).norm() / ref.norm()).item()
    lerr = (lse - lse_ref).abs().max().item()
    worst = max(worst, rel)
    print(f"B{B} S{S} H{H}: rel-L2 {rel:.2e}  lse max-abs {lerr:.2e}", flush=True)
assert worst < 4e-3, worst
B, S, H = 8, 4608, 24
qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
o = torch.empty(B, S, H * 128, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(q, k, v, out=o)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention(q, k, v, out=o)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("impl", os.environ.get("AFB_ATTN_IMPL", "default"), f"{ms:.3f} ms", f"{4*B*H*S*S*128/ms/1e9:.0f} TFLOP/s", flush=True)
