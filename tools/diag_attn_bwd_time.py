"""Times afb_attention_backward (delta + dQ kernel + dK/dV kernel) at the training shape: B 4, H 24, S 4608."""
import sys, torch
sys.path.insert(0, ".")
from arcflow_b200 import ops
B, S, H = 4, 4608, 24
dev = "cuda"
qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
d_o = torch.randn(B, S, H * 128, device=dev).bfloat16()
lse = torch.empty(B, H, S, device=dev, dtype=torch.float32)
o = ops.attention(q, k, v, lse=lse)
for _ in range(2):
    ops.attention_backward(q, k, v, o, d_o, lse)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention_backward(q, k, v, o, d_o, lse)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"attention backward {ms:.3f} ms  {10 * B * H * S * S * 128 / ms / 1e9:.0f} TFLOP/s (5 matmuls of algorithmic work)", flush=True)
