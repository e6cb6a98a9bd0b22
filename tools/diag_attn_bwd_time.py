"""Times afb_attention (forward) and afb_attention_backward (delta + dQ kernel + dK/dV kernel) at the training shape:
B 4, H 24, S 4608. AFB_ATTN_BWD_LEGACY=1 selects the un-overlapped dK/dV kernel (read once per process: A/B = two runs)."""
import json, os, sys, torch
sys.path.insert(0, ".")
from arcflow_b200 import ops
B, S, H = 4, 4608, 24
dev = "cuda"
qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
d_o = torch.randn(B, S, H * 128, device=dev).bfloat16()
lse = torch.empty(B, H, S, device=dev, dtype=torch.float32)
o = ops.attention(q, k, v, lse=lse)


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


fwd = timed(lambda: ops.attention(q, k, v, lse=lse))
bwd = timed(lambda: ops.attention_backward(q, k, v, o, d_o, lse))
unit = 2.0 * B * H * S * S * 128          # one S x S x 128 product
print(json.dumps(dict(what="attention forward / backward, B 4, H 24, S 4608, isolated loop of 10",
                      legacy_dkdv=os.environ.get("AFB_ATTN_BWD_LEGACY", "0"), forward_ms=fwd, backward_ms=bwd,
                      ratio=bwd / fwd, forward_tflops=2 * unit / fwd / 1e9,
                      backward_tflops_7_products=7 * unit / bwd / 1e9,
                      backward_tflops_5_products=5 * unit / bwd / 1e9)), flush=True)
