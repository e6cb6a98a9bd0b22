import sys, time, torch, os
sys.path.insert(0, ".")
from arcflow_b200 import ops
dev = "cuda"
B, S, H = 8, 4608, 24
qkv = torch.randn(B, S, 3 * H * 128, device=dev).bfloat16()
q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
o = torch.empty(B, S, H * 128, device=dev, dtype=torch.bfloat16)
bound = float(os.environ.get("AFB_DIAG_BOUND", "0"))   # > 0: the fixed-reference (bounded-score) kernel
if bound > 0:   # make the bound true: unit-RMS rows per head, so |q.k| / sqrt(128) <= sqrt(128) = 11.3
    def unit(t):
        t = t.float().reshape(B, S, H, 128)
        return (t * torch.rsqrt(t.pow(2).mean(-1, keepdim=True))).reshape(B, S, H * 128).bfloat16()
    qkv = torch.cat([unit(q), unit(k), v], -1).contiguous()
    q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
for _ in range(3):
    ops.attention(q, k, v, out=o, score_bound=bound)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.attention(q, k, v, out=o, score_bound=bound)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("mode", os.environ.get("AFB_ATTN_DEBUG_MODE", "0"), "bound", bound, f"{ms:.3f} ms", f"{4*B*H*S*S*128/ms/1e9:.0f} TFLOP/s", flush=True)
