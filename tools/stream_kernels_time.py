"""HBM-rate check of the streaming kernels of the train step at their real sizes (B 4, S 4608): dropout_rows (M-wide and
D-wide, plain and accumulate), gelu_bwd. Prints ms and achieved GB/s (algorithmic bytes: every operand once)."""
import json, sys, torch
sys.path.insert(0, ".")
from arcflow_b200 import ops
dev = "cuda"
B, S, D, M = 4, 4608, 3072, 12288
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


out = {}
for name, cols in (("M", M), ("D", D)):
    x = torch.randn(B, S, cols, device=dev).bfloat16()
    y = torch.zeros_like(x)
    nbytes = x.numel() * 2
    ms = timed(lambda: ops.dropout_rows(x, 7, 3, 0.05, out=y))
    out[f"dropout_rows_{name}"] = dict(ms=ms, gbs=2 * nbytes / ms / 1e6)
    ms = timed(lambda: ops.dropout_rows(x, 7, 3, 0.05, out=y, accumulate=True))
    out[f"dropout_rows_{name}_accumulate"] = dict(ms=ms, gbs=3 * nbytes / ms / 1e6)
x = torch.randn(B * S, M, device=dev).bfloat16()
d = torch.randn(B * S, M, device=dev).bfloat16()
ms = timed(lambda: ops.gelu_bwd(d, x))
out["gelu_bwd_M"] = dict(ms=ms, gbs=3 * x.numel() * 2 / ms / 1e6)
print(json.dumps(out))
