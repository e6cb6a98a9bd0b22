import ctypes as C, os, sys
os.environ.setdefault("AFB_ATTN_DEBUG_MODE", "7")
import torch
sys.path.insert(0, ".")
from arcflow_b200 import ops, _lib
B, S, H = 8, 4608, 24
qkv = torch.randn(B, S, 3 * H * 128, device="cuda").bfloat16()
q, k, v = qkv[..., :H * 128], qkv[..., H * 128:2 * H * 128], qkv[..., 2 * H * 128:]
for _ in range(3):
    ops.attention(q, k, v)
torch.cuda.synchronize()
buf = (C.c_int64 * (2 * 64 * 8))()
_lib.check(_lib.load().afb_debug_attention_trace(buf, 2 * 64 * 8), "trace")
import numpy as np
a = np.array(buf[:]).reshape(2, 64, 8)
t0 = a[0, 0, 0]
print("stamps: 0 loop-top | 1 S ready | 2 max done | 3 turn granted | 4 exp done | 5 P published   (cycles, relative)")
for j in range(12, 18):
    for t in range(2):
        r = a[t, j, :6] - t0
        print(f"j={j} wg={t}  top {r[0]:8d}  wait_S {r[1]-r[0]:5d}  ld {a[t,j,6]-t0-r[1]:5d} max {r[2]-(a[t,j,6]-t0):5d}  wait_turn {r[3]-r[2]:5d}  exp {r[4]-r[3]:5d}  store {r[5]-r[4]:5d}   iter {a[t,j+1,0]-a[t,j,0]:5d}")
