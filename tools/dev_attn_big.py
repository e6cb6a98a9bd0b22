import sys, math, torch
sys.path.insert(0, ".")
from arcflow_b200 import ops
DEV="cuda"
for hi, S in [(30.0,1024),(60.0,1000),(200.0,777)]:
    g = torch.Generator(device=DEV).manual_seed(4)
    q = torch.randn(1, S, 128, device=DEV, generator=g)
    k = torch.randn(1, S, 128, device=DEV, generator=g)
    k = k * torch.linspace(0.5, hi, S, device=DEV)[None, :, None]
    v = torch.randn(1, S, 128, device=DEV, generator=g)
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    lse = torch.empty(1, 1, S, device=DEV, dtype=torch.float32)
    o = ops.attention(q, k, v, lse=lse)
    qh, kh, vh = [t.reshape(1, S, 1, 128).float().transpose(1, 2) for t in (q, k, v)]
    ref = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh).transpose(1, 2).reshape(1, S, 128)
    sc = (qh.double() @ kh.double().transpose(-1, -2)) / math.sqrt(128)
    lse_ref = torch.logsumexp(sc, -1) * 1.4426950408889634
    err = (o.float() - ref)
    rowerr = err.norm(dim=-1)[0] / ref.norm(dim=-1)[0]
    print(hi, S, "finite", torch.isfinite(o.float()).all().item(), "rel", (err.norm() / ref.norm()).item(),
          "lse err", (lse.double() - lse_ref).abs().max().item(), "worst rows", rowerr.topk(5).indices.tolist(), rowerr.topk(5).values.tolist(), flush=True)
