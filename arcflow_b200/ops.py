"""Tensor-level wrappers over the C ABI (torch is used only for device memory and the stream).

These are the reference-facing operators of the hot path; each takes CUDA tensors, validates them,
and enqueues exactly one kernel of libarcflow_b200.so on the current stream. A CPU tensor or a missing
library is an error — the oracle under `oracle/` is test infrastructure and is never reached from here.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import AfbError, AttnDesc, GemmDesc

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise AfbError(f"{name}: expected a CUDA tensor (no CPU fallback exists)")
    if t.dtype != dtype:
        raise AfbError(f"{name}: expected dtype {dtype}, got {t.dtype}")


def _rows3(t: torch.Tensor, name: str):
    """View a [batches, rows, cols] (or [rows, cols]) tensor as (ptr, ld, batch_stride, batches, rows, cols)."""
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.stride(2) != 1:
        raise AfbError(f"{name}: need [batches, rows, cols] with contiguous last dim, got {tuple(t.shape)} "
                       f"strides {t.stride()}")
    return t.data_ptr(), t.stride(1), t.stride(0), t.shape[0], t.shape[1], t.shape[2]


def gemm(a: Sequence[torch.Tensor] | torch.Tensor, w: torch.Tensor, out: torch.Tensor,
         bias: Optional[torch.Tensor] = None, epilogue: int = _lib.AFB_EPI_BIAS,
         gate: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None, transposed: bool = False,
         w2: Optional[torch.Tensor] = None, alpha: float = 1.0, qk_norm_rope: Optional[dict] = None) -> torch.Tensor:
    """out[b, r, :] = epi(sum_s a_s[b, r, :] @ w[:, koff_s : koff_s + K_s].T).

    `a` is one tensor or up to three K-segments sharing [batches, rows]; `w` is [N, sum K_s] (torch
    Linear layout); `gate` is [batches, N]; `res` has the shape of `out` and may alias it.
    transposed=True: `w` is [K, N] (dX = dY W with the forward's [out, in] weight); `w2` [K2, N] continues the K rows.
    """
    lib = _lib.load()
    segs = [a] if isinstance(a, torch.Tensor) else list(a)
    if not 1 <= len(segs) <= 3:
        raise AfbError("gemm: 1..3 A segments")
    d = GemmDesc()
    batches = rows = None
    for i, s in enumerate(segs):
        _chk(s, BF16, f"gemm a[{i}]")
        ptr, ld, bs, nb, nr, k = _rows3(s, f"gemm a[{i}]")
        if batches is None:
            batches, rows = nb, nr
        elif (nb, nr) != (batches, rows):
            raise AfbError("gemm: A segments disagree on [batches, rows]")
        d.a[i], d.a_ld[i], d.a_batch_stride[i], d.a_k[i] = ptr, ld, bs, k
    _chk(w, BF16, "gemm w")
    _chk(out, BF16, "gemm out")
    if w.dim() != 2 or w.stride(1) != 1:
        raise AfbError("gemm: w must be 2-D with a contiguous last dim")
    optr, old, obs, ob, orows, on = _rows3(out, "gemm out")
    n_w = w.shape[1] if transposed else w.shape[0]
    k_w = (w.shape[0] + (w2.shape[0] if w2 is not None else 0)) if transposed else w.shape[1]
    if (ob, orows) != (batches, rows) or on != n_w:
        raise AfbError(f"gemm: out shape {tuple(out.shape)} does not match [{batches}, {rows}, {n_w}]")
    if k_w != sum(s.shape[-1] for s in segs):
        raise AfbError("gemm: w K does not match the A segments")
    d.batches, d.rows_per_batch = batches, rows
    d.w, d.w_ld, d.n = w.data_ptr(), w.stride(0), n_w
    if transposed:
        d.w_transposed, d.w_k = 1, w.shape[0]
        if w2 is not None:
            _chk(w2, BF16, "gemm w2")
            if w2.dim() != 2 or w2.stride(1) != 1 or w2.shape[1] != n_w:
                raise AfbError("gemm: w2 must be [K2, N]")
            d.w2, d.w2_ld = w2.data_ptr(), w2.stride(0)
    elif w2 is not None:
        raise AfbError("gemm: w2 needs transposed=True")
    d.epilogue = epilogue
    if qk_norm_rope is not None:
        # fused QKV projection epilogue: dict(norm_q=bf16 [128], norm_k=bf16 [128], rope=rope_pack(cos, sin),
        # qk_cols=leading q + k columns, row0=table position of output row 0 of each batch)
        q = qk_norm_rope
        _chk(q["norm_q"], BF16, "gemm norm_q")
        _chk(q["norm_k"], BF16, "gemm norm_k")
        _chk(q["rope"], torch.float32, "gemm rope")
        if q["rope"].dim() != 5 or tuple(q["rope"].shape[1:]) != (2, 16, 32, 4) or not q["rope"].is_contiguous():
            raise AfbError("gemm: rope must be the contiguous fp32 table rope_pack() returns")
        if q["rope"].shape[0] * 32 < int(q.get("row0", 0)) + rows:
            raise AfbError("gemm: rope table has fewer positions than row0 + rows")
        d.epilogue = _lib.AFB_EPI_BIAS_QKNORM_ROPE
        d.norm_q, d.norm_k, d.rope = q["norm_q"].data_ptr(), q["norm_k"].data_ptr(), q["rope"].data_ptr()
        d.rope_row0, d.qk_cols, d.norm_eps = int(q.get("row0", 0)), int(q["qk_cols"]), float(q.get("eps", 1e-6))
    d.alpha = float(alpha)      # scales the accumulator before bias / epilogue (0 in the struct means 1)
    d.out, d.out_ld, d.out_batch_stride = optr, old, obs
    if bias is not None:
        _chk(bias, BF16, "gemm bias")
        d.bias = bias.data_ptr()
    if gate is not None:
        _chk(gate, BF16, "gemm gate")
        if gate.dim() != 2 or gate.stride(1) != 1 or gate.shape[0] != batches:
            raise AfbError("gemm: gate must be [batches, N]")
        d.gate, d.gate_batch_stride = gate.data_ptr(), gate.stride(0)
    if res is not None:
        _chk(res, BF16, "gemm res")
        rptr, rld, rbs, rb, rr, rn = _rows3(res, "gemm res")
        if (rb, rr, rn) != (batches, rows, on):
            raise AfbError("gemm: res shape mismatch")
        d.res, d.res_ld, d.res_batch_stride = rptr, rld, rbs
    _lib.check(lib.afb_gemm(C.byref(d), _stream()), "afb_gemm")
    return out


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: Optional[torch.Tensor] = None,
              scale: float = 0.0, lse: Optional[torch.Tensor] = None, score_bound: float = 0.0) -> torch.Tensor:
    """Joint non-causal attention. q/k/v: [batch, seq, heads*128] views (last dim contiguous).
    lse (optional, fp32 [batch, heads, seq]) receives the log2-domain logsumexp needed by attention_backward.
    score_bound > 0: a guaranteed bound on |scale * q.k| (see qk_score_bound) -> the fixed-reference softmax kernel."""
    lib = _lib.load()
    d = AttnDesc()
    shp = None
    for name, t in (("q", q), ("k", k), ("v", v)):
        _chk(t, BF16, f"attention {name}")
        ptr, ld, bs, nb, ns, nc = _rows3(t, f"attention {name}")
        if shp is None:
            shp = (nb, ns, nc)
        elif shp != (nb, ns, nc):
            raise AfbError("attention: q/k/v shapes differ")
        setattr(d, name, ptr)
        setattr(d, f"{name}_ld", ld)
        setattr(d, f"{name}_batch_stride", bs)
    nb, ns, nc = shp
    if nc % 128:
        raise AfbError("attention: last dim must be heads*128")
    if out is None:
        out = torch.empty((nb, ns, nc), dtype=BF16, device=q.device)
    _chk(out, BF16, "attention out")
    optr, old, obs, ob, os_, oc = _rows3(out, "attention out")
    if (ob, os_, oc) != shp:
        raise AfbError("attention: out shape mismatch")
    d.o, d.o_ld, d.o_batch_stride = optr, old, obs
    d.batch, d.seq, d.heads = nb, ns, nc // 128
    d.scale = scale
    d.score_bound = float(score_bound)
    if lse is not None:
        _chk(lse, torch.float32, "attention lse")
        if tuple(lse.shape) != (nb, nc // 128, ns) or not lse.is_contiguous():
            raise AfbError("attention: lse must be contiguous fp32 [batch, heads, seq]")
        d.lse = lse.data_ptr()
    _lib.check(lib.afb_attention(C.byref(d), _stream()), "afb_attention")
    return out


def rope_pack(cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """fp32 [rows, 128] cos / sin tables (adjacent-pair layout) -> the warp-coalesced table the fused QK-norm + RoPE GEMM
    epilogue reads: fp32 [ceil(rows / 32), 2, 16, 32, 4] (see afb_rope_pack)."""
    lib = _lib.load()
    _chk(cos, torch.float32, "rope_pack cos")
    _chk(sin, torch.float32, "rope_pack sin")
    if cos.dim() != 2 or cos.shape[1] != 128 or cos.shape != sin.shape or not cos.is_contiguous() or not sin.is_contiguous():
        raise AfbError("rope_pack: cos / sin must be contiguous fp32 [rows, 128]")
    out = torch.empty(((cos.shape[0] + 31) // 32, 2, 16, 32, 4), dtype=torch.float32, device=cos.device)
    _lib.check(lib.afb_rope_pack(cos.data_ptr(), sin.data_ptr(), out.data_ptr(), cos.shape[0], _stream()), "afb_rope_pack")
    return out


def qk_score_bound(*norm_weight_pairs, head_dim: int = 128, margin: float = 1.05) -> float:
    """Bound on |q . k| / sqrt(head_dim) for per-head RMS-normalised, rotated q and k: RMSNorm makes ||q_hat||^2 <= head_dim,
    the elementwise weight scales it by at most max|w|, RoPE is a rotation of pairs — so
    |q . k| <= ||q|| ||k|| <= head_dim * max|w_q| * max|w_k|. Arguments: (w_q, w_k) tensors per stream of the joint sequence
    (the bound must hold for every query against every key, so the maxima are taken over all streams). `margin` covers the
    bf16 roundings after the norm, the weight multiply and the rotation (each < 0.4 %)."""
    wq = max(float(p[0].float().abs().max()) for p in norm_weight_pairs)
    wk = max(float(p[1].float().abs().max()) for p in norm_weight_pairs)
    return margin * head_dim * wq * wk / (head_dim ** 0.5)


def attention_backward(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, d_o: torch.Tensor,
                       lse: torch.Tensor, scale: float = 0.0):
    """dq, dk, dv (views of one fused [batch, seq, 3*heads*128] gradient buffer) for afb_attention.
    q/k/v must share leading dim and batch stride (views of one fused buffer); o and d_o likewise."""
    lib = _lib.load()
    for name, t in (("q", q), ("k", k), ("v", v), ("o", o), ("d_o", d_o)):
        _chk(t, BF16, f"attention_backward {name}")
    qp, qld, qbs, nb, ns, nc = _rows3(q, "attention_backward q")
    for t in (k, v):
        _, ld, bs, b2, s2, c2 = _rows3(t, "attention_backward k/v")
        if (ld, bs, b2, s2, c2) != (qld, qbs, nb, ns, nc):
            raise AfbError("attention_backward: q/k/v must be views of one fused buffer")
    op, old, obs, ob, os_, oc = _rows3(o, "attention_backward o")
    dp, dld, dbs, db_, ds_, dc = _rows3(d_o, "attention_backward d_o")
    if (old, obs) != (dld, dbs) or (ob, os_, oc) != (nb, ns, nc) or (db_, ds_, dc) != (nb, ns, nc):
        raise AfbError("attention_backward: o and d_o must share shape and strides")
    _chk(lse, torch.float32, "attention_backward lse")
    heads = nc // 128
    dqkv = torch.empty((nb, ns, 3 * nc), dtype=BF16, device=q.device)
    delta = torch.empty((nb, heads, ns), dtype=torch.float32, device=q.device)
    d = _lib.AttnBwdDesc()
    d.q, d.k, d.v, d.qkv_ld, d.qkv_batch_stride = qp, k.data_ptr(), v.data_ptr(), qld, qbs
    d.o, d.d_o, d.o_ld, d.o_batch_stride = op, dp, old, obs
    d.lse, d.delta_ws = lse.data_ptr(), delta.data_ptr()
    d.dq, d.dk, d.dv = dqkv.data_ptr(), dqkv[..., nc:].data_ptr(), dqkv[..., 2 * nc:].data_ptr()
    d.dqkv_ld, d.dqkv_batch_stride = 3 * nc, ns * 3 * nc
    d.batch, d.seq, d.heads, d.scale = nb, ns, heads, scale
    _lib.check(lib.afb_attention_backward(C.byref(d), _stream()), "afb_attention_backward")
    return dqkv[..., :nc], dqkv[..., nc:2 * nc], dqkv[..., 2 * nc:]


def ln_modulate(x: torch.Tensor, scale: torch.Tensor, shift: torch.Tensor,
                out: Optional[torch.Tensor] = None, eps: float = 1e-6) -> torch.Tensor:
    """LayerNorm(x) * (1 + scale[b]) + shift[b]; x: [batches, rows, dim] (rows contiguous), scale/shift: [batches, dim] views."""
    lib = _lib.load()
    _chk(x, BF16, "ln_modulate x")
    xptr, xld, xbs, nb, nr, dim = _rows3(x, "ln_modulate x")
    if xld != dim:
        raise AfbError("ln_modulate: x rows must be contiguous")
    if out is None:
        out = torch.empty_like(x)
    _chk(out, BF16, "ln_modulate out")
    optr, old, obs, ob, orr, od = _rows3(out, "ln_modulate out")
    if (ob, orr, od) != (nb, nr, dim) or old != dim:
        raise AfbError("ln_modulate: out shape mismatch")
    for name, t in (("scale", scale), ("shift", shift)):
        _chk(t, BF16, f"ln_modulate {name}")
        if t.dim() != 2 or t.shape != (nb, dim) or t.stride(1) != 1:
            raise AfbError(f"ln_modulate: {name} must be [batches, dim]")
    if scale.stride(0) != shift.stride(0):
        raise AfbError("ln_modulate: scale/shift must share a batch stride")
    _lib.check(lib.afb_ln_modulate(xptr, xbs, optr, obs, scale.data_ptr(), shift.data_ptr(),
                                   scale.stride(0), nb, nr, dim, eps, _stream()), "afb_ln_modulate")
    return out


def rmsnorm_rope(qkv: torch.Tensor, q_off: int, k_off: int, heads: int, txt_rows: int,
                 wq_img: torch.Tensor, wk_img: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
                 wq_txt: Optional[torch.Tensor] = None, wk_txt: Optional[torch.Tensor] = None,
                 eps: float = 1e-6) -> torch.Tensor:
    """In-place per-head RMSNorm + RoPE on the q and k thirds of a fused [batch, seq, ld] buffer."""
    lib = _lib.load()
    _chk(qkv, BF16, "rmsnorm_rope qkv")
    ptr, ld, bs, nb, ns, _ = _rows3(qkv, "rmsnorm_rope qkv")
    for name, t in (("cos", cos), ("sin", sin)):
        _chk(t, torch.float32, f"rmsnorm_rope {name}")
        if tuple(t.shape) != (ns, 128) or not t.is_contiguous():
            raise AfbError(f"rmsnorm_rope: {name} must be contiguous [seq, 128] fp32")
    for t in (wq_img, wk_img, wq_txt, wk_txt):
        if t is not None:
            _chk(t, BF16, "rmsnorm_rope weight")
    _lib.check(lib.afb_rmsnorm_rope(
        ptr, ld, bs, q_off, k_off, nb, ns, heads, txt_rows,
        wq_txt.data_ptr() if wq_txt is not None else None,
        wk_txt.data_ptr() if wk_txt is not None else None,
        wq_img.data_ptr(), wk_img.data_ptr(), cos.data_ptr(), sin.data_ptr(), eps, _stream()),
        "afb_rmsnorm_rope")
    return qkv


def small_linear(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, silu_in: bool = False,
                 accumulate: bool = False) -> torch.Tensor:
    """y[m, n] (+)= act(x) @ w.T + bias for m <= 8 rows (HBM-bound weight streaming)."""
    lib = _lib.load()
    _chk(x, BF16, "small_linear x")
    _chk(w, BF16, "small_linear w")
    m, k = x.shape
    n = w.shape[0]
    if out is None:
        if accumulate:
            raise AfbError("small_linear: accumulate needs `out`")
        out = torch.empty((m, n), dtype=BF16, device=x.device)
    _chk(out, BF16, "small_linear out")
    if bias is not None:
        _chk(bias, BF16, "small_linear bias")
    flags = (_lib.AFB_SL_SILU_IN if silu_in else 0) | (_lib.AFB_SL_ACCUMULATE if accumulate else 0)
    _lib.check(lib.afb_small_linear(x.data_ptr(), x.stride(0), w.data_ptr(), w.stride(0),
                                    bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                    out.stride(0), m, n, k, flags, _stream()), "afb_small_linear")
    return out


def timestep_embed(t: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _chk(t, torch.float32, "timestep_embed t")
    out = torch.empty((t.numel(), 256), dtype=BF16, device=t.device)
    _lib.check(lib.afb_timestep_embed(t.data_ptr(), out.data_ptr(), t.numel(), _stream()),
               "afb_timestep_embed")
    return out


def sampler_step(head: torch.Tensor, x: torch.Tensor, sigma_src: float, sigma_start: float,
                 sigma_end: float, num_gaussians: int = 16, eps: float = 1e-4,
                 want_bf16: bool = False):
    """One analytic momentum-integration step in packed token layout (see arcflow_b200.h)."""
    lib = _lib.load()
    _chk(head, BF16, "sampler_step head")
    _chk(x, torch.float32, "sampler_step x")
    if head.dim() != 2 or head.stride(1) != 1:
        raise AfbError("sampler_step: head must be [tokens, head_ld]")
    x2 = x.reshape(-1, 64)
    if not x2.is_contiguous() or x2.shape[0] != head.shape[0]:
        raise AfbError("sampler_step: x must be contiguous [tokens, 64]")
    out = torch.empty_like(x2)
    out_bf = torch.empty(x2.shape, dtype=BF16, device=x.device) if want_bf16 else None
    _lib.check(lib.afb_sampler_step(head.data_ptr(), head.stride(0), x2.data_ptr(), out.data_ptr(),
                                    out_bf.data_ptr() if want_bf16 else None, x2.shape[0],
                                    num_gaussians, sigma_src, sigma_start, sigma_end, eps, _stream()),
               "afb_sampler_step")
    out = out.reshape(x.shape)
    return (out, out_bf.reshape(x.shape)) if want_bf16 else out


# ------------------------------------------------------------------------------------------------
# training-side (trajectory distillation roll-out) operators
# ------------------------------------------------------------------------------------------------
def _host_floats(v, n: int, name: str):
    vals = [float(x) for x in (v.tolist() if isinstance(v, torch.Tensor) else v)]
    if len(vals) == 1 and n > 1:
        vals = vals * n
    if len(vals) != n:
        raise AfbError(f"{name}: expected {n} per-sample values, got {len(vals)}")
    return (C.c_float * n)(*vals)


def policy_eval(head: torch.Tensor, mode: int, sigma_src, sigma_start, sigma_end=None, x: Optional[torch.Tensor] = None,
                batch: Optional[int] = None, drop_mask=None, small=None, num_gaussians: int = 16, eps: float = 1e-4,
                want_bf16: bool = False):
    """Per-sample-time policy evaluation on the raw head tensor [batch*tokens, head_ld] (see arcflow_b200.h:
    AFB_POLICY_INTEGRATE / VELOCITY / AVERAGE_U). sigma_* are host sequences/tensors of length batch."""
    lib = _lib.load()
    _chk(head, BF16, "policy_eval head")
    if head.dim() != 2 or head.stride(1) != 1:
        raise AfbError("policy_eval: head must be [batch*tokens, head_ld]")
    if batch is None:
        batch = len(sigma_src)
    if head.shape[0] % batch:
        raise AfbError("policy_eval: rows of head not divisible by batch")
    tokens = head.shape[0] // batch
    a = _lib.PolicyArgs()
    a.head, a.head_ld, a.batch, a.tokens = head.data_ptr(), head.stride(0), batch, tokens
    a.num_gaussians, a.mode, a.eps = num_gaussians, mode, eps
    keep = [_host_floats(sigma_src, batch, "sigma_src"), _host_floats(sigma_start, batch, "sigma_start")]
    a.sigma_src, a.sigma_start = keep[0], keep[1]
    if sigma_end is not None:
        keep.append(_host_floats(sigma_end, batch, "sigma_end"))
        a.sigma_end = keep[-1]
    if drop_mask is not None:
        dm = torch.as_tensor(drop_mask).to(torch.uint8).reshape(batch, num_gaussians).contiguous().cpu()
        keep.append((C.c_uint8 * dm.numel())(*dm.flatten().tolist()))
        a.drop_mask = keep[-1]
    if small is not None:
        sm = [int(bool(v)) for v in (small.tolist() if isinstance(small, torch.Tensor) else small)]
        keep.append((C.c_uint8 * batch)(*sm))
        a.small = keep[-1]
    if mode == _lib.AFB_POLICY_INTEGRATE:
        if x is None:
            raise AfbError("policy_eval: INTEGRATE needs x")
        _chk(x, torch.float32, "policy_eval x")
        x2 = x.reshape(-1, 64)
        if not x2.is_contiguous() or x2.shape[0] != head.shape[0]:
            raise AfbError("policy_eval: x must be contiguous [batch*tokens, 64]")
        a.x_in = x2.data_ptr()
    out = torch.empty((head.shape[0], 64), dtype=torch.float32, device=head.device)
    out_bf = torch.empty((head.shape[0], 64), dtype=BF16, device=head.device) if want_bf16 else None
    a.out = out.data_ptr()
    a.out_bf16 = out_bf.data_ptr() if want_bf16 else None
    _lib.check(lib.afb_policy_eval(C.byref(a), _stream()), "afb_policy_eval")
    shape = (batch, tokens, 64)
    return (out.reshape(shape), out_bf.reshape(shape)) if want_bf16 else out.reshape(shape)


def policy_backward(head: torch.Tensor, tgt_u: torch.Tensor, sigma_src, sigma_start, sigma_end, coef: float,
                    dhead: Optional[torch.Tensor] = None, small=None, num_gaussians: int = 16, eps: float = 1e-4):
    """dL/dhead for L = coef/2 * sum (policy_average_u(head) - tgt_u)^2, accumulated into `dhead`
    (fp32 [batch*tokens, head_ld]) when given, else written to a new tensor."""
    lib = _lib.load()
    _chk(head, BF16, "policy_backward head")
    tgt_f32 = _target_is_f32(tgt_u, "policy_backward tgt_u")
    batch = len(sigma_src)
    tokens = head.shape[0] // batch
    tg = tgt_u.reshape(-1, 64)
    if head.dim() != 2 or head.stride(1) != 1 or not tg.is_contiguous() or tg.shape[0] != head.shape[0]:
        raise AfbError("policy_backward: head must be [batch*tokens, head_ld], tgt_u contiguous [batch*tokens, 64]")
    accumulate = dhead is not None
    if dhead is None:
        dhead = torch.zeros((head.shape[0], head.shape[1]), dtype=torch.float32, device=head.device)
    _chk(dhead, torch.float32, "policy_backward dhead")
    if dhead.shape != head.shape or dhead.stride(1) != 1:
        raise AfbError("policy_backward: dhead must match head's shape")
    a = _lib.PolicyArgs()
    a.head, a.head_ld, a.batch, a.tokens = head.data_ptr(), head.stride(0), batch, tokens
    a.num_gaussians, a.mode, a.eps = num_gaussians, _lib.AFB_POLICY_AVERAGE_U, eps
    keep = [_host_floats(sigma_src, batch, "sigma_src"), _host_floats(sigma_start, batch, "sigma_start"),
            _host_floats(sigma_end, batch, "sigma_end")]
    a.sigma_src, a.sigma_start, a.sigma_end = keep
    if small is not None:
        sm = [int(bool(v)) for v in (small.tolist() if isinstance(small, torch.Tensor) else small)]
        keep.append((C.c_uint8 * batch)(*sm))
        a.small = keep[-1]
    _lib.check(lib.afb_policy_backward(C.byref(a), tg.data_ptr(), dhead.data_ptr(), dhead.stride(0), float(coef),
                                       int(accumulate), tgt_f32, _stream()), "afb_policy_backward")
    return dhead


def colsum_f32(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[n] += sum over rows of x[:, n] (fp32)."""
    lib = _lib.load()
    _chk(x, torch.float32, "colsum x")
    if x.dim() != 2 or x.stride(1) != 1:
        raise AfbError("colsum: x must be [rows, n] with contiguous columns")
    if out is None:
        out = torch.zeros(x.shape[1], dtype=torch.float32, device=x.device)
    _chk(out, torch.float32, "colsum out")
    _lib.check(lib.afb_colsum_f32(x.data_ptr(), x.stride(0), out.data_ptr(), x.shape[0], x.shape[1], _stream()),
               "afb_colsum_f32")
    return out


def gemm_tn(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[m, n] += sum_t a[t, m] * b[t, n] (weight gradients); a, b bf16 [tokens, *], out fp32 [m, n]."""
    lib = _lib.load()
    _chk(a, BF16, "gemm_tn a")
    _chk(b, BF16, "gemm_tn b")
    if a.dim() != 2 or b.dim() != 2 or a.shape[0] != b.shape[0] or a.stride(1) != 1 or b.stride(1) != 1:
        raise AfbError("gemm_tn: a [tokens, m] and b [tokens, n] with contiguous rows")
    if out is None:
        out = torch.zeros((a.shape[1], b.shape[1]), dtype=torch.float32, device=a.device)
    _chk(out, torch.float32, "gemm_tn out")
    if tuple(out.shape) != (a.shape[1], b.shape[1]) or out.stride(1) != 1:
        raise AfbError("gemm_tn: out must be [m, n] fp32")
    _lib.check(lib.afb_gemm_tn(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0),
                               a.shape[0], a.shape[1], b.shape[1], _stream()), "afb_gemm_tn")
    return out


def ln_mod_param_grad(x: torch.Tensor, dy: torch.Tensor, dscale: Optional[torch.Tensor] = None,
                      dshift: Optional[torch.Tensor] = None, eps: float = 1e-6):
    """dscale[b, d] += sum_rows dy * LN(x), dshift[b, d] += sum_rows dy; x, dy bf16 [batches, rows, dim]."""
    lib = _lib.load()
    _chk(x, BF16, "ln_mod_param_grad x")
    _chk(dy, BF16, "ln_mod_param_grad dy")
    xptr, xld, xbs, nb, nr, dim = _rows3(x, "ln_mod_param_grad x")
    dptr, dld, dbs, db_, dr, dd = _rows3(dy, "ln_mod_param_grad dy")
    if (nb, nr, dim) != (db_, dr, dd) or xld != dim or dld != dim:
        raise AfbError("ln_mod_param_grad: x and dy must share [batches, rows, dim] with contiguous rows")
    if dscale is None:
        dscale = torch.zeros((nb, dim), dtype=torch.float32, device=x.device)
    if dshift is None:
        dshift = torch.zeros((nb, dim), dtype=torch.float32, device=x.device)
    ws = torch.empty(2 * nb * nr, dtype=torch.float32, device=x.device)
    _lib.check(lib.afb_ln_mod_param_grad(xptr, xbs, dptr, dbs, ws.data_ptr(), dscale.data_ptr(), dshift.data_ptr(),
                                         nb, nr, dim, eps, _stream()), "afb_ln_mod_param_grad")
    return dscale, dshift


def rowlinear_param_grad(de: torch.Tensor, t: torch.Tensor, dw: torch.Tensor, dbias: Optional[torch.Tensor] = None,
                         silu_in: bool = True):
    """dw[j, d] += sum_b de[b, j] * act(t[b, d]); dbias[j] += sum_b de[b, j] (fp32 grads of a batch-row Linear)."""
    lib = _lib.load()
    _chk(de, torch.float32, "rowlinear_param_grad de")
    _chk(t, BF16, "rowlinear_param_grad t")
    _chk(dw, torch.float32, "rowlinear_param_grad dw")
    m, n_out = de.shape
    k_in = t.shape[1]
    if t.shape[0] != m or tuple(dw.shape) != (n_out, k_in) or de.stride(1) != 1 or t.stride(1) != 1 or dw.stride(1) != 1:
        raise AfbError("rowlinear_param_grad: shape mismatch")
    _lib.check(lib.afb_rowlinear_param_grad(de.data_ptr(), de.stride(0), t.data_ptr(), t.stride(0), dw.data_ptr(),
                                            dw.stride(0), dbias.data_ptr() if dbias is not None else None, m, n_out,
                                            k_in, int(silu_in), _stream()), "afb_rowlinear_param_grad")
    return dw, dbias


def _target_is_f32(t: torch.Tensor, what: str) -> int:
    """Teacher targets are bf16 (a network output) or fp32 (the true-CFG combination of two)."""
    if not t.is_cuda or t.dtype not in (BF16, torch.float32):
        raise AfbError(f"{what}: expected a CUDA bf16 or fp32 tensor, got {t.dtype} on {t.device}")
    return int(t.dtype == torch.float32)


def cfg_combine(both: torch.Tensor, guidance_scale: float) -> torch.Tensor:
    """[neg; pos] (bf16, batch-doubled network output) -> pos + (pos - neg) * (g - 1), fp32 [batch, ...]."""
    lib = _lib.load()
    _chk(both, BF16, "cfg_combine both")
    if both.shape[0] % 2 or not both.is_contiguous():
        raise AfbError("cfg_combine: need a contiguous tensor with an even batch ([neg; pos])")
    out = torch.empty((both.shape[0] // 2,) + tuple(both.shape[1:]), dtype=torch.float32, device=both.device)
    _lib.check(lib.afb_cfg_combine(both.data_ptr(), out.data_ptr(), out.numel(), float(guidance_scale), _stream()),
               "afb_cfg_combine")
    return out


def axpy_rows(x: torch.Tensor, u: torch.Tensor, coef, want_bf16: bool = False):
    """out[b] = x[b] + coef[b] * u[b]; x fp32, u bf16 or fp32 (teacher velocity), coef host per-sample values."""
    lib = _lib.load()
    _chk(x, torch.float32, "axpy_rows x")
    u_f32 = _target_is_f32(u, "axpy_rows u")
    if x.shape != u.shape or not x.is_contiguous() or not u.is_contiguous():
        raise AfbError("axpy_rows: x and u must be contiguous with the same shape")
    batch = x.shape[0]
    per = x.numel() // batch
    out = torch.empty_like(x)
    out_bf = torch.empty(x.shape, dtype=BF16, device=x.device) if want_bf16 else None
    _lib.check(lib.afb_axpy_rows(x.data_ptr(), u.data_ptr(), _host_floats(coef, batch, "coef"), out.data_ptr(),
                                 out_bf.data_ptr() if want_bf16 else None, batch, per, u_f32, _stream()), "afb_axpy_rows")
    return (out, out_bf) if want_bf16 else out


def mse_rows(pred: torch.Tensor, tgt: torch.Tensor) -> torch.Tensor:
    """Per-sample mean squared error over all trailing dims; pred fp32, tgt bf16 or fp32 -> fp32 [batch] (device)."""
    lib = _lib.load()
    _chk(pred, torch.float32, "mse_rows pred")
    tgt_f32 = _target_is_f32(tgt, "mse_rows tgt")
    if pred.shape != tgt.shape or not pred.is_contiguous() or not tgt.is_contiguous():
        raise AfbError("mse_rows: pred and tgt must be contiguous with the same shape")
    batch = pred.shape[0]
    out = torch.empty(batch, dtype=torch.float32, device=pred.device)
    _lib.check(lib.afb_mse_rows(pred.data_ptr(), tgt.data_ptr(), out.data_ptr(), batch, pred.numel() // batch,
                                tgt_f32, _stream()), "afb_mse_rows")
    return out


# ------------------------------------------------------------------------------------------------
# activation-gradient kernels of the streaming ops
# ------------------------------------------------------------------------------------------------
def ln_modulate_bwd(x: torch.Tensor, dy: torch.Tensor, scale: torch.Tensor, dh: Optional[torch.Tensor] = None,
                    eps: float = 1e-6) -> torch.Tensor:
    """dh (+)= LNmod_bwd(dy); x, dy, dh: bf16 [batches, rows, dim]; scale: [batches, dim] view. dh given -> accumulate."""
    lib = _lib.load()
    xp, xld, xbs, nb, nr, dim = _rows3(x, "ln_modulate_bwd x")
    dp, dld, dbs, *_ = _rows3(dy, "ln_modulate_bwd dy")
    acc = dh is not None
    if dh is None:
        dh = torch.empty((nb, nr, dim), dtype=BF16, device=x.device)
    hp, hld, hbs, *_ = _rows3(dh, "ln_modulate_bwd dh")
    for t in (x, dy, dh, scale):
        _chk(t, BF16, "ln_modulate_bwd")
    if xld != dim or dld != dim or hld != dim:
        raise AfbError("ln_modulate_bwd: rows must be contiguous")
    _lib.check(lib.afb_ln_modulate_bwd(xp, xbs, dp, dbs, hp, hbs, scale.data_ptr(), scale.stride(0), nb, nr, dim, eps,
                                       int(acc), _stream()), "afb_ln_modulate_bwd")
    return dh


def rowscale(x: torch.Tensor, vec: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    xp, xld, xbs, nb, nr, cols = _rows3(x, "rowscale x")
    if out is None:
        out = torch.empty((nb, nr, cols), dtype=BF16, device=x.device)
    op, old, obs, *_ = _rows3(out, "rowscale out")
    _chk(x, BF16, "rowscale x"); _chk(vec, BF16, "rowscale vec"); _chk(out, BF16, "rowscale out")
    _lib.check(lib.afb_rowscale(xp, xld, xbs, vec.data_ptr(), vec.stride(0), op, old, obs, nb, nr, cols, _stream()),
               "afb_rowscale")
    return out


def gelu_bwd(dm: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    """In place: dm *= gelu_tanh'(pre); 2-D views [rows, cols] (any leading dim)."""
    lib = _lib.load()
    _chk(dm, BF16, "gelu_bwd dm"); _chk(pre, BF16, "gelu_bwd pre")
    if dm.dim() != 2 or pre.shape != dm.shape or dm.stride(1) != 1 or pre.stride(1) != 1:
        raise AfbError("gelu_bwd: need 2-D [rows, cols] views with contiguous columns")
    _lib.check(lib.afb_gelu_bwd(dm.data_ptr(), dm.stride(0), pre.data_ptr(), pre.stride(0), dm.shape[0], dm.shape[1],
                                _stream()), "afb_gelu_bwd")
    return dm


def rmsnorm_rope_bwd(dqkv: torch.Tensor, raw: torch.Tensor, q_off: int, k_off: int, heads: int, txt_rows: int,
                     wq_img, wk_img, cos, sin, wq_txt=None, wk_txt=None, eps: float = 1e-6) -> torch.Tensor:
    lib = _lib.load()
    _chk(dqkv, BF16, "rmsnorm_rope_bwd dqkv"); _chk(raw, BF16, "rmsnorm_rope_bwd raw")
    ptr, ld, bs, nb, ns, _ = _rows3(dqkv, "rmsnorm_rope_bwd dqkv")
    rp, rld, rbs, *_ = _rows3(raw, "rmsnorm_rope_bwd raw")
    if (rld, rbs) != (ld, bs):
        raise AfbError("rmsnorm_rope_bwd: dqkv and raw must share strides")
    _lib.check(lib.afb_rmsnorm_rope_bwd(
        ptr, rp, ld, bs, q_off, k_off, nb, ns, heads, txt_rows,
        wq_txt.data_ptr() if wq_txt is not None else None, wk_txt.data_ptr() if wk_txt is not None else None,
        wq_img.data_ptr(), wk_img.data_ptr(), cos.data_ptr(), sin.data_ptr(), eps, _stream()), "afb_rmsnorm_rope_bwd")
    return dqkv


def dropout_rows(x: torch.Tensor, seed: int, layer_id: int, p: float, out: Optional[torch.Tensor] = None,
                 logical_cols: Optional[int] = None, col0: int = 0, silu_in: bool = False, accumulate: bool = False):
    """out (=|+=) keep (.) act(x) / (1 - p); x bf16 [batches, rows, cols] view, counter-based mask (afb_dropout_rows)."""
    lib = _lib.load()
    _chk(x, BF16, "dropout_rows x")
    xp, xld, xbs, nb, nr, cols = _rows3(x, "dropout_rows x")
    if out is None:
        out = torch.zeros((nb, nr, cols), dtype=BF16, device=x.device)
    _chk(out, BF16, "dropout_rows out")
    op, old, obs, *_ = _rows3(out, "dropout_rows out")
    _lib.check(lib.afb_dropout_rows(xp, xld, xbs, op, old, obs, nb, nr, cols, logical_cols or cols, col0, int(seed),
                                    int(layer_id), float(p), int(silu_in), int(accumulate), _stream()), "afb_dropout_rows")
    return out


# ------------------------------------------------------------------------------------------------------------------
# FLUX VAE decoder building blocks (NHWC bf16 activations) — see arcflow_b200/vae.py
# ------------------------------------------------------------------------------------------------------------------
def _nhwc(t: torch.Tensor, name: str):
    _chk(t, BF16, name)
    if t.dim() != 4 or t.stride(3) != 1 or t.stride(1) != t.shape[2] * t.stride(2) or t.stride(0) != t.shape[1] * t.stride(1):
        raise AfbError(f"{name}: need a dense-pixel NHWC tensor [n, h, w, c] (channel slices allowed), got {tuple(t.shape)} "
                       f"strides {t.stride()}")
    return t.data_ptr(), t.stride(2)


def conv3x3(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            res: Optional[torch.Tensor] = None) -> torch.Tensor:
    """3x3 convolution, stride 1, zero padding 1. x: NHWC bf16 [n, h, w, c_in] (c_in % 64 == 0); w: [c_out, 9 * c_in] packed
    tap-major (`pack_conv3x3_weight`); res: added to the output (may alias out). Implicit GEMM on the tcgen05 kernel."""
    lib = _lib.load()
    xp, xld = _nhwc(x, "conv3x3 x")
    _chk(w, BF16, "conv3x3 w")
    n, h, wpx, cin = x.shape
    if w.dim() != 2 or not w.is_contiguous() or w.shape[1] != 9 * cin:
        raise AfbError(f"conv3x3: w must be contiguous [c_out, {9 * cin}], got {tuple(w.shape)}")
    cout = w.shape[0]
    if out is None:
        out = torch.empty((n, h, wpx, cout), dtype=BF16, device=x.device)
    op, old = _nhwc(out, "conv3x3 out")
    if tuple(out.shape) != (n, h, wpx, cout):
        raise AfbError("conv3x3: out shape mismatch")
    d = _lib.ConvDesc()
    d.x, d.w, d.out, d.x_ld, d.out_ld = xp, w.data_ptr(), op, xld, old
    d.n, d.h, d.w_px, d.c_in, d.c_out = n, h, wpx, cin, cout
    d.epilogue = _lib.AFB_EPI_BIAS
    if bias is not None:
        _chk(bias, BF16, "conv3x3 bias")
        d.bias = bias.data_ptr()
    if res is not None:
        rp, rld = _nhwc(res, "conv3x3 res")
        if tuple(res.shape) != tuple(out.shape):
            raise AfbError("conv3x3: res shape mismatch")
        d.res, d.res_ld, d.epilogue = rp, rld, _lib.AFB_EPI_BIAS_RES
    _lib.check(lib.afb_conv3x3(C.byref(d), _stream()), "afb_conv3x3")
    return out


def pack_conv3x3_weight(w: torch.Tensor, c_in_pad: Optional[int] = None, c_out_pad: Optional[int] = None) -> torch.Tensor:
    """torch Conv2d weight [c_out, c_in, 3, 3] -> [c_out_pad, 9 * c_in_pad] bf16, tap-major / channel-minor, zero-padded."""
    co, ci = w.shape[:2]
    cip, cop = c_in_pad or ci, c_out_pad or co
    p = torch.zeros((cop, 3, 3, cip), dtype=BF16, device=w.device)
    p[:co, :, :, :ci] = w.permute(0, 2, 3, 1).to(BF16)
    return p.reshape(cop, 9 * cip).contiguous()


def groupnorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, silu: bool = True, out: Optional[torch.Tensor] = None,
              eps: float = 1e-6) -> torch.Tensor:
    """GroupNorm(32, affine) (+ swish) over NHWC bf16 [n, h, w, c] (contiguous); gamma / beta fp32 [c]."""
    lib = _lib.load()
    _chk(x, BF16, "groupnorm x")
    _chk(gamma, torch.float32, "groupnorm gamma")
    _chk(beta, torch.float32, "groupnorm beta")
    if x.dim() != 4 or not x.is_contiguous():
        raise AfbError("groupnorm: x must be a contiguous NHWC tensor")
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    _chk(out, BF16, "groupnorm out")
    if out.shape != x.shape or not out.is_contiguous():
        raise AfbError("groupnorm: out must be contiguous with x's shape")
    nws = int(lib.afb_groupnorm_ws_floats(n, h * w))
    ws = torch.empty(nws, dtype=torch.float32, device=x.device)
    _lib.check(lib.afb_groupnorm(x.data_ptr(), out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), ws.data_ptr(), nws, n, h * w,
                                 c, eps, int(silu), _stream()), "afb_groupnorm")
    return out


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    _chk(x, BF16, "upsample2x x")
    if x.dim() != 4 or not x.is_contiguous():
        raise AfbError("upsample2x: x must be a contiguous NHWC tensor")
    n, h, w, c = x.shape
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=BF16, device=x.device)
    _lib.check(lib.afb_upsample2x(x.data_ptr(), out.data_ptr(), n, h, w, c, _stream()), "afb_upsample2x")
    return out


def softmax_rows_(x: torch.Tensor) -> torch.Tensor:
    """In-place softmax over the last dim of a bf16 matrix [rows, cols]."""
    lib = _lib.load()
    _chk(x, BF16, "softmax_rows x")
    if x.dim() != 2 or x.stride(1) != 1:
        raise AfbError("softmax_rows: need [rows, cols] with a contiguous last dim")
    _lib.check(lib.afb_softmax_rows(x.data_ptr(), x.stride(0), x.shape[0], x.shape[1], _stream()), "afb_softmax_rows")
    return x


def vae_pre(latents: torch.Tensor, c_pad: int, scale: float, shift: float) -> torch.Tensor:
    """fp32 NCHW latents [n, c, h, w] -> (z / scale + shift) as NHWC bf16 [n, h, w, c_pad] (zero-padded channels)."""
    lib = _lib.load()
    _chk(latents, torch.float32, "vae_pre latents")
    if latents.dim() != 4 or not latents.is_contiguous():
        raise AfbError("vae_pre: latents must be contiguous fp32 [n, c, h, w]")
    n, c, h, w = latents.shape
    out = torch.empty((n, h, w, c_pad), dtype=BF16, device=latents.device)
    _lib.check(lib.afb_vae_pre(latents.data_ptr(), out.data_ptr(), n, c, h, w, c_pad, float(scale), float(shift), _stream()),
               "afb_vae_pre")
    return out


def vae_post(x: torch.Tensor, c_out: int = 3) -> torch.Tensor:
    """NHWC bf16 [n, h, w, >= 8] -> fp32 NCHW [n, c_out, h, w] (first c_out channels)."""
    lib = _lib.load()
    xp, xld = _nhwc(x, "vae_post x")
    n, h, w, _ = x.shape
    out = torch.empty((n, c_out, h, w), dtype=torch.float32, device=x.device)
    _lib.check(lib.afb_vae_post(xp, xld, out.data_ptr(), n, c_out, h, w, _stream()), "afb_vae_post")
    return out
