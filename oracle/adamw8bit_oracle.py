"""ORACLE — TEST INFRASTRUCTURE ONLY. Never imported by the product path (arcflow_b200/, lakonlab/).

PARITY UNPINNED: numpy restatement of bitsandbytes' block-wise 8-bit AdamW — the optimizer the reference's configs select
(`optimizer=dict(diffusion=dict(type='AdamW8bit', ...))`, configs/flux/_ddp_train.py:18-26; the class is registered from
`bitsandbytes.optim` by lakonlab/runner/optimizer/builder.py:11-24 and stepped by BaseModel.step_optimizer,
lakonlab/models/base.py:76-103). bitsandbytes is a third-party dependency that is absent from /root/reference and from this
image, and the reference does not pin its version (requirements.txt:11 `bitsandbytes`). What is restated here is its
published algorithm: Dettmers, Lewis, Shleifer, Zettlemoyer, "8-bit Optimizers via Block-wise Quantization" (ICLR 2022),
in the form of bitsandbytes >= 0.44:
  * `functional.create_dynamic_map(signed, max_exponent_bits=7, total_bits=8)`  -> dynamic_map()
  * `optim.optimizer.Optimizer2State.init_state`: tensors with numel >= min_8bit_size (4096) keep uint8 state1 / state2 (zeros),
    qmap1 = dynamic map signed, qmap2 = dynamic map unsigned, one fp32 absmax per block of 256 elements; smaller tensors keep
    fp32 moments;
  * `kOptimizerStatic8bit2StateBlockwise` (csrc/kernels.cu), ADAM branch: per block
        s1 = qmap1[c1] * absmax1,  s2 = qmap2[c2] * absmax2
        s1 = s1 * beta1 + (1 - beta1) * g,  s2 = s2 * beta2 + (1 - beta2) * g * g      (g already scaled by the clip coefficient)
        absmax1', absmax2' = max |s1|, max |s2| over the block
        p += step_size * s1 / (sqrt(s2) + correction2 * eps),  correction1 = 1 - beta1^t,  correction2 = sqrt(1 - beta2^t),
                                                                step_size = -lr * correction2 / correction1
        p *= 1 - lr * weight_decay                              (only when weight_decay > 0)
        c1' = nearest code of s1 / absmax1' (moved one code towards the sign of s1 if the code's sign differs),
        c2' = nearest code of s2 / absmax2'
    a non-finite gradient element zeroes its two moments and leaves its parameter unchanged.
No golden vector of bitsandbytes exists in the reference's tests; the known values the map must contain (0, 1, the largest
fraction 0.9929688 and the smallest 5.5e-7 of the signed map) are checked in tests/test_oracle_adamw8bit.py.
"""
from __future__ import annotations

import numpy as np

BLOCK = 256
MIN_8BIT_SIZE = 4096


def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """torch.linspace in float32: step = (end - start) / (steps - 1); the first half counts up from start, the second half
    down from end (the symmetric evaluation torch uses)."""
    start, end = np.float32(start), np.float32(end)
    if steps == 1:
        return np.array([start], np.float32)
    step = np.float32((end - start) / np.float32(steps - 1))
    i = np.arange(steps)
    half = steps // 2
    up = (start + step * i.astype(np.float32)).astype(np.float32)
    down = (end - step * (steps - 1 - i).astype(np.float32)).astype(np.float32)
    return np.where(i < half, up, down).astype(np.float32)


def dynamic_map(signed: bool = True, max_exponent_bits: int = 7, total_bits: int = 8) -> np.ndarray:
    data = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    for i in range(max_exponent_bits):
        if signed:
            fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1)
        else:
            fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        b = _linspace_f32(0.1, 1.0, fraction_items)
        means = ((b[:-1] + b[1:]) / np.float32(2.0)).astype(np.float32)
        scaled = (np.float32(10.0 ** (-(max_exponent_bits - 1) + i)) * means).astype(np.float32)
        data += scaled.tolist()
        if signed:
            data += (-scaled).tolist()
    if additional_items > 0:
        b = _linspace_f32(0.1, 1.0, additional_items + 1)
        means = ((b[:-1] + b[1:]) / np.float32(2.0)).astype(np.float32)
        scaled = (np.float32(10.0 ** (-(max_exponent_bits - 1) + i)) * means).astype(np.float32)
        data += scaled.tolist()
        if signed:
            data += (-scaled).tolist()
    data += [0.0, 1.0]
    assert len(data) == 2 ** total_bits
    return np.sort(np.asarray(data, np.float32))


def nearest_code(qmap: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Index of the code-book entry nearest to x (qmap ascending); exact mid-points go to the lower entry."""
    hi = np.clip(np.searchsorted(qmap, x, side="left"), 0, len(qmap) - 1)
    lo = np.clip(hi - 1, 0, len(qmap) - 1)
    pick_hi = (x - qmap[lo]) > (qmap[hi] - x)
    return np.where(pick_hi, hi, lo).astype(np.int64)


def quantize_blockwise(x: np.ndarray, qmap: np.ndarray, block: int = BLOCK):
    """-> (codes uint8 [n], absmax fp32 [n / block]) of a flat fp32 array whose length is a multiple of `block`."""
    xb = x.reshape(-1, block).astype(np.float32)
    absmax = np.abs(xb).max(axis=1).astype(np.float32)
    inv = np.where(absmax > 0, np.float32(1.0) / np.where(absmax > 0, absmax, 1), 0).astype(np.float32)
    codes = nearest_code(qmap, (xb * inv[:, None]).astype(np.float32))
    return codes.reshape(-1).astype(np.uint8), absmax


def dequantize_blockwise(codes: np.ndarray, absmax: np.ndarray, qmap: np.ndarray, block: int = BLOCK) -> np.ndarray:
    return (qmap[codes.reshape(-1, block).astype(np.int64)] * absmax[:, None]).astype(np.float32).reshape(-1)


def adamw8bit_step(p, g, c1, c2, absmax1, absmax2, qmap1, qmap2, step: int, lr, beta1=0.9, beta2=0.95, eps=1e-8,
                   weight_decay=0.0, block: int = BLOCK):
    """One block-wise 8-bit AdamW step on flat arrays (length a multiple of `block`). `lr` may be a scalar or a per-element
    array (parameter groups). Returns (p', c1', c2', absmax1', absmax2'); fp32 arithmetic throughout, as the kernel."""
    f = np.float32
    p, g = p.astype(f), g.astype(f)
    lr = np.broadcast_to(np.asarray(lr, f), p.shape)
    corr1 = f(1.0 - float(beta1) ** step)
    corr2 = f(np.sqrt(f(1.0 - float(beta2) ** step)))
    step_size = (-lr * corr2 / corr1).astype(f)
    s1 = dequantize_blockwise(c1, absmax1, qmap1, block)
    s2 = dequantize_blockwise(c2, absmax2, qmap2, block)
    finite = np.isfinite(g)
    gz = np.where(finite, g, 0).astype(f)
    s2 = np.where(finite, (s2 * f(beta2)).astype(f) + ((f(1.0) - f(beta2)) * gz * gz).astype(f), 0).astype(f)
    s1 = np.where(finite, (s1 * f(beta1)).astype(f) + ((f(1.0) - f(beta1)) * gz).astype(f), 0).astype(f)
    new1 = np.abs(s1.reshape(-1, block)).max(axis=1).astype(f)
    new2 = np.abs(s2.reshape(-1, block)).max(axis=1).astype(f)
    upd = (step_size * (s1 / (np.sqrt(s2).astype(f) + corr2 * f(eps)))).astype(f)
    pn = np.where(finite, p + upd, p).astype(f)
    if weight_decay > 0:
        pn = np.where(finite, pn * (f(1.0) - lr * f(weight_decay)), pn).astype(f)
    inv1 = np.where(new1 > 0, f(1.0) / np.where(new1 > 0, new1, 1), 0).astype(f)
    inv2 = np.where(new2 > 0, f(1.0) / np.where(new2 > 0, new2, 1), 0).astype(f)
    x1 = (s1.reshape(-1, block) * inv1[:, None]).astype(f).reshape(-1)
    x2 = (s2.reshape(-1, block) * inv2[:, None]).astype(f).reshape(-1)
    n1 = nearest_code(qmap1, x1)
    wrong = np.signbit(qmap1[n1]) != np.signbit(s1)
    n1 = np.clip(np.where(wrong, n1 + np.where(s1 > 0, 1, -1), n1), 0, 255)
    n2 = nearest_code(qmap2, x2)
    return pn, n1.astype(np.uint8), n2.astype(np.uint8), new1, new2


def adamw32_step(p, g, m, v, step: int, lr, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0):
    """The fp32-state update bitsandbytes applies to tensors below `min_8bit_size` (kOptimizer32bit2State, ADAM): same update
    rule, decay applied before the step."""
    f = np.float32
    lr = np.broadcast_to(np.asarray(lr, f), p.shape)
    pn = (p * (f(1.0) - lr * f(weight_decay))).astype(f) if weight_decay > 0 else p.astype(f)
    m = (f(beta1) * m + (f(1.0) - f(beta1)) * g).astype(f)
    v = (f(beta2) * v + (f(1.0) - f(beta2)) * g * g).astype(f)
    corr1 = f(1.0 - float(beta1) ** step)
    corr2 = f(1.0 - float(beta2) ** step)
    pn = (pn - lr * (m / corr1) / (np.sqrt(v / corr2).astype(f) + f(eps))).astype(f)
    return pn, m, v
