"""Step glue after the backward: DDP gradient all-reduce, grad clip, AdamW, Karras EMA — over ONE flat fp32 arena.

Reference:
  BaseModel.step_optimizer                     lakonlab/models/base.py:76-103       clip 50.0 from iteration 100, skip on NaN/Inf
  optimizer / lr config                        configs/flux/_ddp_train.py:13-31     AdamW8bit lr 1e-4 betas (.9,.95) wd 0,
                                                                                    proj_out_loggamma lr x 0.1, linear warm-up 100 it from 1e-3
  ExponentialMovingAverageHookMod (Karras)     lakonlab/runner/hooks/ema_hook.py:86-121, configs/flux/arcflux_2nfe_k16.py:135-145
  DDP all-reduce of the trainable submodule    lakonlab/parallel/ddp_wrapper.py:7-26
All trainable tensors live back to back in one buffer, so the all-reduce is ONE NCCL call on the gradient arena and the
update is two kernel launches (afb_grad_norm_sq, afb_adamw_ema_step) regardless of how many tensors there are.
Two state modes: `state_bits=32` (fp32 moments, torch.optim.AdamW arithmetic) and `state_bits=8` — the block-wise 8-bit
moments of bitsandbytes' AdamW8bit, the optimizer the reference's configs name (`type='AdamW8bit'`): one code byte per
element and moment + one fp32 absmax per 256-element block for tensors of >= 4096 elements, fp32 moments for smaller ones
(bitsandbytes' `min_8bit_size`). bitsandbytes is not vendored / pinned by the reference; the algorithm is restated from its
published form (include/arcflow_b200.h, oracle/adamw8bit_oracle.py).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import AfbError


def warmup_lr(base_lr: float, iteration: int, warmup_iters: int = 100, warmup_ratio: float = 0.001) -> float:
    """mmcv 'linear' warm-up of a fixed-policy LR (SURVEY.md Appendix A.9 iv)."""
    if iteration >= warmup_iters:
        return base_lr
    k = (1 - iteration / warmup_iters) * (1 - warmup_ratio)
    return base_lr * (1 - k)


def karras_momentum(iteration: int, start_iter: int = 100, gamma: float = 7.0, max_momentum: float = 1.0) -> float:
    t = max(iteration + 1 - start_iter, 1)
    return min((1 - 1 / t) ** (gamma + 1), max_momentum)


def create_dynamic_map(signed: bool = True, max_exponent_bits: int = 7, total_bits: int = 8) -> torch.Tensor:
    """bitsandbytes.functional.create_dynamic_map: the 256-entry "dynamic" code book of the 8-bit optimizers. Exponent i of
    `max_exponent_bits` decades (10^(i-6)) carries 2^i (signed) or 2^(i+1) (unsigned) linear fraction steps, given by the
    mid-points of linspace(0.1, 1, steps + 1); plus the values 0 and 1; ascending. Signed: +- both; 254 + 2 = 256 entries."""
    data = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    for i in range(max_exponent_bits):
        fraction_items = int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                             else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1)
        boundaries = torch.linspace(0.1, 1, fraction_items)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data += [0.0, 1.0]
    if len(data) != 2 ** total_bits:
        raise AssertionError(f"dynamic map has {len(data)} entries")
    data.sort()
    return torch.tensor(data, dtype=torch.float32)


QBLOCK = 256            # quantisation block of the 8-bit state (bitsandbytes >= 0.44)
MIN_8BIT_SIZE = 4096    # tensors below this keep fp32 moments (bitsandbytes `min_8bit_size`)


class FlatAdamW:
    def __init__(self, shapes: Dict[str, Tuple[int, ...]], device, lr: float = 1e-4, betas=(0.9, 0.95), eps: float = 1e-8,
                 weight_decay: float = 0.0, lr_mult_key: str = "proj_out_loggamma", lr_mult: float = 0.1,
                 max_norm: float = 50.0, clip_begin_iter: int = 100, clip_skip_ratio: float = 0.0, warmup_iters: int = 100, warmup_ratio: float = 0.001,
                 ema_gamma: float = 7.0, ema_start_iter: int = 100, state_bits: int = 32):
        if state_bits not in (8, 32):
            raise ValueError("state_bits must be 32 (fp32 moments) or 8 (block-wise 8-bit moments, AdamW8bit)")
        self.lib = _lib.load() if torch.device(device).type == "cuda" else None
        self.device = torch.device(device)
        self.state_bits = state_bits

        def numel(n):
            size = 1
            for d in shapes[n]:
                size *= d
            return size

        is_lo = lambda n: lr_mult_key in n
        self.views, off = {}, 0
        self.lo = [0, 0]
        if state_bits == 32:
            # tensors matching lr_mult_key are laid out contiguously so one [begin, end) range carries the multiplier
            names = sorted(shapes, key=lambda n: (not is_lo(n), n))
            for n in names:
                if is_lo(n):
                    self.lo[1] = off + numel(n)
                self.views[n] = (off, tuple(shapes[n]))
                off += (numel(n) + 3) // 4 * 4   # keep every tensor 16-byte aligned
            self.n8 = 0
        else:
            # 8-bit tensors first (slots padded to the quantisation block, so blocks start at tensor starts as in
            # bitsandbytes and never straddle two tensors), small tensors behind them with fp32 moments; the lr-multiplier
            # tensors sit at the seam so they still form ONE contiguous range: [big | big lo | small lo | small]
            small = lambda n: numel(n) < MIN_8BIT_SIZE
            names = sorted(shapes, key=lambda n: (small(n), is_lo(n) != small(n), n))
            self.n8, lo_open = 0, False
            for n in names:
                pad = 4 if small(n) else QBLOCK
                if is_lo(n) and not lo_open:
                    self.lo[0], lo_open = off, True
                self.views[n] = (off, tuple(shapes[n]))
                off += (numel(n) + pad - 1) // pad * pad
                if is_lo(n):
                    self.lo[1] = off if not small(n) else self.views[n][0] + numel(n)
                if not small(n):
                    self.n8 = off
        self.n = off
        z = lambda dt=torch.float32, n=None: torch.zeros(self.n if n is None else n, dtype=dt, device=self.device)
        self.params, self.grads, self.ema = z(), z(), z()
        n32 = self.n - self.n8                         # elements with fp32 moments
        self.exp_avg, self.exp_avg_sq = z(n=n32), z(n=n32)
        if state_bits == 8:
            self.state1, self.state2 = z(torch.uint8, self.n8), z(torch.uint8, self.n8)
            self.absmax1, self.absmax2 = z(n=self.n8 // QBLOCK), z(n=self.n8 // QBLOCK)
            self.qmap1 = create_dynamic_map(signed=True).to(self.device)
            self.qmap2 = create_dynamic_map(signed=False).to(self.device)
        self.shadow = z(torch.bfloat16)
        self.norm_sq = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.skipped = torch.zeros(1, dtype=torch.int32, device=self.device)
        # partials + ticket of the deterministic norm reduction: one per optimizer instance (never shared across streams)
        self.norm_scratch = torch.zeros(max(int(self.lib.afb_grad_norm_scratch_floats()), 1), dtype=torch.float32,
                                        device=self.device) if self.device.type == "cuda" else None
        self.hp = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, lr_mult=lr_mult, max_norm=max_norm,
                       clip_begin_iter=clip_begin_iter, clip_skip_ratio=clip_skip_ratio, warmup_iters=warmup_iters, warmup_ratio=warmup_ratio,
                       ema_gamma=ema_gamma, ema_start_iter=ema_start_iter, state_bits=state_bits)
        self.steps_taken = 0

    _STATE32 = ("params", "exp_avg", "exp_avg_sq", "ema")
    _STATE8 = ("state1", "state2", "absmax1", "absmax2")

    def view(self, buf: torch.Tensor, name: str) -> torch.Tensor:
        off, shape = self.views[name]
        size = 1
        for d in shape:
            size *= d
        return buf[off:off + size].view(shape)

    def param(self, name):
        return self.view(self.params, name)

    def grad(self, name):
        return self.view(self.grads, name)

    def load_params(self, tensors: Dict[str, torch.Tensor]):
        for n, t in tensors.items():
            self.param(n).copy_(t.to(torch.float32))
        self.ema.copy_(self.params)
        self.shadow.copy_(self.params)

    def state_dict(self) -> Dict:
        """Arenas + layout, for bit-exact resume (lakonlab/runner/checkpoint.py)."""
        names = self._STATE32 + (self._STATE8 if self.state_bits == 8 else ())
        sd = {k: getattr(self, k).cpu() for k in names}
        sd.update(steps_taken=self.steps_taken, layout={n: (o, list(s)) for n, (o, s) in self.views.items()}, hp=dict(self.hp))
        return sd

    def load_state_dict(self, sd: Dict):
        bits = int((sd.get("hp") or {}).get("state_bits", 32))
        if bits != self.state_bits:
            raise AfbError(f"optimizer checkpoint holds {bits}-bit moments, this optimizer keeps {self.state_bits}-bit ones")
        layout = {n: (o, tuple(s)) for n, (o, s) in sd["layout"].items()}
        if layout != self.views:
            raise AfbError("optimizer checkpoint was written for a different set of adapter tensors")
        for name in self._STATE32 + (self._STATE8 if self.state_bits == 8 else ()):
            getattr(self, name).copy_(sd[name].to(self.device))
        self.shadow.copy_(self.params)
        self.steps_taken = int(sd["steps_taken"])

    def all_reduce_grads(self):
        """DDP semantics: gradients averaged over ranks; one collective on the whole arena (NCCL on GPU, gloo in CPU tests)."""
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grads)
            self.grads.div_(dist.get_world_size())

    @torch.no_grad()
    def step(self, iteration: int) -> Dict[str, float]:
        hp = self.hp
        if self.device.type != "cuda":
            raise AfbError("FlatAdamW.step needs CUDA tensors (no CPU fallback exists)")
        stream = torch.cuda.current_stream().cuda_stream
        self.all_reduce_grads()
        clip = hp["max_norm"] > 0 and iteration >= hp["clip_begin_iter"]
        _lib.check(self.lib.afb_grad_norm_sq_ws(self.grads.data_ptr(), self.n, self.norm_sq.data_ptr(),
                                                self.norm_scratch.data_ptr(), self.norm_scratch.numel(), stream), "afb_grad_norm_sq_ws")
        lr = warmup_lr(hp["lr"], iteration, hp["warmup_iters"], hp["warmup_ratio"])
        ratio = hp.get("clip_skip_ratio", 0.0)

        def fill(a, begin, end):
            """hyper-parameters + the arena slice [begin, end) (element offsets; fp32 moments are indexed from n8)"""
            f32 = lambda t, o: t.data_ptr() + 4 * o
            a.params, a.grads, a.ema = f32(self.params, begin), f32(self.grads, begin), f32(self.ema, begin)
            a.bf16_shadow, a.n = self.shadow.data_ptr() + 2 * begin, end - begin
            a.lr = lr
            a.beta1, a.beta2, a.eps, a.weight_decay = hp["betas"][0], hp["betas"][1], hp["eps"], hp["weight_decay"]
            a.step = self.steps_taken + 1
            a.max_norm = hp["max_norm"] if clip else 0.0
            a.skip_norm = hp["max_norm"] * ratio if (clip and ratio > 0) else 0.0
            a.grad_norm_sq, a.skipped = self.norm_sq.data_ptr(), self.skipped.data_ptr()
            a.ema_copy = int(iteration < hp["ema_start_iter"])
            a.ema_momentum = karras_momentum(iteration, hp["ema_start_iter"], hp["ema_gamma"])
            a.lr_mult_begin = min(max(self.lo[0] - begin, 0), end - begin)
            a.lr_mult_end = min(max(self.lo[1] - begin, 0), end - begin)
            a.lr_mult = hp["lr_mult"]

        if self.n8 > 0:
            q = _lib.Adamw8bitArgs()
            fill(q.base, 0, self.n8)
            q.state1, q.state2 = self.state1.data_ptr(), self.state2.data_ptr()
            q.absmax1, q.absmax2 = self.absmax1.data_ptr(), self.absmax2.data_ptr()
            q.qmap1, q.qmap2, q.blocksize = self.qmap1.data_ptr(), self.qmap2.data_ptr(), QBLOCK
            _lib.check(self.lib.afb_adamw8bit_ema_step(C.byref(q), stream), "afb_adamw8bit_ema_step")
        a = _lib.AdamwArgs()
        if self.n > self.n8:
            fill(a, self.n8, self.n)
            a.exp_avg, a.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
            _lib.check(self.lib.afb_adamw_ema_step(C.byref(a), stream), "afb_adamw_ema_step")
        a.lr = lr
        skipped = bool(self.skipped.item()) if clip else False
        if not skipped:
            self.steps_taken += 1
        norm = float(self.norm_sq.sqrt().item())
        return {"diffusion_grad_norm": float("nan") if skipped else norm, "skipped": skipped, "lr": a.lr}
