"""ArcFlow-Qwen-Image on the native engine: config, synthetic weights, RoPE tables, packing, model.

Reference: `_ArcQwenImageTransformer2DModel` (lakonlab/models/architecture/arcflow/arcqwen.py:23-174) over
diffusers 0.35.1 `QwenImageTransformerBlock` / `QwenEmbedRope` / `QwenTimestepProjEmbeddings` (SURVEY.md
Appendix A.4), config configs/qwen/arcqwen_2nfe_k16.py:35-56: 60 double-stream blocks, D = 3072, text width
3584 behind an RMSNorm, timestep-only conditioning, rank-256 LoRA on img_mlp (all blocks), txt_mlp (blocks
0..58) and the timestep embedder. The last block's text-stream output is never read, so it is not computed.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import asdict, dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib, ops
from ._lib import AfbError
from .model import BF16, _cat_k, EngineModelBase
from .schedule import denoise_sigmas


@dataclass
class ArcQwenConfig:
    num_gaussians: int = 16
    logweights_channels: int = 4
    in_channels: int = 64
    out_channels: int = 64
    num_layers: int = 60
    attention_head_dim: int = 128
    num_attention_heads: int = 24
    joint_attention_dim: int = 3584
    axes_dims_rope: Tuple[int, int, int] = (16, 56, 56)
    lora_rank: int = 256
    mlp_ratio: int = 4

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    @property
    def mlp_dim(self) -> int:
        return self.mlp_ratio * self.inner_dim

    @property
    def head_dims(self):
        k = self.num_gaussians
        return k * self.out_channels, k * self.logweights_channels, (k - 1) * self.logweights_channels

    def to_dict(self):
        d = asdict(self)
        d["axes_dims_rope"] = list(self.axes_dims_rope)
        return d


def qwen_image() -> ArcQwenConfig:
    return ArcQwenConfig()


def qwen_tiny(num_layers: int = 3, heads: int = 2) -> ArcQwenConfig:
    return ArcQwenConfig(num_layers=num_layers, num_attention_heads=heads, joint_attention_dim=256)


def qwen_lora_targets(cfg: ArcQwenConfig) -> List[str]:
    """configs/qwen/arcqwen_2nfe_k16.py:47-56 — txt_mlp of the LAST block is excluded (range(num_layers - 1))."""
    t = ["time_text_embed.timestep_embedder.linear_1", "time_text_embed.timestep_embedder.linear_2"]
    for i in range(cfg.num_layers):
        t += [f"transformer_blocks.{i}.img_mlp.net.0.proj", f"transformer_blocks.{i}.img_mlp.net.2"]
        if i < cfg.num_layers - 1:
            t += [f"transformer_blocks.{i}.txt_mlp.net.0.proj", f"transformer_blocks.{i}.txt_mlp.net.2"]
    return t


def qwen_linear_shapes(cfg: ArcQwenConfig) -> Dict[str, tuple]:
    D, M = cfg.inner_dim, cfg.mlp_dim
    s = {"img_in": (D, cfg.in_channels), "txt_in": (D, cfg.joint_attention_dim),
         "time_text_embed.timestep_embedder.linear_1": (D, 256),
         "time_text_embed.timestep_embedder.linear_2": (D, D)}
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        s[p + "img_mod.1"] = (6 * D, D)
        s[p + "txt_mod.1"] = (6 * D, D)
        for n in ("to_q", "to_k", "to_v", "add_q_proj", "add_k_proj", "add_v_proj", "to_out.0", "to_add_out"):
            s[p + "attn." + n] = (D, D)
        for side in ("img_mlp", "txt_mlp"):
            s[p + side + ".net.0.proj"] = (M, D)
            s[p + side + ".net.2"] = (D, M)
    s["norm_out.linear"] = (2 * D, D)
    nm, nw, ng = cfg.head_dims
    s["proj_out_means"], s["proj_out_logweights"], s["proj_out_loggamma"] = (nm, D), (nw, D), (ng, D)
    return s


def make_qwen_state_dict(cfg: ArcQwenConfig, seed: int = 1234, device="cpu", dtype=BF16) -> Dict[str, torch.Tensor]:
    """Seeded synthetic ArcFlow-Qwen student; distributions as in arcflow_b200.synthetic."""
    g = torch.Generator(device=device).manual_seed(seed)

    def normal(shape, std, mean=0.0):
        return torch.empty(shape, device=device, dtype=torch.float32).normal_(mean, std, generator=g).to(dtype)

    sd: Dict[str, torch.Tensor] = {}
    r = cfg.lora_rank
    targets = set(qwen_lora_targets(cfg)) if r > 0 else set()
    for name, (o, i) in qwen_linear_shapes(cfg).items():
        sd[name + ".weight"] = normal((o, i), 0.02)
        sd[name + ".bias"] = normal((o,), 0.02)
        if name in targets:
            sd[name + ".lora_A.weight"] = normal((r, i), 1.0 / r)
            sd[name + ".lora_B.weight"] = normal((o, r), 0.02)
    sd["txt_norm.weight"] = normal((cfg.joint_attention_dim,), 0.02, mean=1.0)
    for i in range(cfg.num_layers):
        for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k"):
            sd[f"transformer_blocks.{i}.attn.{n}.weight"] = normal((cfg.attention_head_dim,), 0.02, mean=1.0)
    gam = torch.logspace(math.log10(0.2), math.log10(4.0), cfg.num_gaussians - 1, base=10).log()
    sd["proj_out_loggamma.bias"] = gam.unsqueeze(1).repeat(1, cfg.logweights_channels).flatten().to(device=device, dtype=dtype)
    return sd


def make_qwen_inputs(cfg: ArcQwenConfig, batch: int, height: int, width: int, txt_len: int = 512, seed: int = 42,
                     device="cpu"):
    g = torch.Generator(device=device).manual_seed(seed)
    gh, gw = height // 16, width // 16
    lat = torch.empty(batch, 16, 2 * gh, 2 * gw, device=device).normal_(generator=g)
    x = lat.view(batch, 16, gh, 2, gw, 2).permute(0, 2, 4, 1, 3, 5).reshape(batch, gh * gw, 64).contiguous()
    txt = torch.empty(batch, txt_len, cfg.joint_attention_dim, device=device).normal_(generator=g).to(BF16)
    return x, txt


def qwen_rope_tables(txt_len: int, grid_h: int, grid_w: int, axes_dims=(16, 56, 56), theta: float = 10000.0,
                     device="cpu"):
    """diffusers QwenEmbedRope(theta, axes_dim, scale_rope=True) for one (1, h, w) image + text, as fp32
    [txt_len + h*w, 128] cos/sin tables in the adjacent-pair layout (complex element j -> columns 2j, 2j+1).
    Text rows come first (the joint sequence is cat([txt, img])); text positions start at max(h//2, w//2)."""
    def rope_params(index: torch.Tensor, dim: int) -> torch.Tensor:
        return torch.outer(index.to(torch.float32),
                           1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))

    def axis(index, i):
        return rope_params(index, axes_dims[i])

    h2, w2 = grid_h // 2, grid_w // 2
    hpos = torch.cat([torch.arange(-(grid_h - h2), 0), torch.arange(0, h2)])
    wpos = torch.cat([torch.arange(-(grid_w - w2), 0), torch.arange(0, w2)])
    f_frame = axis(torch.zeros(1, dtype=torch.long), 0).view(1, 1, -1).expand(grid_h, grid_w, -1)
    f_h = axis(hpos, 1).view(grid_h, 1, -1).expand(grid_h, grid_w, -1)
    f_w = axis(wpos, 2).view(1, grid_w, -1).expand(grid_h, grid_w, -1)
    img = torch.cat([f_frame, f_h, f_w], dim=-1).reshape(grid_h * grid_w, -1)
    tpos = torch.arange(max(h2, w2), max(h2, w2) + txt_len)
    txt = torch.cat([axis(tpos, 0), axis(tpos, 1), axis(tpos, 2)], dim=-1)
    ang = torch.cat([txt, img], dim=0)                       # [S, 64] fp32 angles
    cos = torch.cos(ang).repeat_interleave(2, dim=1).contiguous()
    sin = torch.sin(ang).repeat_interleave(2, dim=1).contiguous()
    return cos.to(device), sin.to(device)


class PackedQwenWeights:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: ArcQwenConfig, device, consume: bool = False):
        self.cfg = cfg
        self.keep: List[torch.Tensor] = []
        self.adapter_views: Dict[str, torch.Tensor] = {}   # see PackedFluxWeights.adapter_views
        self.lora_base: Dict[str, torch.Tensor] = {}
        D, r = cfg.inner_dim, cfg.lora_rank

        def get(name, required=True):
            t = sd.pop(name) if (consume and name in sd) else sd.get(name)
            if t is None:
                if required:
                    raise AfbError(f"state dict is missing '{name}'")
                return None
            return t.to(device=device, dtype=BF16)

        def hold(t, view_name=None):
            if t is None:
                return None
            t = t.contiguous()
            self.keep.append(t)
            if view_name is not None:
                self.adapter_views[view_name] = t
            return t.data_ptr()

        def hold_packed(w_, lb_, prefix):
            t = _cat_k(w_, lb_)
            self.keep.append(t)
            if lb_ is not None:
                self.adapter_views[prefix + ".lora_B.weight"] = t[:, w_.shape[1]:]
            return t.data_ptr()

        def lora(prefix):
            if r <= 0:
                return None, None
            a, b = get(prefix + ".lora_A.weight", False), get(prefix + ".lora_B.weight", False)
            if (a is None) != (b is None):
                raise AfbError(f"LoRA pair incomplete for '{prefix}'")
            return a, b

        w = _lib.Weights()
        w.x_emb_w, w.x_emb_b = hold(get("img_in.weight")), hold(get("img_in.bias"))
        w.ctx_w, w.ctx_b = hold(get("txt_in.weight")), hold(get("txt_in.bias"))
        w.txt_norm_w = hold(get("txt_norm.weight"))
        for li in (1, 2):
            pre = f"time_text_embed.timestep_embedder.linear_{li}"
            setattr(w, f"t{li}_w", hold(get(pre + ".weight")))
            self.lora_base[pre] = self.keep[-1]          # un-packed base weight of a LoRA target (fuse_lora)
            setattr(w, f"t{li}_b", hold(get(pre + ".bias")))
            a, b = lora(pre)
            setattr(w, f"t{li}_la", hold(a, pre + ".lora_A.weight"))
            setattr(w, f"t{li}_lb", hold(b, pre + ".lora_B.weight"))
        mod_w, mod_b, mod_off = [], [], 0

        def add_mod(prefix):
            nonlocal mod_off
            mw, mb = get(prefix + ".weight"), get(prefix + ".bias")
            mod_w.append(mw)
            mod_b.append(mb)
            off = mod_off
            mod_off += mw.shape[0]
            return off

        self.dbl = (_lib.DoubleBlock * max(cfg.num_layers, 1))()
        for i in range(cfg.num_layers):
            p = f"transformer_blocks.{i}."
            k = self.dbl[i]
            k.img_mod_off = add_mod(p + "img_mod.1")
            k.txt_mod_off = add_mod(p + "txt_mod.1")
            for side, names, out_name, ff in (("img", ("to_q", "to_k", "to_v"), "to_out.0", "img_mlp"),
                                              ("txt", ("add_q_proj", "add_k_proj", "add_v_proj"), "to_add_out", "txt_mlp")):
                setattr(k, f"{side}_qkv_w", hold(torch.cat([get(p + f"attn.{n}.weight") for n in names], 0)))
                setattr(k, f"{side}_qkv_b", hold(torch.cat([get(p + f"attn.{n}.bias") for n in names], 0)))
                setattr(k, f"{side}_out_w", hold(get(p + f"attn.{out_name}.weight")))
                setattr(k, f"{side}_out_b", hold(get(p + f"attn.{out_name}.bias")))
                la, lb = lora(p + f"{ff}.net.0.proj")
                setattr(k, f"{side}_up_w", hold_packed(get(p + f"{ff}.net.0.proj.weight"), lb, p + f"{ff}.net.0.proj"))
                setattr(k, f"{side}_up_b", hold(get(p + f"{ff}.net.0.proj.bias")))
                setattr(k, f"{side}_up_la", hold(la, p + f"{ff}.net.0.proj.lora_A.weight"))
                la, lb = lora(p + f"{ff}.net.2")
                setattr(k, f"{side}_down_w", hold_packed(get(p + f"{ff}.net.2.weight"), lb, p + f"{ff}.net.2"))
                setattr(k, f"{side}_down_b", hold(get(p + f"{ff}.net.2.bias")))
                setattr(k, f"{side}_down_la", hold(la, p + f"{ff}.net.2.lora_A.weight"))
            norms = [get(p + f"attn.{n}.weight") for n in ("norm_q", "norm_k", "norm_added_q", "norm_added_k")]
            k.img_nq, k.img_nk, k.txt_nq, k.txt_nk = (hold(t) for t in norms)
            # per-head RMSNorm on q and k bounds the joint attention scores (see PackedFluxWeights)
            k.qk_bound = ops.qk_score_bound((norms[0], norms[1]), (norms[2], norms[3]))
        self.sgl = (_lib.SingleBlock * 1)()
        w.norm_out_mod_off = add_mod("norm_out.linear")
        mod_w_t, mod_b_t = torch.cat(mod_w, 0), torch.cat(mod_b, 0)
        w.mod_w, w.mod_b, w.mod_total = hold(mod_w_t), hold(mod_b_t), mod_off
        self.adapter_views["norm_out.linear.weight"] = mod_w_t[w.norm_out_mod_off:]
        self.adapter_views["norm_out.linear.bias"] = mod_b_t[w.norm_out_mod_off:]
        del mod_w, mod_b
        hw = [get("proj_out_means.weight"), get("proj_out_logweights.weight"), get("proj_out_loggamma.weight")]
        hb = [get("proj_out_means.bias"), get("proj_out_logweights.bias"), get("proj_out_loggamma.bias")]
        n = sum(t.shape[0] for t in hw)
        pad = (-n) % 8
        if pad:
            hw.append(torch.zeros(pad, D, device=device, dtype=BF16))
            hb.append(torch.zeros(pad, device=device, dtype=BF16))
        self.head_w_tensor = torch.cat(hw, 0).contiguous()
        head_b_tensor = torch.cat(hb, 0).contiguous()
        row = 0
        for hn, t in zip(("proj_out_means", "proj_out_logweights", "proj_out_loggamma"), hw):
            self.adapter_views[hn + ".weight"] = self.head_w_tensor[row:row + t.shape[0]]
            self.adapter_views[hn + ".bias"] = head_b_tensor[row:row + t.shape[0]]
            row += t.shape[0]
        w.head_w, w.head_b, w.head_n = hold(self.head_w_tensor), hold(head_b_tensor), n + pad
        self.head_n = n + pad
        w.dbl = C.cast(self.dbl, C.POINTER(_lib.DoubleBlock))
        w.sgl = C.cast(self.sgl, C.POINTER(_lib.SingleBlock))
        self.struct = w

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.keep)


def qwen_time_input(sigma: float) -> float:
    """What the time embedder sees: `timestep.to(bf16)` (arcqwen.py:126), then Timesteps(scale=1000) in fp32."""
    return float(torch.tensor(sigma, dtype=torch.float32).to(BF16).float().item()) * 1000.0


class ArcQwenEngineModel(EngineModelBase):
    """ArcFlow-Qwen-Image student transformer + N-NFE sampler on the native engine (inference and training)."""
    arch = "qwen"
    _DBL_LORA = (("img_up", "img_mlp.net.0.proj"), ("img_down", "img_mlp.net.2"), ("txt_up", "txt_mlp.net.0.proj"),
                 ("txt_down", "txt_mlp.net.2"))

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: ArcQwenConfig, device="cuda",
                 consume_state_dict: bool = False):
        weights = PackedQwenWeights(state_dict, cfg, torch.device(device), consume=consume_state_dict)
        md = _lib.ModelDesc(
            arch=_lib.AFB_ARCH_QWEN, num_double=cfg.num_layers, num_single=0, dim=cfg.inner_dim,
            heads=cfg.num_attention_heads, mlp_dim=cfg.mlp_dim, in_channels=cfg.in_channels,
            txt_dim=cfg.joint_attention_dim, pooled_dim=0, guidance=0, num_gaussians=cfg.num_gaussians,
            lora_rank=cfg.lora_rank, head_mode=0)
        super().__init__(cfg, weights, md, device)

    def rope(self, txt_len, grid_h, grid_w):
        key = (txt_len, grid_h, grid_w)
        if key not in self._rope_cache:
            self._rope_cache[key] = qwen_rope_tables(txt_len, grid_h, grid_w, self.cfg.axes_dims_rope, device=self.device)
        return self._rope_cache[key]

    def _check(self, latents, txt, grid_hw):
        cfg = self.cfg
        if latents.dim() != 3 or latents.shape[2] != cfg.in_channels:
            raise AfbError(f"latents must be [batch, tokens, {cfg.in_channels}], got {tuple(latents.shape)}")
        if txt.dim() != 3 or txt.shape[2] != cfg.joint_attention_dim or txt.shape[0] != latents.shape[0]:
            raise AfbError(f"text embeds must be [batch, txt_len, {cfg.joint_attention_dim}], got {tuple(txt.shape)}")
        if grid_hw[0] * grid_hw[1] != latents.shape[1]:
            raise AfbError(f"token grid {grid_hw} does not match {latents.shape[1]} image tokens")
        if not (latents.is_cuda and txt.is_cuda):
            raise AfbError("inputs must be CUDA tensors (no CPU fallback exists)")

    @torch.no_grad()
    def forward_heads(self, latents, txt, sigma, grid_hw: Sequence[int], train: bool = False) -> torch.Tensor:
        """One network call; `sigma`: a scalar or per-sample values. train=True also stores the per-block checkpoints
        `backward_trunk` recomputes from (see ArcFluxEngineModel.forward_heads)."""
        self._check(latents, txt, grid_hw)
        B, Si, _ = latents.shape
        lat, txt = latents.to(BF16).contiguous(), txt.to(BF16).contiguous()
        self._reserve(B, txt.shape[1], Si)
        sig = [float(v) for v in (sigma.tolist() if isinstance(sigma, torch.Tensor) else
                                  (sigma if isinstance(sigma, (list, tuple)) else [sigma] * B))]
        if len(sig) != B:
            raise AfbError(f"sigma: expected a scalar or {B} per-sample values")
        tdev = torch.tensor([qwen_time_input(v) for v in sig], dtype=torch.float32).to(self.device)
        cos, sin = self.rope(txt.shape[1], grid_hw[0], grid_hw[1])
        out = torch.empty(B, Si, self.weights.head_n, dtype=BF16, device=self.device)
        a = self._fwd_args(txt, None, tdev, None, cos, sin, B, Si)
        a.latents, a.head_out = lat.data_ptr(), out.data_ptr()
        self._launch_forward(a, (lat, txt, tdev, cos, sin, out), train)
        return out

    @torch.no_grad()
    def denoise(self, latents, txt, grid_hw: Sequence[int], num_inference_steps: int = 2, total_substeps: int = 128,
                timestep_ratio: float = 1.0, shift: float = 3.2, eps: float = 1e-4, cuda_graph: bool = False) -> torch.Tensor:
        """The whole N-NFE loop in one C-ABI call (arcqwen_pipeline.py:399-463). cuda_graph=True captures the fixed kernel
        sequence once per (shape, schedule) and replays it, as ArcFluxEngineModel.denoise does."""
        self._check(latents, txt, grid_hw)
        if latents.dtype != torch.float32:
            raise AfbError("denoise: latents must be fp32 packed tokens")
        B, Si, _ = latents.shape
        txt = txt.to(BF16).contiguous()
        self._reserve(B, txt.shape[1], Si)
        sig = denoise_sigmas(num_inference_steps, total_substeps, timestep_ratio, shift)
        tin = [qwen_time_input(s) for s in sig[:-1]]
        cos, sin = self.rope(txt.shape[1], grid_hw[0], grid_hw[1])

        def launch(x, txt_):
            d = _lib.DenoiseArgs()
            d.fwd = self._fwd_args(txt_, None, None, None, cos, sin, B, Si)
            d.nfe = num_inference_steps
            sig_arr, tin_arr = (C.c_float * len(sig))(*sig), (C.c_float * len(tin))(*tin)
            d.sigmas, d.timesteps, d.x, d.eps = sig_arr, tin_arr, x.data_ptr(), eps
            _lib.check(self.lib.afb_engine_denoise(self.handle, C.byref(d), torch.cuda.current_stream().cuda_stream),
                       "afb_engine_denoise")

        if not cuda_graph:
            x = latents.contiguous().clone()
            launch(x, txt)
            return x
        key = (B, txt.shape[1], Si, tuple(grid_hw), tuple(sig), tuple(tin), eps, self._reserved)
        ent = self._graphs.get(key)
        if ent is None:
            st = dict(x=torch.empty_like(latents, memory_format=torch.contiguous_format), txt=torch.empty_like(txt))
            st["x"].copy_(latents), st["txt"].copy_(txt)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up outside capture: one-time function attributes, lazy module load
                launch(st["x"], st["txt"])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                launch(st["x"], st["txt"])
            ent = self._graphs[key] = dict(graph=graph, **st)
        ent["x"].copy_(latents), ent["txt"].copy_(txt)
        ent["graph"].replay()
        return ent["x"].clone()


def make_qwen_teacher_extras(cfg: ArcQwenConfig, seed: int = 4321, device="cpu", dtype=BF16) -> Dict[str, torch.Tensor]:
    """The teacher-only tensors of the stock Qwen-Image transformer (its `norm_out.linear` and `proj_out`); the trunk is
    tied to the student's frozen base layers (lakonlab/models/base_diffusion.py:93-94)."""
    g = torch.Generator(device=device).manual_seed(seed)
    D = cfg.inner_dim
    normal = lambda shape: torch.empty(shape, device=device, dtype=torch.float32).normal_(0.0, 0.02, generator=g).to(dtype)
    return {"norm_out.linear.weight": normal((2 * D, D)), "norm_out.linear.bias": normal((2 * D,)),
            "proj_out.weight": normal((cfg.out_channels, D)), "proj_out.bias": normal((cfg.out_channels,))}


class QwenTeacherEngine(ArcQwenEngineModel):
    """Stock Qwen-Image velocity network TIED to a student's frozen trunk (lakonlab/models/base_diffusion.py:93-94,
    lakonlab/utils/misc.py:116-132), with true classifier-free guidance as the reference's teacher uses it
    (GaussianFlow.forward_u, lakonlab/models/diffusions/gaussian_flow.py:224-254; teacher_guidance_scale = 4.0,
    configs/qwen/arcqwen_2nfe_k16.py:100): ONE batch-doubled forward on [neg; pos] text, then
    pos + (pos - neg)(g - 1) in fp32 (afb_cfg_combine). Forward = lakonlab/models/architecture/diffusers/qwen.py:107-139."""

    def __init__(self, student: ArcQwenEngineModel, teacher_sd: Dict[str, torch.Tensor]):
        from types import SimpleNamespace
        cfg, dev = student.cfg, student.device
        self.student = student   # keeps the shared packed tensors alive
        keep = []

        def hold(name):
            if name not in teacher_sd:
                raise AfbError(f"teacher state dict is missing '{name}'")
            t = teacher_sd[name].to(device=dev, dtype=BF16).contiguous()
            keep.append(t)
            return t

        w = _lib.Weights()
        C.memmove(C.byref(w), C.byref(student.weights.struct), C.sizeof(w))
        pw, pb = hold("proj_out.weight"), hold("proj_out.bias")
        nw, nb = hold("norm_out.linear.weight"), hold("norm_out.linear.bias")
        if pw.shape != (cfg.out_channels, cfg.inner_dim) or nw.shape != (2 * cfg.inner_dim, cfg.inner_dim):
            raise AfbError("teacher proj_out / norm_out.linear have unexpected shapes")
        w.head_w, w.head_b, w.head_n = pw.data_ptr(), pb.data_ptr(), cfg.out_channels
        w.alt_norm_out_w, w.alt_norm_out_b = nw.data_ptr(), nb.data_ptr()
        weights = SimpleNamespace(struct=w, keep=keep, head_n=cfg.out_channels, adapter_views={})
        md = _lib.ModelDesc(
            arch=_lib.AFB_ARCH_QWEN, num_double=cfg.num_layers, num_single=0, dim=cfg.inner_dim,
            heads=cfg.num_attention_heads, mlp_dim=cfg.mlp_dim, in_channels=cfg.in_channels,
            txt_dim=cfg.joint_attention_dim, pooled_dim=0, guidance=0, num_gaussians=cfg.num_gaussians,
            lora_rank=cfg.lora_rank, head_mode=1, ignore_lora=1)
        EngineModelBase.__init__(self, cfg, weights, md, dev)

    def velocity(self, latents, txt_pos, txt_neg, sigma, guidance_scale: float, grid_hw) -> torch.Tensor:
        """u(x_t, t) in packed-token layout [batch, tokens, 64]: bf16 without guidance, fp32 with true CFG."""
        if not guidance_scale > 1.0:
            return self.forward_heads(latents, txt_pos, sigma, grid_hw)
        if txt_neg is None or txt_neg.shape != txt_pos.shape:
            raise AfbError("true CFG needs negative text embeds of the positive ones' shape ([neg; pos] is one batch)")
        B = latents.shape[0]
        sig = [float(v) for v in (sigma.tolist() if isinstance(sigma, torch.Tensor) else
                                  (sigma if isinstance(sigma, (list, tuple)) else [sigma] * B))]
        both = self.forward_heads(torch.cat([latents, latents], 0), torch.cat([txt_neg, txt_pos], 0), sig + sig, grid_hw)
        return ops.cfg_combine(both, guidance_scale)

    def denoise(self, *a, **k):
        raise AfbError("the teacher has no ArcFlow heads; denoise() is a student method")
