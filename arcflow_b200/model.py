"""ArcFlow-FLUX transformer on the native engine: weight packing + the C-ABI handle.

`ArcFluxEngineModel` is what the reference's `_ArcFluxTransformer2DModel`
(lakonlab/models/architecture/arcflow/arcflux.py:25-257) becomes here: it takes the reference's state
dict (diffusers FLUX keys + ArcFlow adapter keys), packs it once into the layouts the kernels want
(fused QKV, [W | lora_B] K-extended weights, all AdaLN Linears concatenated, the three heads fused and
padded to a multiple of 8 rows) and drives `afb_engine_forward` / `afb_engine_denoise`.
There is no PyTorch compute path in this class.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from types import SimpleNamespace
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import AfbError
from .config import ArcFluxConfig
from .ops import qk_score_bound
from .rope import flux_rope_tables
from .schedule import denoise_sigmas, flux_time_inputs

BF16 = torch.bfloat16


def _cat_k(w: torch.Tensor, lora_b: Optional[torch.Tensor]) -> torch.Tensor:
    """[W | B] along K so that [x | xA^T] @ [W | B]^T = xW^T + (xA^T)B^T (peft scaling alpha/r = 1)."""
    return (torch.cat([w, lora_b], dim=1) if lora_b is not None else w).contiguous()


class PackedFluxWeights:
    """Owns the packed device tensors and the ctypes structs that point into them."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: ArcFluxConfig, device, consume: bool = False):
        self.cfg = cfg
        self.keep: List[torch.Tensor] = []
        # state-dict name of every trainable (adapter) tensor -> the bf16 view inside the packed buffers the engine reads
        # (the optimizer's write-back target: lora_B lives in the K-extension columns of [W | B], the heads are row slices
        # of the fused head weight, norm_out.linear is the last row block of the fused modulation weight)
        self.adapter_views: Dict[str, torch.Tensor] = {}
        self.lora_base: Dict[str, torch.Tensor] = {}
        D = cfg.inner_dim
        r = cfg.lora_rank

        def get(name: str, required: bool = True) -> Optional[torch.Tensor]:
            t = sd.pop(name) if (consume and name in sd) else sd.get(name)
            if t is None:
                if required:
                    raise AfbError(f"state dict is missing '{name}'")
                return None
            return t.to(device=device, dtype=BF16)

        def hold(t: Optional[torch.Tensor], view_name: Optional[str] = None):
            if t is None:
                return None
            t = t.contiguous()
            self.keep.append(t)
            if view_name is not None:
                self.adapter_views[view_name] = t
            return t.data_ptr()

        def hold_packed(w_: torch.Tensor, lb_: Optional[torch.Tensor], prefix: str):
            t = _cat_k(w_, lb_)
            self.keep.append(t)
            if lb_ is not None:
                self.adapter_views[prefix + ".lora_B.weight"] = t[:, w_.shape[1]:]
            return t.data_ptr()

        def lora(prefix: str):
            if r <= 0:
                return None, None
            a = get(prefix + ".lora_A.weight", required=False)
            b = get(prefix + ".lora_B.weight", required=False)
            if (a is None) != (b is None):
                raise AfbError(f"LoRA pair incomplete for '{prefix}'")
            if a is not None and (a.shape[0] != r or b.shape[1] != r):
                raise AfbError(f"LoRA rank mismatch for '{prefix}': {tuple(a.shape)}")
            return a, b

        w = _lib.Weights()
        w.x_emb_w, w.x_emb_b = hold(get("x_embedder.weight")), hold(get("x_embedder.bias"))
        w.ctx_w, w.ctx_b = hold(get("context_embedder.weight")), hold(get("context_embedder.bias"))
        te = "time_text_embed."
        for tag, name, with_lora in (("t", "timestep_embedder", True), ("g", "guidance_embedder", False),
                                     ("p", "text_embedder", False)):
            if tag == "g" and not cfg.guidance_embeds:
                continue
            for li in (1, 2):
                pre = f"{te}{name}.linear_{li}"
                base_w = get(pre + ".weight")
                setattr(w, f"{tag}{li}_w", hold(base_w))
                if with_lora:
                    self.lora_base[pre] = self.keep[-1]      # un-packed base weight of a LoRA target (fuse_lora)
                setattr(w, f"{tag}{li}_b", hold(get(pre + ".bias")))
                if with_lora:
                    a, b = lora(pre)
                    setattr(w, f"{tag}{li}_la", hold(a, pre + ".lora_A.weight"))
                    setattr(w, f"{tag}{li}_lb", hold(b, pre + ".lora_B.weight"))

        mod_w: List[torch.Tensor] = []
        mod_b: List[torch.Tensor] = []
        mod_off = 0

        def add_mod(prefix: str) -> int:
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
            k.img_mod_off = add_mod(p + "norm1.linear")
            k.txt_mod_off = add_mod(p + "norm1_context.linear")
            for side, names, out_name, ff in (("img", ("to_q", "to_k", "to_v"), "to_out.0", "ff"),
                                              ("txt", ("add_q_proj", "add_k_proj", "add_v_proj"), "to_add_out", "ff_context")):
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
            # q and k are RMS-normalised per head, so the attention scores of the joint sequence are bounded by the norm
            # weights alone -> fixed-reference softmax in the attention kernel (afb_attn_desc.score_bound)
            k.qk_bound = qk_score_bound((norms[0], norms[1]), (norms[2], norms[3]))

        self.sgl = (_lib.SingleBlock * max(cfg.num_single_layers, 1))()
        for i in range(cfg.num_single_layers):
            p = f"single_transformer_blocks.{i}."
            k = self.sgl[i]
            k.mod_off = add_mod(p + "norm.linear")
            k.qkv_w = hold(torch.cat([get(p + f"attn.{n}.weight") for n in ("to_q", "to_k", "to_v")], 0))
            k.qkv_b = hold(torch.cat([get(p + f"attn.{n}.bias") for n in ("to_q", "to_k", "to_v")], 0))
            nq_t, nk_t = get(p + "attn.norm_q.weight"), get(p + "attn.norm_k.weight")
            k.nq, k.nk = hold(nq_t), hold(nk_t)
            k.qk_bound = qk_score_bound((nq_t, nk_t))
            la, lb = lora(p + "proj_mlp")
            k.mlp_w = hold_packed(get(p + "proj_mlp.weight"), lb, p + "proj_mlp")
            k.mlp_b, k.mlp_la = hold(get(p + "proj_mlp.bias")), hold(la, p + "proj_mlp.lora_A.weight")
            la, lb = lora(p + "proj_out")
            k.out_w = hold_packed(get(p + "proj_out.weight"), lb, p + "proj_out")
            k.out_b, k.out_la = hold(get(p + "proj_out.bias")), hold(la, p + "proj_out.lora_A.weight")

        w.norm_out_mod_off = add_mod("norm_out.linear")
        mod_w_t, mod_b_t = torch.cat(mod_w, 0), torch.cat(mod_b, 0)
        w.mod_w, w.mod_b = hold(mod_w_t), hold(mod_b_t)
        w.mod_total = mod_off
        self.adapter_views["norm_out.linear.weight"] = mod_w_t[w.norm_out_mod_off:]
        self.adapter_views["norm_out.linear.bias"] = mod_b_t[w.norm_out_mod_off:]
        del mod_w, mod_b

        heads_w = [get("proj_out_means.weight"), get("proj_out_logweights.weight"), get("proj_out_loggamma.weight")]
        heads_b = [get("proj_out_means.bias"), get("proj_out_logweights.bias"), get("proj_out_loggamma.bias")]
        n = sum(t.shape[0] for t in heads_w)
        pad = (-n) % 8
        if pad:
            heads_w.append(torch.zeros(pad, D, device=device, dtype=BF16))
            heads_b.append(torch.zeros(pad, device=device, dtype=BF16))
        self.head_w_tensor = torch.cat(heads_w, 0).contiguous()
        head_b_tensor = torch.cat(heads_b, 0).contiguous()
        row = 0
        for hn, hw in zip(("proj_out_means", "proj_out_logweights", "proj_out_loggamma"), heads_w):
            self.adapter_views[hn + ".weight"] = self.head_w_tensor[row:row + hw.shape[0]]
            self.adapter_views[hn + ".bias"] = head_b_tensor[row:row + hw.shape[0]]
            row += hw.shape[0]
        w.head_w, w.head_b, w.head_n = hold(self.head_w_tensor), hold(head_b_tensor), n + pad
        self.head_n = n + pad
        w.dbl = C.cast(self.dbl, C.POINTER(_lib.DoubleBlock))
        w.sgl = C.cast(self.sgl, C.POINTER(_lib.SingleBlock))
        self.struct = w

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.keep)


class EngineModelBase:
    """Shared plumbing of the engine-backed transformers: handle lifetime, workspace, RoPE cache, profiling."""

    def __init__(self, cfg, weights, model_desc: "_lib.ModelDesc", device="cuda"):
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise AfbError("the engine needs a CUDA device (there is no CPU path)")
        self.num_gaussians = cfg.num_gaussians
        self.dtype = BF16
        self.weights = weights
        h = C.c_void_p()
        _lib.check(self.lib.afb_engine_create(C.byref(model_desc), C.byref(h)), "afb_engine_create")
        self.handle = h
        _lib.check(self.lib.afb_engine_bind(self.handle, C.byref(self.weights.struct)), "afb_engine_bind")
        self._rope_cache = {}
        self._graphs = {}            # captured denoise loops, keyed by shape + schedule (see denoise(cuda_graph=True))
        self._reserved = (0, 0, 0)

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.afb_engine_destroy(h)
            except Exception:
                pass
            self.handle = None

    # -- attributes the pipelines read off `pipe.transformer` -----------------------
    @property
    def config(self):
        return SimpleNamespace(**self.cfg.to_dict())

    def cache_context(self, name: str):
        return contextlib.nullcontext()

    def _reserve(self, batch: int, txt_len: int, img_len: int):
        rb, rt, ri = self._reserved
        if batch > rb or txt_len > rt or img_len > ri:
            nb, nt, ni = max(batch, rb), max(txt_len, rt), max(img_len, ri)
            _lib.check(self.lib.afb_engine_reserve(self.handle, nb, nt, ni), "afb_engine_reserve")
            self._reserved = (nb, nt, ni)
            self._graphs.clear()     # captured graphs point into the old workspace

    def set_profiling(self, on: bool):
        _lib.check(self.lib.afb_engine_set_profiling(self.handle, int(on)), "afb_engine_set_profiling")

    def read_profile(self) -> dict:
        p = _lib.Profile()
        _lib.check(self.lib.afb_engine_read_profile(self.handle, C.byref(p)), "afb_engine_read_profile")
        return {n: getattr(p, n) for n, _ in p._fields_}

    def export_activation(self, which: str, batch: int, txt_len: int, img_len: int) -> torch.Tensor:
        """Copy of an internal activation of the last forward: 'hidden' (final image hidden states), 'head_in'
        (norm_out output = A operand of the head GEMM), 'temb'. Training keeps them for the backward."""
        idx = {"hidden": 0, "head_in": 1, "temb": 2}[which]
        D = self.cfg.inner_dim
        shape = (batch, D) if idx == 2 else (batch, img_len, D)
        out = torch.empty(shape, dtype=BF16, device=self.device)
        _lib.check(self.lib.afb_engine_export(self.handle, idx, out.data_ptr(), batch, txt_len, img_len,
                                              torch.cuda.current_stream().cuda_stream), "afb_engine_export")
        return out

    def workspace_bytes(self, batch: int, txt_len: int, img_len: int) -> int:
        return int(self.lib.afb_engine_workspace_bytes(self.handle, batch, txt_len, img_len))

    def _fwd_args(self, txt, pooled, timestep, guidance, cos, sin, batch, img_len) -> "_lib.ForwardArgs":
        a = _lib.ForwardArgs()
        a.batch, a.txt_len, a.img_len = batch, txt.shape[1], img_len
        a.txt = txt.data_ptr()
        a.pooled = pooled.data_ptr() if pooled is not None else None
        a.timestep = timestep.data_ptr() if timestep is not None else None
        a.guidance = guidance.data_ptr() if guidance is not None else None
        a.rope_cos, a.rope_sin = cos.data_ptr(), sin.data_ptr()
        return a

    # ---- training: checkpointed forward + adapter-only backward (shared by the FLUX and Qwen students) -------------
    # LoRA tensor of the engine's block structs -> state-dict suffix (export_arcflow_to_diffusers.py:104-127 names);
    # subclasses set these. Tensors the state dict does not carry (Qwen: txt_mlp of the last block) are skipped.
    _DBL_LORA: tuple = ()
    _SGL_LORA: tuple = ()

    def set_lora_scale(self, scale: float):
        """Runtime adapter scale (peft `scaling` = alpha / r x the weight given to `pipe.set_adapters` /
        `joint_attention_kwargs={'scale': s}`): every LoRA branch contributes scale x B(A(x)). Inference only."""
        if float(scale) == getattr(self, "_lora_scale", 1.0):
            return
        if getattr(self, "lora_fused", False):
            raise AfbError("the adapter was merged into the base weights (fuse_lora) with its scale; re-load it to change it")
        _lib.check(self.lib.afb_engine_set_lora_scale(self.handle, float(scale)), "afb_engine_set_lora_scale")
        self._lora_scale = float(scale)
        self._graphs.clear()     # the captured A-projection launches carry the old scale

    @torch.no_grad()
    def fuse_lora(self):
        """Merge the adapter's low-rank branches into the base weights, W <- W + scale * B A, and stop computing them
        (diffusers' `fuse_lora()`): removes the 2 * r * (in + out) FLOPs per token of every adapted
        Linear (5.6 % of a FLUX forward) and the A-projection launches. The merge itself runs on the GEMM kernel
        (dY W mode with A as the [K = r, N = in] operand, residual epilogue, in place on the packed [W | B] buffers).
        The merged weights are rounded to bf16 once, so outputs differ from the un-merged path at the bf16 level.
        One-way: re-load the adapter to get the separate branch (and runtime `scale`, training) back."""
        if getattr(self, "lora_fused", False):
            return self
        views = self.weights.adapter_views
        for name, a in views.items():
            if not name.endswith(".lora_A.weight"):
                continue
            pre = name[:-len(".lora_A.weight")]
            b = views[pre + ".lora_B.weight"]
            if pre in getattr(self.weights, "lora_base", {}):
                w = self.weights.lora_base[pre]
            else:    # lora_B is the K-extension of the packed [W | B]: W is the same rows, the `in` columns before it
                w = torch.as_strided(b, (b.shape[0], a.shape[1]), b.stride(), b.storage_offset() - a.shape[1])
            from . import ops
            ops.gemm(b.unsqueeze(0), a, w.unsqueeze(0), epilogue=_lib.AFB_EPI_BIAS_RES, res=w.unsqueeze(0), transposed=True,
                     alpha=getattr(self, "_lora_scale", 1.0))
        _lib.check(self.lib.afb_engine_set_ignore_lora(self.handle, 1), "afb_engine_set_ignore_lora")
        self.lora_fused = True
        self._graphs.clear()
        return self

    def set_lora_dropout(self, p: float, seed: int = 0):
        """peft lora_dropout for the NEXT forward_heads(train=True) and its backward (counter-based mask, see
        afb_engine_set_lora_dropout); p = 0 disables. Inference forwards never drop."""
        _lib.check(self.lib.afb_engine_set_lora_dropout(self.handle, float(p), int(seed) & 0xFFFFFFFFFFFFFFFF),
                   "afb_engine_set_lora_dropout")

    def set_activation_stash(self, mode="auto", headroom_bytes: int = 12 << 30):
        """Keep each block's outputs of the train forward for the backward instead of recomputing the block
        (afb_engine_set_activation_stash). mode: True / False / "auto" = on when the device has room for it plus
        `headroom_bytes` (FLUX bs 4 at 1024 px: ~62 GB — fits a 180 GB B200 next to weights, checkpoints and optimizer)."""
        self._stash_mode = mode
        self._stash_headroom = int(headroom_bytes)
        self._stash_shape = None      # decided per shape at the next train forward

    def _apply_stash(self):
        mode = getattr(self, "_stash_mode", "auto")
        if getattr(self, "_stash_shape", None) == (mode, self._reserved):
            return
        on = False
        if mode is True or mode == "auto":
            need = int(self.lib.afb_engine_stash_bytes(self.handle, *self._reserved))
            free, _ = torch.cuda.mem_get_info(self.device)
            cached = torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
            on = mode is True or need + getattr(self, "_stash_headroom", 12 << 30) <= free + cached
            if on and need > free:
                torch.cuda.empty_cache()
        rc = self.lib.afb_engine_set_activation_stash(self.handle, int(on), *self._reserved)
        if rc != 0 and mode == "auto":     # the device could not hold it after all: keep recomputing
            rc = self.lib.afb_engine_set_activation_stash(self.handle, 0, *self._reserved)
            on = False
        _lib.check(rc, "afb_engine_set_activation_stash")
        self.activation_stash = on
        self._stash_shape = (mode, self._reserved)

    def _launch_forward(self, a: "_lib.ForwardArgs", keep: tuple, train: bool):
        stream = torch.cuda.current_stream().cuda_stream
        if train and getattr(self, "lora_fused", False):
            raise AfbError("the adapter was merged into the base weights (fuse_lora): training needs the separate branch")
        if train:
            _lib.check(self.lib.afb_engine_train_reserve(self.handle, *self._reserved), "afb_engine_train_reserve")
            self._apply_stash()
            _lib.check(self.lib.afb_engine_forward_train(self.handle, C.byref(a), stream), "afb_engine_forward_train")
            self._train_ctx = dict(args=a, keep=keep)   # inputs stay alive until the backward has run
        else:
            _lib.check(self.lib.afb_engine_forward(self.handle, C.byref(a), stream), "afb_engine_forward")

    def head_weight(self) -> torch.Tensor:
        """The fused [head_n, D] head weight (means | logits | loggamma | pad) the engine reads."""
        return self.weights.head_w_tensor

    def trunk_lora_shapes(self) -> Dict[str, tuple]:
        views = self.weights.adapter_views
        return {n: tuple(views[n].shape) for n in self.trunk_lora_names()}

    def trunk_lora_names(self):
        """State-dict names of the LoRA tensors the trunk backward produces gradients for."""
        names = []
        for i in range(self.cfg.num_layers):
            names += [f"transformer_blocks.{i}.{n}.lora_{ab}.weight" for _, n in self._DBL_LORA for ab in "AB"]
        for i in range(getattr(self.cfg, "num_single_layers", 0)):
            names += [f"single_transformer_blocks.{i}.{n}.lora_{ab}.weight" for _, n in self._SGL_LORA for ab in "AB"]
        views = self.weights.adapter_views
        return [n for n in names if n in views]

    _TEMB_LORA = (("t1", "time_text_embed.timestep_embedder.linear_1"), ("t2", "time_text_embed.timestep_embedder.linear_2"))

    def embed_lora_shapes(self) -> Dict[str, tuple]:
        views = self.weights.adapter_views
        names = [f"{name}.lora_{ab}.weight" for _, name in self._TEMB_LORA for ab in "AB"]
        return {n: tuple(views[n].shape) for n in names if n in views}

    @property
    def mod_total(self) -> int:
        return int(self.weights.struct.mod_total)

    @property
    def norm_out_mod_off(self) -> int:
        return int(self.weights.struct.norm_out_mod_off)

    @torch.no_grad()
    def backward_embed(self, d_mod: torch.Tensor, grads: Dict[str, torch.Tensor]) -> None:
        """d_mod (fp32 [batch, mod_total], every AdaLN vector's gradient) -> timestep-embedder LoRA gradients (+=)."""
        ctx = getattr(self, "_train_ctx", None)
        if ctx is None:
            raise AfbError("backward_embed: run forward_heads(train=True) first")
        a = ctx["args"]
        if d_mod.dtype != torch.float32 or tuple(d_mod.shape) != (a.batch, self.mod_total) or not d_mod.is_contiguous():
            raise AfbError(f"d_mod must be contiguous fp32 [{a.batch}, {self.mod_total}]")
        g = _lib.EmbedGrads()
        for tag, name in self._TEMB_LORA:
            for ab in "ab":
                t = grads.get(f"{name}.lora_{ab.upper()}.weight")
                setattr(g, f"{tag}_l{ab}", t.data_ptr() if t is not None else None)
        _lib.check(self.lib.afb_engine_backward_embed(self.handle, C.byref(a), d_mod.data_ptr(), C.byref(g),
                                                      torch.cuda.current_stream().cuda_stream), "afb_engine_backward_embed")

    @torch.no_grad()
    def backward_trunk(self, d_head_in: torch.Tensor, grads: Dict[str, torch.Tensor],
                       d_mod: Optional[torch.Tensor] = None) -> None:
        """Accumulate (+=) the LoRA gradients of the last `forward_heads(train=True)` into `grads` (fp32 tensors keyed by
        state-dict name, shapes of the LoRA tensors; missing names are skipped). d_head_in: bf16 [batch, tokens, dim],
        the gradient w.r.t. the norm_out output. Replaces torch autograd + checkpointing through the diffusers blocks
        (lakonlab/models/base_diffusion.py:14-62)."""
        ctx = getattr(self, "_train_ctx", None)
        if ctx is None:
            raise AfbError("backward_trunk: run forward_heads(train=True) first")
        a = ctx["args"]
        D = self.cfg.inner_dim
        if d_head_in.dtype != BF16 or tuple(d_head_in.shape) != (a.batch, a.img_len, D) or not d_head_in.is_contiguous():
            raise AfbError(f"d_head_in must be contiguous bf16 [{a.batch}, {a.img_len}, {D}]")

        def ptr(name):
            t = grads.get(name)
            if t is None:
                return None
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise AfbError(f"gradient buffer '{name}' must be a contiguous fp32 CUDA tensor")
            return t.data_ptr()

        dbl = (_lib.DoubleBlockGrads * max(self.cfg.num_layers, 1))()
        for i in range(self.cfg.num_layers):
            for field, n in self._DBL_LORA:
                setattr(dbl[i], field + "_la", ptr(f"transformer_blocks.{i}.{n}.lora_A.weight"))
                setattr(dbl[i], field + "_lb", ptr(f"transformer_blocks.{i}.{n}.lora_B.weight"))
        n_single = getattr(self.cfg, "num_single_layers", 0)
        sgl = (_lib.SingleBlockGrads * max(n_single, 1))()
        for i in range(n_single):
            for field, n in self._SGL_LORA:
                setattr(sgl[i], field + "_la", ptr(f"single_transformer_blocks.{i}.{n}.lora_A.weight"))
                setattr(sgl[i], field + "_lb", ptr(f"single_transformer_blocks.{i}.{n}.lora_B.weight"))
        b = _lib.BackwardArgs()
        b.fwd = a
        b.d_head_in = d_head_in.data_ptr()
        b.dbl = C.cast(dbl, C.POINTER(_lib.DoubleBlockGrads))
        b.sgl = C.cast(sgl, C.POINTER(_lib.SingleBlockGrads))
        if d_mod is not None:
            if d_mod.dtype != torch.float32 or tuple(d_mod.shape) != (a.batch, self.mod_total) or not d_mod.is_contiguous():
                raise AfbError(f"d_mod must be contiguous fp32 [{a.batch}, {self.mod_total}]")
            b.d_mod = d_mod.data_ptr()
        _lib.check(self.lib.afb_engine_backward(self.handle, C.byref(b), torch.cuda.current_stream().cuda_stream),
                   "afb_engine_backward")

    def split_heads(self, head: torch.Tensor):
        """Raw head tensor -> the reference's ArcFlowModelOutput fields in token layout
        (means [B,S,K,64], logweights [B,S,K,4] log-softmaxed over K in bf16, loggammas [B,S,K-1,4])."""
        cfg = self.cfg
        nm, nw, ng = cfg.head_dims
        B, S, _ = head.shape
        means = head[..., :nm].reshape(B, S, cfg.num_gaussians, cfg.out_channels)
        logw = head[..., nm:nm + nw].reshape(B, S, cfg.num_gaussians, cfg.logweights_channels).log_softmax(dim=-2)
        gam = head[..., nm + nw:nm + nw + ng].reshape(B, S, cfg.num_gaussians - 1, cfg.logweights_channels)
        return dict(means=means, logweights=logw, loggammas=gam)


class ArcFluxEngineModel(EngineModelBase):
    """The ArcFlow-FLUX student transformer + N-NFE sampler on the native engine (inference and training)."""
    arch = "flux"
    _DBL_LORA = (("img_up", "ff.net.0.proj"), ("img_down", "ff.net.2"), ("txt_up", "ff_context.net.0.proj"),
                 ("txt_down", "ff_context.net.2"))
    _SGL_LORA = (("mlp", "proj_mlp"), ("out", "proj_out"))

    def __init__(self, state_dict: Dict[str, torch.Tensor], cfg: ArcFluxConfig, device="cuda",
                 consume_state_dict: bool = False):
        weights = PackedFluxWeights(state_dict, cfg, torch.device(device), consume=consume_state_dict)
        md = _lib.ModelDesc(
            arch=_lib.AFB_ARCH_FLUX, num_double=cfg.num_layers, num_single=cfg.num_single_layers,
            dim=cfg.inner_dim, heads=cfg.num_attention_heads, mlp_dim=cfg.mlp_dim,
            in_channels=cfg.in_channels, txt_dim=cfg.joint_attention_dim, pooled_dim=cfg.pooled_projection_dim,
            guidance=int(cfg.guidance_embeds), num_gaussians=cfg.num_gaussians, lora_rank=cfg.lora_rank,
            head_mode=0)
        super().__init__(cfg, weights, md, device)

    def rope(self, txt_len: int, grid_h: int, grid_w: int):
        key = (txt_len, grid_h, grid_w)
        if key not in self._rope_cache:
            self._rope_cache[key] = flux_rope_tables(txt_len, grid_h, grid_w, self.cfg.axes_dims_rope,
                                                     round_bf16=True, device=self.device)
        return self._rope_cache[key]

    def _check_inputs(self, latents, txt, pooled, grid_hw):
        cfg = self.cfg
        if latents.dim() != 3 or latents.shape[2] != cfg.in_channels:
            raise AfbError(f"latents must be [batch, tokens, {cfg.in_channels}], got {tuple(latents.shape)}")
        if txt.dim() != 3 or txt.shape[2] != cfg.joint_attention_dim or txt.shape[0] != latents.shape[0]:
            raise AfbError(f"text embeds must be [batch, txt_len, {cfg.joint_attention_dim}], got {tuple(txt.shape)}")
        if pooled is None or tuple(pooled.shape) != (latents.shape[0], cfg.pooled_projection_dim):
            raise AfbError("pooled projections must be [batch, pooled_projection_dim]")
        if grid_hw[0] * grid_hw[1] != latents.shape[1]:
            raise AfbError(f"token grid {grid_hw} does not match {latents.shape[1]} image tokens")
        for t in (latents, txt, pooled):
            if not t.is_cuda:
                raise AfbError("inputs must be CUDA tensors (no CPU fallback exists)")

    @torch.no_grad()
    def forward_heads(self, latents: torch.Tensor, txt: torch.Tensor, pooled: torch.Tensor,
                      sigma, guidance_scale: float, grid_hw: Sequence[int], train: bool = False) -> torch.Tensor:
        """One network call. `sigma`: a scalar or per-sample values. Returns the raw head tensor bf16
        [batch, tokens, head_n] (means K*64 | logits K*4 | loggamma (K-1)*4 | pad) — logits NOT yet log-softmaxed
        (for a teacher engine: the velocity [batch, tokens, 64]). train=True also stores the per-block residual-stream
        checkpoints `backward_trunk` recomputes from."""
        self._check_inputs(latents, txt, pooled, grid_hw)
        B, Si, _ = latents.shape
        lat = latents.to(BF16).contiguous()
        txt = txt.to(BF16).contiguous()
        pooled = pooled.to(BF16).contiguous()
        self._reserve(B, txt.shape[1], Si)
        sig = [float(v) for v in (sigma.tolist() if isinstance(sigma, torch.Tensor) else
                                  (sigma if isinstance(sigma, (list, tuple)) else [sigma] * B))]
        if len(sig) != B:
            raise AfbError(f"sigma: expected a scalar or {B} per-sample values")
        tdev = torch.tensor([flux_time_inputs(v, guidance_scale)[0] for v in sig], dtype=torch.float32).to(self.device)
        g_in = flux_time_inputs(sig[0], guidance_scale)[1]
        gdev = torch.full((B,), g_in, dtype=torch.float32, device=self.device) if self.cfg.guidance_embeds else None
        cos, sin = self.rope(txt.shape[1], grid_hw[0], grid_hw[1])
        out = torch.empty(B, Si, self.weights.head_n, dtype=BF16, device=self.device)
        a = self._fwd_args(txt, pooled, tdev, gdev, cos, sin, B, Si)
        a.latents, a.head_out = lat.data_ptr(), out.data_ptr()
        self._launch_forward(a, (lat, txt, pooled, tdev, gdev, cos, sin, out), train)
        return out

    @torch.no_grad()
    def denoise(self, latents: torch.Tensor, txt: torch.Tensor, pooled: torch.Tensor, grid_hw: Sequence[int],
                num_inference_steps: int = 2, total_substeps: int = 128, timestep_ratio: float = 1.0,
                shift: float = 3.2, guidance_scale: float = 3.5, eps: float = 1e-4, cuda_graph: bool = False) -> torch.Tensor:
        """The whole N-NFE loop (network + analytic momentum integration) in one C-ABI call.
        latents: fp32 packed tokens [batch, tokens, 64]; returns the final fp32 packed latents.
        cuda_graph=True captures the call's fixed kernel sequence once per (shape, schedule) and replays it — the engine
        does no host sync and no allocation, so the ~670 launches per NFE collapse into one graph launch (what matters for
        small batches / resolutions, where the step is launch-bound; replaces the reference's eager Python loop,
        arcflux_pipeline.py:453-524)."""
        self._check_inputs(latents, txt, pooled, grid_hw)
        if latents.dtype != torch.float32:
            raise AfbError("denoise: latents must be fp32 (the sampler state is fp32, arcflux_pipeline.py:407)")
        B, Si, _ = latents.shape
        txt = txt.to(BF16).contiguous()
        pooled = pooled.to(BF16).contiguous()
        self._reserve(B, txt.shape[1], Si)
        sig = denoise_sigmas(num_inference_steps, total_substeps, timestep_ratio, shift)
        tin = [flux_time_inputs(s, guidance_scale)[0] for s in sig[:-1]]
        g_in = flux_time_inputs(sig[0], guidance_scale)[1]
        cos, sin = self.rope(txt.shape[1], grid_hw[0], grid_hw[1])

        def launch(x, txt_, pooled_, gdev_):
            d = _lib.DenoiseArgs()
            d.fwd = self._fwd_args(txt_, pooled_, None, gdev_, cos, sin, B, Si)
            d.nfe = num_inference_steps
            sig_arr = (C.c_float * len(sig))(*sig)
            tin_arr = (C.c_float * len(tin))(*tin)
            d.sigmas, d.timesteps = sig_arr, tin_arr
            d.x = x.data_ptr()
            d.eps = eps
            _lib.check(self.lib.afb_engine_denoise(self.handle, C.byref(d), torch.cuda.current_stream().cuda_stream),
                       "afb_engine_denoise")

        mk_g = lambda: (torch.full((B,), g_in, dtype=torch.float32, device=self.device)
                        if self.cfg.guidance_embeds else None)
        if not cuda_graph:
            x = latents.contiguous().clone()
            launch(x, txt, pooled, mk_g())
            return x
        key = (B, txt.shape[1], Si, tuple(grid_hw), tuple(sig), tuple(tin), g_in, eps, self._reserved)
        ent = self._graphs.get(key)
        if ent is None:
            st = dict(x=torch.empty_like(latents, memory_format=torch.contiguous_format), txt=torch.empty_like(txt),
                      pooled=torch.empty_like(pooled), g=mk_g())
            st["x"].copy_(latents), st["txt"].copy_(txt), st["pooled"].copy_(pooled)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):      # warm-up outside capture: one-time function attributes, lazy module load
                launch(st["x"], st["txt"], st["pooled"], st["g"])
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                launch(st["x"], st["txt"], st["pooled"], st["g"])
            ent = self._graphs[key] = dict(graph=graph, **st)
        ent["x"].copy_(latents), ent["txt"].copy_(txt), ent["pooled"].copy_(pooled)
        ent["graph"].replay()
        return ent["x"].clone()


class FluxTeacherEngine(ArcFluxEngineModel):
    """Stock FLUX velocity network TIED to a student's frozen trunk (lakonlab/models/base_diffusion.py:93-94,
    lakonlab/utils/misc.py:116-132: the teacher shares storage with the student's LoRA base layers): it borrows the
    student's packed weights ([W | lora_B] buffers included — the extra K columns are never read) and adds only its own
    `norm_out.linear` and `proj_out`. Forward = lakonlab/models/architecture/diffusers/flux.py:122-156."""

    def __init__(self, student: ArcFluxEngineModel, teacher_sd: Dict[str, torch.Tensor]):
        cfg = student.cfg
        dev = student.device
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
        weights = SimpleNamespace(struct=w, keep=keep, head_n=cfg.out_channels)
        md = _lib.ModelDesc(
            arch=_lib.AFB_ARCH_FLUX, num_double=cfg.num_layers, num_single=cfg.num_single_layers,
            dim=cfg.inner_dim, heads=cfg.num_attention_heads, mlp_dim=cfg.mlp_dim, in_channels=cfg.in_channels,
            txt_dim=cfg.joint_attention_dim, pooled_dim=cfg.pooled_projection_dim, guidance=int(cfg.guidance_embeds),
            num_gaussians=cfg.num_gaussians, lora_rank=cfg.lora_rank, head_mode=1, ignore_lora=1)
        EngineModelBase.__init__(self, cfg, weights, md, dev)

    def velocity(self, latents, txt, pooled, sigma, guidance_scale, grid_hw) -> torch.Tensor:
        """u(x_t, t) as bf16 [batch, tokens, 64] (packed-token layout)."""
        return self.forward_heads(latents, txt, pooled, sigma, guidance_scale, grid_hw)

    def denoise(self, *a, **k):
        raise AfbError("the teacher has no ArcFlow heads; denoise() is a student method")
