"""ctypes binding of libarcflow_b200.so — mirrors include/arcflow_b200.h one to one.

There is NO fallback: if the shared library is missing or an entry point fails, the caller gets an
exception (`AfbError`), never a silent PyTorch path.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libarcflow_b200.so"

AFB_OK = 0
AFB_EPI_BIAS, AFB_EPI_BIAS_GELU, AFB_EPI_BIAS_GATE_RES, AFB_EPI_BIAS_RES, AFB_EPI_BIAS_QKNORM_ROPE = 0, 1, 2, 3, 4
AFB_SL_SILU_IN, AFB_SL_ACCUMULATE = 1, 2
AFB_ARCH_FLUX, AFB_ARCH_QWEN = 0, 1
AFB_ABI_VERSION = 3


class AfbError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", C.c_void_p * 3),
        ("a_ld", C.c_int64 * 3),
        ("a_batch_stride", C.c_int64 * 3),
        ("a_k", C.c_int32 * 3),
        ("batches", C.c_int32),
        ("rows_per_batch", C.c_int32),
        ("w", C.c_void_p),
        ("w_ld", C.c_int64),
        ("n", C.c_int32),
        ("epilogue", C.c_int32),
        ("out", C.c_void_p),
        ("out_ld", C.c_int64),
        ("out_batch_stride", C.c_int64),
        ("bias", C.c_void_p),
        ("gate", C.c_void_p),
        ("gate_batch_stride", C.c_int64),
        ("res", C.c_void_p),
        ("res_ld", C.c_int64),
        ("res_batch_stride", C.c_int64),
        ("w_transposed", C.c_int32),
        ("w_k", C.c_int32),
        ("w2", C.c_void_p),
        ("w2_ld", C.c_int64),
        ("alpha", C.c_float),
        ("norm_q", C.c_void_p), ("norm_k", C.c_void_p), ("rope", C.c_void_p),
        ("rope_row0", C.c_int32), ("qk_cols", C.c_int32), ("norm_eps", C.c_float),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("o", C.c_void_p),
        ("q_ld", C.c_int64), ("k_ld", C.c_int64), ("v_ld", C.c_int64), ("o_ld", C.c_int64),
        ("q_batch_stride", C.c_int64), ("k_batch_stride", C.c_int64),
        ("v_batch_stride", C.c_int64), ("o_batch_stride", C.c_int64),
        ("batch", C.c_int32), ("seq", C.c_int32), ("heads", C.c_int32),
        ("scale", C.c_float), ("lse", C.c_void_p), ("score_bound", C.c_float),
    ]


class AttnBwdDesc(C.Structure):
    _fields_ = [("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p), ("qkv_ld", C.c_int64), ("qkv_batch_stride", C.c_int64),
                ("o", C.c_void_p), ("d_o", C.c_void_p), ("o_ld", C.c_int64), ("o_batch_stride", C.c_int64),
                ("lse", C.c_void_p), ("delta_ws", C.c_void_p), ("dq", C.c_void_p), ("dk", C.c_void_p), ("dv", C.c_void_p),
                ("dqkv_ld", C.c_int64), ("dqkv_batch_stride", C.c_int64),
                ("batch", C.c_int32), ("seq", C.c_int32), ("heads", C.c_int32), ("scale", C.c_float)]


class ModelDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "arch", "num_double", "num_single", "dim", "heads", "mlp_dim", "in_channels", "txt_dim",
        "pooled_dim", "guidance", "num_gaussians", "lora_rank", "head_mode", "ignore_lora")]


_P = C.c_void_p


class DoubleBlock(C.Structure):
    _fields_ = [(n, _P) for n in (
        "img_qkv_w", "img_qkv_b", "img_nq", "img_nk", "img_out_w", "img_out_b",
        "img_up_w", "img_up_b", "img_up_la", "img_down_w", "img_down_b", "img_down_la",
        "txt_qkv_w", "txt_qkv_b", "txt_nq", "txt_nk", "txt_out_w", "txt_out_b",
        "txt_up_w", "txt_up_b", "txt_up_la", "txt_down_w", "txt_down_b", "txt_down_la")] + [
        ("img_mod_off", C.c_int64), ("txt_mod_off", C.c_int64), ("qk_bound", C.c_float), ("reserved0", C.c_float)]


class SingleBlock(C.Structure):
    _fields_ = [(n, _P) for n in (
        "qkv_w", "qkv_b", "nq", "nk", "mlp_w", "mlp_b", "mlp_la", "out_w", "out_b", "out_la")] + [
        ("mod_off", C.c_int64), ("qk_bound", C.c_float), ("reserved0", C.c_float)]


class Weights(C.Structure):
    _fields_ = [(n, _P) for n in (
        "x_emb_w", "x_emb_b", "ctx_w", "ctx_b", "txt_norm_w",
        "t1_w", "t1_b", "t1_la", "t1_lb", "t2_w", "t2_b", "t2_la", "t2_lb",
        "g1_w", "g1_b", "g2_w", "g2_b", "p1_w", "p1_b", "p2_w", "p2_b",
        "mod_w", "mod_b")] + [
        ("mod_total", C.c_int64), ("norm_out_mod_off", C.c_int64),
        ("head_w", _P), ("head_b", _P), ("alt_norm_out_w", _P), ("alt_norm_out_b", _P), ("head_n", C.c_int32),
        ("dbl", C.POINTER(DoubleBlock)), ("sgl", C.POINTER(SingleBlock))]


class ForwardArgs(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("txt_len", C.c_int32), ("img_len", C.c_int32),
        ("latents", _P), ("txt", _P), ("pooled", _P),
        ("timestep", _P), ("guidance", _P), ("rope_cos", _P), ("rope_sin", _P),
        ("head_out", _P),
    ]


class DenoiseArgs(C.Structure):
    _fields_ = [
        ("fwd", ForwardArgs),
        ("nfe", C.c_int32),
        ("sigmas", C.POINTER(C.c_float)),
        ("timesteps", C.POINTER(C.c_float)),
        ("x", _P),
        ("eps", C.c_float),
    ]


class PolicyArgs(C.Structure):
    _fields_ = [("head", _P), ("head_ld", C.c_int64), ("batch", C.c_int32), ("tokens", C.c_int32),
                ("num_gaussians", C.c_int32), ("mode", C.c_int32),
                ("sigma_src", C.POINTER(C.c_float)), ("sigma_start", C.POINTER(C.c_float)),
                ("sigma_end", C.POINTER(C.c_float)), ("drop_mask", C.POINTER(C.c_uint8)),
                ("small", C.POINTER(C.c_uint8)), ("x_in", _P), ("out", _P), ("out_bf16", _P), ("eps", C.c_float)]


AFB_POLICY_INTEGRATE, AFB_POLICY_VELOCITY, AFB_POLICY_AVERAGE_U = 0, 1, 2


class AdamwArgs(C.Structure):
    _fields_ = [("params", _P), ("grads", _P), ("exp_avg", _P), ("exp_avg_sq", _P), ("ema", _P), ("bf16_shadow", _P),
                ("n", C.c_int64), ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("step", C.c_int32), ("max_norm", C.c_float), ("grad_norm_sq", _P),
                ("skipped", _P), ("ema_momentum", C.c_float), ("ema_copy", C.c_int32),
                ("lr_mult_begin", C.c_int64), ("lr_mult_end", C.c_int64), ("lr_mult", C.c_float), ("skip_norm", C.c_float)]


class Adamw8bitArgs(C.Structure):
    _fields_ = [("base", AdamwArgs), ("state1", _P), ("state2", _P), ("absmax1", _P), ("absmax2", _P), ("qmap1", _P),
                ("qmap2", _P), ("blocksize", C.c_int32), ("reserved0", C.c_int32)]


class DoubleBlockGrads(C.Structure):
    _fields_ = [(n, _P) for n in ("img_up_la", "img_up_lb", "img_down_la", "img_down_lb",
                                  "txt_up_la", "txt_up_lb", "txt_down_la", "txt_down_lb")]


class SingleBlockGrads(C.Structure):
    _fields_ = [(n, _P) for n in ("mlp_la", "mlp_lb", "out_la", "out_lb")]


class BackwardArgs(C.Structure):
    _fields_ = [("fwd", ForwardArgs), ("d_head_in", _P), ("dbl", C.POINTER(DoubleBlockGrads)),
                ("sgl", C.POINTER(SingleBlockGrads)), ("d_mod", _P)]


class EmbedGrads(C.Structure):
    _fields_ = [(n, _P) for n in ("t1_la", "t1_lb", "t2_la", "t2_lb")]


class ConvDesc(C.Structure):
    _fields_ = [("x", _P), ("w", _P), ("bias", _P), ("res", _P), ("out", _P),
                ("x_ld", C.c_int64), ("out_ld", C.c_int64), ("res_ld", C.c_int64),
                ("n", C.c_int32), ("h", C.c_int32), ("w_px", C.c_int32), ("c_in", C.c_int32), ("c_out", C.c_int32),
                ("epilogue", C.c_int32)]


class Profile(C.Structure):
    _fields_ = [("gemm_ms", C.c_double), ("attn_ms", C.c_double), ("gemm_flops", C.c_double),
                ("attn_flops", C.c_double), ("gemm_launches", C.c_int64), ("attn_launches", C.c_int64),
                ("gemm_fused_qk_ms", C.c_double), ("gemm_fused_qk_flops", C.c_double), ("gemm_fused_qk_launches", C.c_int64)]


# name -> (restype, argtypes); every symbol include/arcflow_b200.h declares
SIGNATURES = {
    "afb_abi_version": (C.c_int, []),
    "afb_last_error": (C.c_char_p, []),
    "afb_launch_count": (C.c_uint64, []),
    "afb_gemm": (C.c_int, [C.POINTER(GemmDesc), _P]),
    "afb_attention": (C.c_int, [C.POINTER(AttnDesc), _P]),
    "afb_attention_backward": (C.c_int, [C.POINTER(AttnBwdDesc), _P]),
    "afb_debug_attention_trace": (C.c_int, [C.POINTER(C.c_int64), C.c_int32]),
    "afb_ln_modulate": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.c_int64, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_float, _P]),
    "afb_rmsnorm_rope": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P,
                                   C.c_float, _P]),
    "afb_small_linear": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.c_int64, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, _P]),
    "afb_timestep_embed": (C.c_int, [_P, _P, C.c_int32, _P]),
    "afb_sampler_step": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_int64, C.c_int32, C.c_float,
                                   C.c_float, C.c_float, C.c_float, _P]),
    "afb_cast_f32_bf16": (C.c_int, [_P, _P, C.c_int64, _P]),
    "afb_policy_eval": (C.c_int, [C.POINTER(PolicyArgs), _P]),
    "afb_policy_backward": (C.c_int, [C.POINTER(PolicyArgs), _P, _P, C.c_int64, C.c_float, C.c_int32, C.c_int32, _P]),
    "afb_dropout_rows": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_uint64, C.c_uint32, C.c_float, C.c_int32, C.c_int32, _P]),
    "afb_cfg_combine": (C.c_int, [_P, _P, C.c_int64, C.c_float, _P]),
    "afb_colsum_f32": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, _P]),
    "afb_gemm_tn": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, _P]),
    "afb_ln_mod_param_grad": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_float, _P]),
    "afb_rowlinear_param_grad": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, _P]),
    "afb_engine_export": (C.c_int, [_P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "afb_ln_modulate_bwd": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_float, C.c_int32, _P]),
    "afb_rowscale": (C.c_int, [_P, C.c_int64, C.c_int64, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                               C.c_int32, _P]),
    "afb_gelu_bwd": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, _P]),
    "afb_rmsnorm_rope_bwd": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, _P, _P, _P, _P, _P, _P, C.c_float, _P]),
    "afb_engine_set_lora_dropout": (C.c_int, [_P, C.c_float, C.c_uint64]),
    "afb_engine_set_ignore_lora": (C.c_int, [_P, C.c_int32]),
    "afb_engine_train_reserve": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "afb_engine_stash_bytes": (C.c_int64, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "afb_engine_set_activation_stash": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "afb_engine_forward_train": (C.c_int, [_P, C.POINTER(ForwardArgs), _P]),
    "afb_engine_backward": (C.c_int, [_P, C.POINTER(BackwardArgs), _P]),
    "afb_engine_backward_embed": (C.c_int, [_P, C.POINTER(ForwardArgs), _P, C.POINTER(EmbedGrads), _P]),
    "afb_rope_pack": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "afb_conv3x3": (C.c_int, [C.POINTER(ConvDesc), _P]),
    "afb_groupnorm_ws_floats": (C.c_int, [C.c_int32, C.c_int64]),
    "afb_groupnorm": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_float, C.c_int32, _P]),
    "afb_upsample2x": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "afb_softmax_rows": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, _P]),
    "afb_vae_pre": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P]),
    "afb_vae_post": (C.c_int, [_P, C.c_int64, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "afb_grad_norm_sq": (C.c_int, [_P, C.c_int64, _P, _P]),
    "afb_grad_norm_scratch_floats": (C.c_int, []),
    "afb_grad_norm_sq_ws": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int64, _P]),
    "afb_adamw_ema_step": (C.c_int, [C.POINTER(AdamwArgs), _P]),
    "afb_adamw8bit_ema_step": (C.c_int, [C.POINTER(Adamw8bitArgs), _P]),
    "afb_axpy_rows": (C.c_int, [_P, _P, C.POINTER(C.c_float), _P, _P, C.c_int32, C.c_int64, C.c_int32, _P]),
    "afb_mse_rows": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int32, _P]),
    "afb_engine_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(_P)]),
    "afb_engine_destroy": (None, [_P]),
    "afb_engine_bind": (C.c_int, [_P, C.POINTER(Weights)]),
    "afb_engine_set_lora_scale": (C.c_int, [_P, C.c_float]),
    "afb_engine_workspace_bytes": (C.c_size_t, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "afb_engine_reserve": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "afb_engine_forward": (C.c_int, [_P, C.POINTER(ForwardArgs), _P]),
    "afb_engine_denoise": (C.c_int, [_P, C.POINTER(DenoiseArgs), _P]),
    "afb_engine_set_profiling": (C.c_int, [_P, C.c_int32]),
    "afb_engine_read_profile": (C.c_int, [_P, C.POINTER(Profile)]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the shared library (building is `__graft_entry__.build()`'s job). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise AfbError(
            f"{LIB_PATH} not found — run `python -m arcflow_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: that is the point
        fn.restype = res
        fn.argtypes = args
    if lib.afb_abi_version() != AFB_ABI_VERSION:
        raise AfbError("libarcflow_b200.so ABI version mismatch — rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != AFB_OK:
        msg = load().afb_last_error().decode(errors="replace")
        raise AfbError(f"{what} failed (code {rc}): {msg}")
