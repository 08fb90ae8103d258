"""ctypes binding of libbsi_b200.so (the C ABI declared in include/bsi_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, the caller
gets an exception carrying bsi_last_error().
"""

from __future__ import annotations

import ctypes as C
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
# BSI_B200_LIB selects an experimental build of the same library (kernel A/B runs: `python -m bsi_b200.build --variant NAME -DFLAG`
# writes libbsi_b200_NAME.so next to the product library); unset = the product library
LIB_PATH = os.environ.get("BSI_B200_LIB") or os.path.join(_PKG, "libbsi_b200.so")


class RowRef(C.Structure):
    _fields_ = [("base", C.c_void_p), ("sample_stride", C.c_int32), ("step_stride", C.c_int32)]


class Noise(C.Structure):
    _fields_ = [("eps", C.c_void_p), ("seed", C.c_uint64), ("sample_base", C.c_uint64), ("draw", C.c_int32), ("key_ptr", C.c_void_p)]


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("C", C.c_void_p), ("bias", C.c_void_p),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("lda", C.c_int32), ("ldw", C.c_int32), ("ldc", C.c_int32),
        ("batch", C.c_int32),
        ("stride_a", C.c_int64), ("stride_w", C.c_int64), ("stride_c", C.c_int64), ("stride_bias", C.c_int64),
        ("epilogue", C.c_int32),
        ("gate", RowRef),
        ("step_ptr", C.c_void_p),
        ("rows_per_sample", C.c_int32),
        ("pos", C.c_void_p),
        ("patch", C.c_int32), ("grid_w", C.c_int32), ("channels", C.c_int32),
        ("aux", C.c_void_p),
    ]  # fmt: skip


class ConvArgs(C.Structure):
    _fields_ = [
        ("X1", C.c_void_p), ("X2", C.c_void_p), ("W", C.c_void_p), ("Y", C.c_void_p), ("bias", C.c_void_p), ("resid", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("Wd", C.c_int32), ("C1", C.c_int32), ("C2", C.c_int32), ("N", C.c_int32),
        ("taps", C.c_int32), ("ldc", C.c_int32), ("epilogue", C.c_int32),
        ("scale", RowRef), ("shift", RowRef), ("step_ptr", C.c_void_p), ("gn_partial", C.c_void_p),
    ]  # fmt: skip


class DitConfig(C.Structure):
    _fields_ = [
        ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
        ("patch", C.c_int32), ("dim", C.c_int32), ("depth", C.c_int32), ("heads", C.c_int32),
        ("fourier_n_min", C.c_int32), ("fourier_n_max", C.c_int32), ("exact", C.c_int32),
    ]  # fmt: skip


class UnetConfig(C.Structure):
    _fields_ = [
        ("channels", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("dim", C.c_int32), ("levels", C.c_int32), ("heads", C.c_int32),
        ("pos_size", C.c_int32), ("pos_mult", C.c_int32), ("fourier_n_min", C.c_int32), ("fourier_n_max", C.c_int32),
    ]  # fmt: skip


class ReduceJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("groups", C.c_int32), ("rows", C.c_int32), ("D", C.c_int32), ("dst_ld", C.c_int32),
                ("accumulate", C.c_int32), ("reserved", C.c_int32)]  # fmt: skip


class AdamWArgs(C.Structure):
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("ema", C.c_void_p),
        ("param_bf16", C.c_void_p), ("grad_sumsq", C.c_void_p), ("numel", C.c_int64), ("step", C.c_int64),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double), ("weight_decay", C.c_double),
        ("max_norm", C.c_float), ("grad_scale", C.c_float), ("ema_weight", C.c_float), ("ema_mode", C.c_int32), ("zero_grad", C.c_int32),
    ]  # fmt: skip


(EPI_BIAS_BF16, EPI_BIAS_GELU_BF16, EPI_BIAS_SILU_BF16, EPI_BIAS_F32, EPI_GATE_RESID_F32, EPI_POS_F32, EPI_UNPATCH_F32, EPI_MOD_SILU_BF16,
 EPI_BIAS_GELU_DUAL_BF16, EPI_MUL_GELU_GRAD_BF16) = range(10)

_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol of include/bsi_b200.h (checked by tests/test_abi.py)
SIGNATURES = {
    "bsi_attention_backward_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, C.c_uint32, _i32, _vp]),
    "bsi_attention_lse_bf16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "bsi_attention_dropout_bf16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, C.c_uint32, _vp]),
    "bsi_layernorm_mod_dropout_bf16": (C.c_int, [_vp, _vp, RowRef, RowRef, _i32, _i64, _i32, _f32, _f32, C.c_uint32, _vp]),
    "bsi_gate_residual": (C.c_int, [_vp, _vp, _vp, RowRef, _i32, _i64, _i32, _vp]),
    "bsi_gate_residual_layernorm_bf16": (C.c_int, [_vp, _vp, _vp, _vp, RowRef, RowRef, RowRef, _vp, _vp, _i32, _i64, _i32, _f32, _f32, C.c_uint32, _vp]),
    "bsi_gate_residual_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, RowRef, _i32, _i32, _i32, _vp]),
    "bsi_gate_residual_backward_rows": (C.c_int, [_vp, _vp, _vp, _vp, _vp, RowRef, _i32, _i32, _i64, _i32, _vp]),
    "bsi_reduce_rows": (C.c_int, [C.POINTER(ReduceJob), _i32, _vp]),
    "bsi_cast_transpose_bf16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "bsi_colsum_bf16": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _i32, _vp]),
    "bsi_gelu_bf16": (C.c_int, [_vp, _vp, _i64, _vp]),
    "bsi_gelu_backward_bf16": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "bsi_layernorm_mod_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, RowRef, _vp, _i32, _i32, _i64, _i32, _f32, _f32, C.c_uint32, _vp]),
    "bsi_gemm_wgrad_bf16": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bsi_grad_sumsq": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "bsi_grad_sumsq_workspace_floats": (_i32, []),
    "bsi_adamw_ema_step": (C.c_int, [C.POINTER(AdamWArgs), _vp]),
    "bsi_ema_update": (C.c_int, [_vp, _vp, _i64, _f32, _i32, _vp]),
    "bsi_abi_version": (C.c_int, []),
    "bsi_last_error": (C.c_char_p, []),
    "bsi_device_arch": (C.c_int, []),
    "bsi_launch_counter": (C.c_longlong, []),
    "bsi_profile_gemm_begin": (C.c_int, []),
    "bsi_profile_gemm_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_i32)]),
    "bsi_sample_init": (C.c_int, [_vp, _vp, Noise, _i64, _i64, _vp]),
    "bsi_step_fused": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, Noise, _vp, _vp, _i64, _i64, _vp]),
    "bsi_step_advance": (C.c_int, [_vp, _vp]),
    "bsi_edm_combine": (C.c_int, [_vp, _vp, _vp, RowRef, RowRef, _vp, _i64, _i64, _vp]),
    "bsi_scale_rows": (C.c_int, [_vp, _vp, RowRef, _vp, _i64, _i64, _vp]),
    "bsi_q_sample": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, Noise, _i64, _i64, _i64, _vp]),
    "bsi_bucketize": (C.c_int, [_vp, _vp, _vp, _f32, _f32, _i32, _i64, _vp]),
    "bsi_to_uint8": (C.c_int, [_vp, _vp, _f32, _f32, _i64, _vp]),
    "bsi_recon_reduce": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _i64, _i64, _i64, _vp]),
    "bsi_sqerr_reduce": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bsi_sqerr_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "bsi_gemm_bf16": (C.c_int, [C.POINTER(GemmArgs), _vp]),
    "bsi_conv_bf16": (C.c_int, [C.POINTER(ConvArgs), _vp]),
    "bsi_groupnorm_apply_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "bsi_gemm_force_cta_group": (C.c_int, [_i32]),
    "bsi_cast_bf16": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _vp]),
    "bsi_layernorm_mod_bf16": (C.c_int, [_vp, _vp, RowRef, RowRef, _vp, _vp, _vp, _i32, _i64, _i32, _f32, _vp]),
    "bsi_attention_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "bsi_split3_bf16": (C.c_int, [_vp, _vp, _i64, _i32, _i64, _i32, _i32, _i32, _vp]),
    "bsi_attention_f32": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "bsi_layernorm_mod_f32": (C.c_int, [_vp, _vp, RowRef, RowRef, _vp, _vp, _vp, _i32, _i64, _i32, _f32, _vp]),
    "bsi_dit_patch_operand_f32": (C.c_int, [_vp, _vp, RowRef, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bsi_attention_force_legacy": (C.c_int, [_i32]),
    "bsi_attention_debug_phases": (C.c_int, [_vp]),
    "bsi_attention_backward_debug_phases": (C.c_int, [_vp]),
    "bsi_dit_patch_operand": (C.c_int, [_vp, _vp, RowRef, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bsi_time_embed": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "bsi_dit_create": (C.c_int, [C.POINTER(DitConfig), C.POINTER(_vp)]),
    "bsi_dit_destroy": (None, [_vp]),
    "bsi_dit_param_bytes": (_i64, [_vp]),
    "bsi_dit_workspace_bytes": (_i64, [_vp, _i32]),
    "bsi_dit_cond_bytes": (_i64, [_vp, _i32]),
    "bsi_dit_bind_params": (C.c_int, [_vp, _vp, _i64]),
    "bsi_dit_set_param": (C.c_int, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "bsi_dit_missing_params": (C.c_int, [_vp]),
    "bsi_dit_cond_scratch_bytes": (_i64, [_vp, _i32]),
    "bsi_dit_conditioning": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "bsi_dit_forward": (C.c_int, [_vp, _vp, _vp, RowRef, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i64, _vp]),
    "bsi_dit_peek": (C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp]),
    "bsi_groupnorm_act_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "bsi_unet_input_bf16": (C.c_int, [_vp, _vp, RowRef, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bsi_unet_decode": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "bsi_pack_conv_weight": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "bsi_attention_d128_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _vp]),
    "bsi_unet_create": (C.c_int, [C.POINTER(UnetConfig), C.POINTER(_vp)]),
    "bsi_unet_destroy": (None, [_vp]),
    "bsi_unet_param_bytes": (_i64, [_vp]),
    "bsi_unet_workspace_bytes": (_i64, [_vp, _i32]),
    "bsi_unet_cond_bytes": (_i64, [_vp, _i32]),
    "bsi_unet_cond_scratch_bytes": (_i64, [_vp, _i32]),
    "bsi_unet_bind_params": (C.c_int, [_vp, _vp, _i64]),
    "bsi_unet_set_param": (C.c_int, [_vp, C.c_char_p, _vp, _i64, _vp]),
    "bsi_unet_missing_params": (C.c_int, [_vp]),
    "bsi_unet_conditioning": (C.c_int, [_vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "bsi_unet_forward": (C.c_int, [_vp, _vp, _vp, RowRef, _vp, _i32, _i32, _i32, _i32, _vp, _i32, _vp, _i64, _vp]),
}

_lib = None


class BsiNativeError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the shared library (building it is `python -m bsi_b200.build` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BsiNativeError(
                f"{LIB_PATH} is missing: build it with `python -m bsi_b200.build` — bsi_b200 has no CPU/PyTorch fallback"
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().bsi_last_error().decode(errors="replace")
        raise BsiNativeError(f"{what or 'bsi_b200 call'} failed with status {rc}: {msg}")


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "native kernels need contiguous CUDA tensors"
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def rowref(t: torch.Tensor | None, sample_stride: int = 0, step_stride: int = 0, offset: int = 0) -> RowRef:
    """Reference into a float32 CUDA tensor (see bsi_rowref in include/bsi_b200.h)."""
    if t is None:
        return RowRef(None, 0, 0)
    assert t.dtype == torch.float32
    return RowRef(ptr(t) + 4 * offset, sample_stride, step_stride)


def noise(eps: torch.Tensor | None = None, seed: int = 0, sample_base: int = 0, draw: int = 0, key: torch.Tensor | None = None) -> Noise:
    """bsi_noise: injected `eps`, or Philox keyed by (seed, sample_base) -- by value, or read on the device from `key`
    (int64[2] CUDA tensor {seed, sample_base}) so that a captured graph can be replayed with another key."""
    if eps is not None:
        assert eps.dtype == torch.float32
    if key is not None:
        assert key.dtype == torch.int64 and key.numel() == 2 and key.is_cuda
    return Noise(ptr(eps), seed & 0xFFFFFFFFFFFFFFFF, sample_base, draw, ptr(key))
