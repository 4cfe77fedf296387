"""ctypes binding of libturboae_b200.so (the C ABI declared in include/turboae_b200.h)."""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TURBOAE_B200_LIB") or os.path.join(_PKG, "lib", "libturboae_b200.so")   # env: A/B builds

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_F16X3 = 2
PRECISIONS = {"fp32": PRECISION_FP32, "bf16": PRECISION_BF16, "f16x3": PRECISION_F16X3}


class TaeDecConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("block_len", "num_iteration", "num_iter_ft", "num_layer", "num_unit",
                                         "kernel_size", "extrinsic")]


class TaeEncConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("block_len", "num_layer", "num_unit", "kernel_size")]


class TaeWgradJob(C.Structure):
    _fields_ = [("a_img", C.c_void_p), ("b_img", C.c_void_p), ("grad", C.c_void_p), ("bias_grad", C.c_void_p)] + \
               [(n, C.c_int32) for n in ("b_chunks", "b_c0", "b_nc", "taps", "n_cols", "m_valid", "n_valid", "n0",
                                         "s_m", "s_n", "s_t", "g0", "g1", "tap_shift")]


IMG_CHUNK_BYTES = 8256
IMG_CHUNKS = 13


class TaeError(RuntimeError):
    pass


_P = C.c_void_p
_SIGNATURES = {
    "tae_version": (C.c_int, []),
    "tae_last_error": (C.c_char_p, []),
    "tae_launch_count": (C.c_uint64, []),
    "tae_interleave_f32": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tae_conv1d_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "tae_conv1d_elu_f32": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, _P, C.c_size_t, _P]),
    "tae_conv1d_bwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "tae_conv1d_elu_bwd_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         _P, C.c_size_t, _P]),
    "tae_dec_param_count": (C.c_size_t, [C.POINTER(TaeDecConfig)]),
    "tae_dec_packed_bytes": (C.c_size_t, [C.POINTER(TaeDecConfig)]),
    "tae_dec_pack_bf16": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P]),
    "tae_dec_workspace_bytes": (C.c_size_t, [C.POINTER(TaeDecConfig), C.c_int32, C.c_int32]),
    "tae_dec_forward": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P,
                                  C.c_size_t, _P]),
    "tae_dec_host_workspace_bytes": (C.c_size_t, [C.POINTER(TaeDecConfig), C.c_int32, C.c_int32]),
    "tae_dec_forward_host": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, C.c_size_t, _P]),
    "tae_enc_param_count": (C.c_size_t, [C.POINTER(TaeEncConfig)]),
    "tae_enc_workspace_bytes": (C.c_size_t, [C.POINTER(TaeEncConfig), C.c_int32]),
    "tae_enc_forward": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P, _P, _P, C.c_int32, _P, C.c_size_t, _P]),
    "tae_enc_packed_bytes": (C.c_size_t, [C.POINTER(TaeEncConfig)]),
    "tae_enc_pack_bf16": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P]),
    "tae_enc_forward_bf16": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P, _P, _P, _P, C.c_int32, _P, C.c_size_t, _P]),
    "tae_dec_packed_bytes_x3": (C.c_size_t, [C.POINTER(TaeDecConfig)]),
    "tae_dec_pack_f16x3": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P]),
    "tae_enc_packed_bytes_x3": (C.c_size_t, [C.POINTER(TaeEncConfig)]),
    "tae_enc_pack_f16x3": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P]),
    "tae_enc_forward_f16x3": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, C.c_size_t, _P]),
    "tae_power_norm_f32": (C.c_int, [_P, _P, C.c_size_t, _P, _P, _P]),
    "tae_power_norm_given_f32": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_float, C.c_float, _P]),
    "tae_power_norm_ste_f32": (C.c_int, [_P, _P, C.c_size_t, _P, _P, C.c_float, C.c_float, _P]),
    "tae_dec_out_backward_f32": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, _P]),
    "tae_dec_input_grad_f32": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tae_enc_out_backward_f32": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P]),
    "tae_power_stats_f32": (C.c_int, [_P, C.c_size_t, _P, _P]),
    "tae_power_norm_bwd_sums_f32": (C.c_int, [_P, _P, C.c_size_t, _P, _P]),
    "tae_power_norm_bwd_f32": (C.c_int, [_P, _P, _P, C.c_size_t, _P, _P, _P, _P]),
    "tae_gru_direction_f32": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tae_gru_direction_bwd_f32": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tae_gru_rows_per_block": (C.c_int32, [C.c_int32]),
    "tae_gru_tile_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "tae_gru_packed_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "tae_gru_pack_bf16": (C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tae_gru_direction_bf16": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_int32, _P, C.c_size_t, _P]),
    "tae_gru_tiles_from_f32": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tae_gru_linear_f32": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "tae_train_groups": (C.c_int32, [C.c_int32, C.c_int32]),
    "tae_train_units": (C.c_int32, [C.c_int32, C.c_int32]),
    "tae_dec_forward_train_bf16": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "tae_dec_bwd_packed_bytes": (C.c_size_t, [C.POINTER(TaeDecConfig)]),
    "tae_dec_pack_bwd_bf16": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P]),
    "tae_dec_backward_bf16": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, C.c_size_t, _P]),
    "tae_dec_backward_range_bf16": (C.c_int, [C.POINTER(TaeDecConfig), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P,
                                              C.c_size_t, _P]),
    "tae_enc_forward_train_bf16": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P, _P, _P, _P, C.c_int32, _P, _P, _P, C.c_size_t, _P]),
    "tae_enc_bwd_packed_bytes": (C.c_size_t, [C.POINTER(TaeEncConfig)]),
    "tae_enc_pack_bwd_bf16": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P]),
    "tae_enc_backward_bf16": (C.c_int, [C.POINTER(TaeEncConfig), _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P, C.c_size_t, _P]),
    "tae_wgrad_bf16": (C.c_int, [C.POINTER(TaeWgradJob), C.c_int32, _P, _P, C.c_size_t, _P]),
    "tae_awgn_f32": (C.c_int, [_P, _P, C.c_size_t, C.c_float, C.c_uint64, C.c_uint64, _P]),
    "tae_error_count_f32": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    # debug / self-test entry points
    "tae_debug_set_timeline": (None, [_P]),
    "tae_debug_gru_timeline": (None, [_P]),
    "tae_debug_probe_rate": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, C.c_int32, _P]),
    "tae_debug_probe_lbo": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "tae_debug_probe_pair": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
}
# symbols every build must export (checked by tests/test_cabi.py against include/turboae_b200.h)
PUBLIC_SYMBOLS = [s for s in _SIGNATURES if not s.startswith("tae_debug_")]

_lock = threading.Lock()
_lib = None


def load():
    """Load the shared library (building it first if nvcc is present and it is stale/missing).
    Raises loudly when it cannot be had -- the product has no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH) or os.environ.get("TURBOAE_B200_REBUILD"):
            from . import build as _build
            _build.build(force=True)
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise TaeError("cannot load %s (%s); run `python -m turboae_b200.build` -- there is no CPU fallback"
                           % (LIB_PATH, e)) from e
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise TaeError("libturboae_b200 error %d: %s" % (rc, load().tae_last_error().decode()))


def ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise TaeError("%s must live on a CUDA device (got %s): turboae_b200 has no CPU fallback" % (what, t.device))


def launch_count() -> int:
    return int(load().tae_launch_count())
