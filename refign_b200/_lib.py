"""ctypes binding of librefign_b200.so -- the C-ABI boundary (include/refign_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a
RuntimeError is raised.  The library is built in-tree by ``refign_b200.build``
(``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librefign_b200.so")
_lib = None

c_int, c_i64, c_u64, c_f32, c_p = (ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float,
                                   ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/refign_b200.h one to one
_SIGNATURES = {
    "rf_last_error": (ctypes.c_char_p, []),
    "rf_version": (c_int, []),
    "rf_device_check": (c_int, []),
    "rf_set_device": (c_int, [c_int]),
    "rf_local_corr_fwd": (c_int, [c_p, c_p, c_p, c_p] + [c_int] * 16 + [c_int, c_p]),
    "rf_local_corr_bwd_scratch_bytes": (c_i64, [c_int] * 16),
    "rf_local_corr_bwd": (c_int, [c_p] * 6 + [c_int] * 16 + [c_p]),
    "rf_relu_l2norm_bwd": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_i64, c_p]),
    "rf_global_corr_workspace_bytes": (c_i64, [c_int, c_i64, c_i64]),
    "rf_global_corr_fwd": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_i64, c_i64, c_int, c_int, c_p]),
    "rf_flow_is_zero": (c_int, [c_p, c_i64, c_p, c_p]),
    "rf_warp_bilinear_fwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p]),
    "rf_warp_bilinear_bwd": (c_int, [c_p] * 6 + [c_int] * 4 + [c_p]),
    "rf_cert_fwd": (c_int, [c_p, c_p, c_i64, c_p]),
    "rf_refine_fwd": (c_int, [c_p] * 10 + [c_int, c_int, c_i64, c_f32, c_u64, c_int, c_p]),
    "rf_dwconv3x3_nhwc_fwd": (c_int, [c_p, c_p, c_p, c_p] + [c_int] * 7 + [c_p]),
    "rf_dwconv3x3_nhwc_bwd_input": (c_int, [c_p, c_p, c_p] + [c_int] * 6 + [c_p]),
    "rf_dwconv3x3_gelu_bwd_pre": (c_int, [c_p] * 5 + [c_int] * 6 + [c_p]),
    "rf_dwconv3x3_nhwc_bwd_weight": (c_int, [c_p] * 4 + [c_int] * 7 + [c_p]),
    "rf_patch_embed_ln_fwd": (c_int, [c_p] * 9 + [c_int, c_int, c_int, c_int, c_f32, c_p]),
    "rf_add_layernorm_fwd": (c_int, [c_p] * 9 + [c_i64, c_int, c_i64, c_f32, c_int, c_int, c_int, c_p]),
    "rf_add_layernorm_bwd": (c_int, [c_p] * 11 + [c_i64, c_int, c_i64, c_int, c_int, c_int, c_int, c_p]),
    "rf_sr_attention_fwd": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_f32, c_p]),
    "rf_sr_attention_bwd_workspace_bytes": (c_i64, [c_int, c_int, c_int, c_int]),
    "rf_sr_attention_bwd": (c_int, [c_p] * 8 + [c_int, c_int, c_int, c_int, c_f32, c_p]),
    "rf_sr_attention_f32_fwd": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_f32, c_p]),
    "rf_sr_attention_f32_bwd_workspace_bytes": (c_i64, [c_int, c_int, c_int]),
    "rf_sr_attention_f32_bwd": (c_int, [c_p] * 8 + [c_int, c_int, c_int, c_int, c_int, c_f32, c_p]),
    "rf_gemm_bf16": (c_int, [c_p, c_p, c_p, c_p] + [c_int] * 7 + [c_p, c_p]),
    "rf_conv3x3_bf16": (c_int, [c_p, c_p, c_p, c_p] + [c_int] * 7 + [c_f32, c_int, c_p]),
    "rf_uncertainty_cnn_param_bytes": (c_int, []),
    "rf_uncertainty_cnn_fwd": (c_int, [c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_f32, c_p]),
    "rf_maxpool2x2_nhwc_bf16": (c_int, [c_p, c_p, c_int, c_int, c_int, c_int, c_p]),
    "rf_conv3x3_wgrad_bf16": (c_int, [c_p, c_p, c_p] + [c_int] * 5 + [c_p]),
    "rf_dacs_count": (c_int, [c_p, c_i64, c_f32, c_p, c_p]),
    "rf_dacs_mix": (c_int, [c_p] * 10 + [c_int] * 5 + [c_p]),
    "rf_dacs_blur": (c_int, [c_p, c_p, c_p, c_int, c_int, c_int, c_p]),
    "rf_ema_update": (c_int, [c_p, c_p, c_i64, ctypes.c_double, c_p]),
    "rf_bn_stats": (c_int, [c_p, c_p, c_i64, c_int, c_int, c_p]),
    "rf_bn_finalize": (c_int, [c_p] * 9 + [c_int, ctypes.c_double, c_f32, c_f32, c_p]),
    "rf_bn_apply": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_int, c_int, c_int, c_p]),
    "rf_bn_bwd_reduce": (c_int, [c_p] * 7 + [c_i64, c_int, c_int, c_int, c_p]),
    "rf_bn_bwd_apply": (c_int, [c_p] * 8 + [c_i64, c_int, ctypes.c_double, c_int, c_int, c_p]),
    "rf_colsum": (c_int, [c_p, c_p, c_i64, c_int, c_int, c_int, c_p]),
    "rf_cast_bf16": (c_int, [c_p, c_p, c_i64, c_p]),
    "rf_upsample_concat_fwd": (c_int, [c_p, c_p, c_p, c_p, c_int, c_p, c_int, c_int, c_int, c_p]),
    "rf_upsample_concat_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_p]),
    "rf_upsample_ce_fwd": (c_int, [c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "rf_upsample_ce_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "rf_upsample_bilinear_f32": (c_int, [c_p, c_p, c_i64, c_int, c_int, c_int, c_int, c_p]),
    "rf_bias_act": (c_int, [c_p, c_p, c_i64, c_int, c_i64, c_int, c_f32, c_int, c_p]),
    "rf_space_to_depth": (c_int, [c_p, c_p, c_int, c_int, c_int, c_int, c_int, c_int, c_p]),
    "rf_ema_update_dev": (c_int, [c_p, c_p, c_i64, c_p, c_p]),
    "rf_adamw_step_dev": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_int, ctypes.POINTER(c_i64), ctypes.POINTER(c_f32),
                                  c_f32, c_f32, c_f32, c_f32, c_p, c_p]),
    "rf_adamw_step": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_int, ctypes.POINTER(c_i64),
                              ctypes.POINTER(c_f32), ctypes.POINTER(c_f32), c_f32, c_f32, c_f32,
                              c_int, c_f32, c_p]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "refign_b200: %s not found -- run `python -m refign_b200.build` "
                "(there is no CPU / PyTorch fallback for the CUDA path)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().rf_last_error().decode("utf-8", "replace")
        raise RuntimeError("refign_b200.%s failed (rc=%d): %s" % (what, rc, msg))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("refign_b200: expected CUDA tensors (got %s); this path has no CPU "
                               "implementation" % t.device)
