"""ctypes binding of libvlb200.so (the C ABI in include/vlb200.h).  Fails loudly when absent."""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvlb200.so")

_lib = None

# name -> (restype, argtypes); must list every symbol include/vlb200.h declares
SIGNATURES = {
    "vlb200_abi_version": (c_int, []),
    "vlb200_last_error": (c_char_p, []),
    "vlb200_launch_count": (c_uint64, []),
    "vlb200_set_attn_fwd_variant": (c_int, [c_int]),
    "vlb200_init_uniform": (c_int, [c_void_p, c_int, c_uint64, c_uint32, c_float, c_float, c_void_p]),
    "vlb200_perturb_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_float, c_float, c_void_p]),
    "vlb200_gemm_bf16": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                 c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vlb200_gemm_bf16_ex": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int,
                                    c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_int,
                                    c_int, c_int, c_void_p]),
    "vlb200_gemm_swiglu_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                        c_int, c_void_p]),
    "vlb200_gemm_swiglu_bwd_bf16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p,
                                            c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vlb200_set_gemm_mode": (c_int, [c_int]),
    "vlb200_set_gemm_raster_mb": (c_int, [c_double]),
    "vlb200_set_gemm_raster_policy": (c_int, [c_int]),
    "vlb200_gemm_plan_raster": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vlb200_gemm_tile_coords": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "vlb200_logps_fwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlb200_logps_bwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_int, c_int, c_void_p, c_int64, c_void_p]),
    "vlb200_dpo_loss": (c_int, [c_void_p, c_void_p, c_int, c_float, c_float, c_int, c_int, c_float, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlb200_rmsnorm_fwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_int, c_float, c_void_p]),
    "vlb200_norm_bwd_workspace_floats": (c_int, [c_int]),
    "vlb200_rmsnorm_bwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                   c_void_p, c_int, c_int, c_void_p]),
    "vlb200_layernorm_fwd": (c_int, [c_void_p, c_int, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_float,
                                     c_void_p]),
    "vlb200_colsum": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "vlb200_colsum_f32": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "vlb200_dot_f32": (c_int, [c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "vlb200_rope": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vlb200_swiglu_fwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "vlb200_swiglu_bwd": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "vlb200_gelu_fwd": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "vlb200_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "vlb200_clip_im2col": (c_int, [c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "vlb200_clip_cls_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vlb200_copy_rows": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_int64, c_int, c_int, c_int,
                                 c_void_p]),
    "vlb200_gather_rows": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "vlb200_scatter_rows": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "vlb200_scatter_add_rows": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_float, c_void_p]),
    "vlb200_memset_zero": (c_int, [c_void_p, c_uint64, c_void_p]),
    "vlb200_llava_merge_index": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlb200_llava_merge_embed": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "vlb200_llava_merge_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_void_p]),
    "vlb200_llava_merge_bwd_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                            c_int, c_int, c_void_p]),
    "vlb200_llavanext_merge_index": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                             c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "vlb200_llavanext_merge_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                           c_void_p]),
    "vlb200_qwen_merge_index": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p]),
    "vlb200_attn_fwd_tc": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                   c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vlb200_attn_delta": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "vlb200_attn_bwd_tc": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                   c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                   c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "vlb200_attn_fwd_tc_varlen": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                          c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                          c_void_p]),
    "vlb200_attn_delta_varlen": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int,
                                         c_int, c_void_p]),
    "vlb200_attn_bwd_tc_varlen": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                          c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                          c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                          c_void_p]),
    "vlb200_attn_fwd_tc_ctx": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                       c_void_p]),
    "vlb200_attn_bwd_tc_ctx": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                       c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_float, c_void_p]),
    "vlb200_share_prefix_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64,
                                         c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p, c_int64,
                                         c_void_p]),
    "vlb200_pack_merge_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                       c_void_p, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "vlb200_sumsq_bf16": (c_int, [c_void_p, c_uint64, c_void_p, c_void_p, c_int, c_void_p]),
    "vlb200_adamw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_float, c_float, c_float,
                             c_float, c_float, c_int, c_float, c_void_p, c_float, c_void_p]),
    "vlb200_cast_f32_to_bf16": (c_int, [c_void_p, c_void_p, c_uint64, c_float, c_void_p]),
    "vlb200_cast_bf16_to_f32": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "vlb200_host_matching_blocks": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int]),
    "vlb200_host_ddpo_row_weights": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int64,
                                             c_int, c_void_p]),
    "vlb200_clip_preprocess_workspace_bytes": (c_size_t, [c_int, c_int]),
    "vlb200_clip_preprocess_u8": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int,
                                          c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_double,
                                          c_void_p, c_void_p, c_int, c_void_p]),
}


class Vlb200Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is the product and there is no fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    try:  # make torch's libcudart / libnccl the ones in the process
        import torch  # noqa: F401
    except Exception:
        pass
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int):
    """Map C status codes to the reference's exception types (ValueError for bad arguments)."""
    if rc == 0:
        return
    msg = load().vlb200_last_error().decode("utf-8", "replace")
    if rc == 1:
        raise ValueError(msg)
    raise Vlb200Error(f"vlb200 error {rc}: {msg}")
