"""ctypes binding of libmmoe_b200.so (the C ABI declared in include/mmoe_b200.h).

There is no fallback: if the shared object is missing it is built with nvcc, and if
that fails the import error is raised to the caller.  Compute entry points raise
RuntimeError with the library's message on any non-zero return code.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# MMOE_B200_LIB: A/B-test another build of the same ABI (developer switch; the default is the in-tree library)
LIB_PATH = os.environ.get("MMOE_B200_LIB") or os.path.join(_HERE, "libmmoe_b200.so")

F32, BF16, F16 = 0, 1, 2


class Epilogue(C.Structure):
    _fields_ = [
        ("out", C.c_void_p), ("out_dtype", C.c_int32), ("accumulate", C.c_int32), ("ldo", C.c_int64),
        ("bias", C.c_void_p), ("preact", C.c_void_p), ("act", C.c_int32), ("bwd_mode", C.c_int32),
        ("aux", C.c_void_p), ("ld_aux", C.c_int64), ("residual", C.c_void_p), ("ld_res", C.c_int64),
        ("colsum", C.c_void_p), ("alpha", C.c_float), ("drop_p", C.c_float),
        ("drop_key0", C.c_uint32), ("drop_key1", C.c_uint32), ("mask_out", C.c_void_p),
    ]


class GemmProblem(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64), ("a_major", C.c_int32),
        ("b", C.c_void_p), ("ldb", C.c_int64), ("b_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("k_splits", C.c_int32),
        ("epi", Epilogue),
    ]


class Call(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("B", C.c_int32), ("training", C.c_int32), ("home", C.c_int32),
        ("drop_p", C.c_float), ("seed", C.c_uint64),
        ("params", C.POINTER(C.c_void_p)), ("grads", C.POINTER(C.c_void_p)),
        ("saved", C.c_void_p), ("saved_bytes", C.c_size_t),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("stream", C.c_void_p),
    ]


class HeadCfg(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_expert", C.c_int32), ("hidden", C.c_int32), ("tower_drop_p", C.c_float)]


class CrossCfg(C.Structure):
    _fields_ = [("d", C.c_int32), ("S", C.c_int32), ("n_head", C.c_int32), ("n_layer", C.c_int32)]


class FuseCfg(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_head", C.c_int32), ("depth", C.c_int32)]


class HomeCfg(C.Structure):
    _fields_ = [("d", C.c_int32), ("n_in", C.c_int32), ("n_shared", C.c_int32), ("n_task", C.c_int32),
                ("tower_hidden", C.c_int32), ("expert_hidden", C.c_int32)]


ABI_STRUCTS = [Epilogue, GemmProblem, Call, HeadCfg, CrossCfg, FuseCfg, HomeCfg]

_P = C.POINTER
_vp, _i32, _i64, _f, _u32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32, C.c_size_t

# name -> (restype, argtypes); every symbol include/mmoe_b200.h declares
PROTOTYPES = {
    "mmoe_abi_version": (C.c_int, []),
    "mmoe_abi_sizeof": (_sz, [C.c_int]),
    "mmoe_last_error": (C.c_char_p, []),
    "mmoe_init": (C.c_int, []),
    "mmoe_launch_count": (_i64, [C.c_int]),
    "mmoe_cast_f32": (C.c_int, [_vp, _vp, _i64, C.c_int, _vp]),
    "mmoe_gemm_grouped": (C.c_int, [_P(GemmProblem), C.c_int, C.c_int, C.c_int, _vp]),
    "mmoe_gemm_timing": (C.c_int, [C.c_int]),
    "mmoe_gemm_timing_read": (C.c_int, [_P(C.c_double), _P(C.c_double), _P(_i64), C.c_int]),
    "mmoe_gemm_timing_dump": (C.c_int, [_P(C.c_double), C.c_int]),
    "mmoe_set_sm_reserve": (C.c_int, [C.c_int]),
    "mmoe_launch_trace": (C.c_int, [C.c_int]),
    "mmoe_launch_trace_read": (C.c_int, [_P(_i32), C.c_int]),
    "mmoe_dropout_mask": (C.c_int, [_u32, _u32, _f, _i64, _vp, _vp]),
    "mmoe_site_keys": (C.c_int, [C.c_uint64, _u32, _P(_u32), _P(_u32)]),
    "mmoe_layernorm_fwd": (C.c_int, [_vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _i64, _i32, C.c_int, _vp]),
    "mmoe_attention_fwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32,
                                     _f, _u32, _u32, C.c_int, _vp]),
    "mmoe_attention_bwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp,
                                     _i32, _i32, _i32, _i32, _i32, _f, _u32, _u32, C.c_int, _vp]),
    "mmoe_head_saved_bytes": (_sz, [_P(HeadCfg), _i32, C.c_int]),
    "mmoe_head_workspace_bytes": (_sz, [_P(HeadCfg), _i32, C.c_int]),
    "mmoe_head_fwd": (C.c_int, [_P(Call), _P(HeadCfg), _vp, _vp, _vp]),
    "mmoe_head_bwd": (C.c_int, [_P(Call), _P(HeadCfg), _vp, _vp, _vp]),
    "mmoe_dense_gate_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "mmoe_dense_gate_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "mmoe_cross_saved_bytes": (_sz, [_P(CrossCfg), _i32, C.c_int]),
    "mmoe_cross_workspace_bytes": (_sz, [_P(CrossCfg), _i32, C.c_int]),
    "mmoe_cross_fwd": (C.c_int, [_P(Call), _P(CrossCfg), _vp, _vp, _vp, _vp, _vp]),
    "mmoe_cross_bwd": (C.c_int, [_P(Call), _P(CrossCfg), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmoe_cross_bwd_stage": (C.c_int, [_P(Call), _P(CrossCfg), C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmoe_cross_saved_offset": (C.c_int, [_P(CrossCfg), _i32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P(_sz), _P(_sz)]),
    "mmoe_fuse_saved_offset": (C.c_int, [_P(FuseCfg), _i32, C.c_int, C.c_int, C.c_int, C.c_int, _P(_sz), _P(_sz)]),
    "mmoe_fuse_saved_bytes": (_sz, [_P(FuseCfg), _i32, C.c_int]),
    "mmoe_fuse_workspace_bytes": (_sz, [_P(FuseCfg), _i32, C.c_int]),
    "mmoe_fuse_fwd": (C.c_int, [_P(Call), _P(FuseCfg), _vp, _vp, _vp]),
    "mmoe_fuse_bwd": (C.c_int, [_P(Call), _P(FuseCfg), _vp, _vp]),
    "mmoe_home_saved_bytes": (_sz, [_P(HomeCfg), _i32, C.c_int]),
    "mmoe_home_workspace_bytes": (_sz, [_P(HomeCfg), _i32, C.c_int]),
    "mmoe_home_fwd": (C.c_int, [_P(Call), _P(HomeCfg), _vp, _vp, _vp]),
    "mmoe_home_bwd": (C.c_int, [_P(Call), _P(HomeCfg), _vp, _vp, _vp]),
    "mmoe_img_pool_fwd": (C.c_int, [_P(Call), _vp, C.c_int, _i32, _i32, _i32, _vp, _vp, _vp]),
    "mmoe_img_pool_bwd": (C.c_int, [_P(Call), _i32, _i32, _i32, _vp, _vp, _vp, _vp, C.c_int]),
    "mmoe_img_proj_saved_bytes": (_sz, [_i32, _i32, _i32, C.c_int]),
    "mmoe_img_proj_workspace_bytes": (_sz, [_i32, _i32, _i32, C.c_int]),
    "mmoe_img_proj_fwd": (C.c_int, [_P(Call), _i32, _i32, _vp, _vp]),
    "mmoe_img_proj_bwd": (C.c_int, [_P(Call), _i32, _i32, _vp, _vp]),
    "mmoe_bn_silu_stack_fwd": (C.c_int, [_P(Call), _i32, _i32, _P(_vp), _vp, _vp, _vp, _P(_vp), _P(_vp), _f, _f]),
    "mmoe_bn_silu_stack_bwd": (C.c_int, [_P(Call), _i32, _i32, _P(_vp), _vp, _vp, _vp, _P(_vp), _P(_vp), _P(_vp), _f]),
    "mmoe_bce2_fwd_bwd": (C.c_int, [_vp, _vp, _vp, _f, _f, _i32, _vp, _vp, _f, _vp]),
    "mmoe_info_nce_saved_bytes": (_sz, [_i32, _i32, _i32, C.c_int]),
    "mmoe_info_nce_workspace_bytes": (_sz, [_i32, _i32, _i32, C.c_int]),
    "mmoe_info_nce_fwd": (C.c_int, [_P(Call), _i32, _i32, _P(_vp), _P(_vp), _f, _vp]),
    "mmoe_info_nce_bwd": (C.c_int, [_P(Call), _i32, _i32, _P(_vp), _P(_vp), _P(_f), _P(_vp), _P(_vp)]),
    "mmoe_auc_workspace_bytes": (_sz, [_i64]),
    "mmoe_auc": (C.c_int, [_vp, _vp, _i64, _vp, _sz, _vp, _vp]),
    "mmoe_patch_u8_to_operand": (C.c_int, [_vp, _vp, _i64, _i32, C.c_int, _vp]),
    "mmoe_patchify": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, C.c_int, _vp]),
    "mmoe_patch_project": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _i64, _i32, _i32, C.c_int, _vp]),
    "mmoe_sent_gather_fwd": (C.c_int, [_P(Call), _i32, _i32, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mmoe_sent_gather_bwd": (C.c_int, [_P(Call), _i32, _i32, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()


def _build_if_needed():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mmoe_b200_build", os.path.join(_HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build(force=False, verbose=True)


def lib():
    """The loaded shared object (built on first use if the .so is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            _build_if_needed()
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)          # AttributeError here = the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        for i, st in enumerate(ABI_STRUCTS):
            n = L.mmoe_abi_sizeof(i)
            if n != C.sizeof(st):
                raise ImportError(f"ABI mismatch: {st.__name__} is {C.sizeof(st)} bytes in Python, {n} in the library")
        _lib = L
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().mmoe_last_error()
        raise RuntimeError(f"mmoe_b200 {what} failed ({rc}): {msg.decode() if msg else 'unknown error'}")


def ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)
