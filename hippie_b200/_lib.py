"""ctypes binding of libhippie_b200.so (the C ABI declared in include/hippie_b200.h).

There is no CPU fallback: if the shared library is missing the import of anything that
needs it raises, telling the user to build it (`python -c "import __graft_entry__ as g; g.build()"`
or `make -C hippie_b200/csrc`).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhippie_b200.so")

ABI_VERSION = 3


class HippieCfg(C.Structure):
    """struct hippie_cfg (include/hippie_b200.h)."""
    _fields_ = [
        ("z_dim", C.c_int32),
        ("class_hidden_dim", C.c_int32),
        ("num_sources", C.c_int32),
        ("num_classes", C.c_int32),
        ("len_wave", C.c_int32),
        ("len_isi", C.c_int32),
        ("multimodal", C.c_int32),
        ("max_batch", C.c_int32),
        ("inference_only", C.c_int32),
        ("conv_path", C.c_int32),
    ]


_lib = None

_f32p = C.c_void_p  # device pointers travel as integers
_i64p = C.c_void_p
_H = C.c_void_p

# name -> (restype, argtypes); mirrors include/hippie_b200.h one to one
SIGNATURES = {
    "hippie_abi_version": (C.c_int, []),
    "hippie_create": (C.c_int, [C.POINTER(HippieCfg), C.POINTER(_H)]),
    "hippie_destroy": (None, [_H]),
    "hippie_last_error": (C.c_char_p, [_H]),
    "hippie_num_params": (C.c_int, [_H]),
    "hippie_param_floats": (C.c_int64, [_H]),
    "hippie_param_info": (C.c_int, [_H, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                    C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "hippie_num_bn": (C.c_int, [_H]),
    "hippie_bn_floats": (C.c_int64, [_H]),
    "hippie_bn_info": (C.c_int, [_H, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "hippie_workspace_bytes": (C.c_size_t, [_H]),
    "hippie_num_tensors": (C.c_int, [_H]),
    "hippie_tensor_info": (C.c_int, [_H, C.c_int, C.c_char_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32),
                                     C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "hippie_bind": (C.c_int, [_H, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _i64p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "hippie_train_fwd_bwd": (C.c_int, [_H, _f32p, _f32p, _i64p, _i64p, _f32p, C.c_int32, C.c_float, C.c_float,
                                       C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "hippie_train_fwd_bwd_part": (C.c_int, [_H, _f32p, _f32p, _i64p, _i64p, _f32p, C.c_int32, C.c_float, C.c_float,
                                            C.c_float, _f32p, C.c_int32, C.c_void_p]),
    "hippie_grad_split": (C.c_int64, [_H]),
    "hippie_grad_bounds": (C.c_int, [_H, C.POINTER(C.c_int64)]),
    "hippie_eval_forward": (C.c_int, [_H, _f32p, _f32p, _i64p, _i64p, _f32p, C.c_int32, C.c_float, C.c_float,
                                      C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "hippie_train_forward": (C.c_int, [_H, _f32p, _f32p, _i64p, _i64p, _f32p, C.c_int32, C.c_float, C.c_float,
                                       C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "hippie_embed": (C.c_int, [_H, _f32p, _f32p, _i64p, _i64p, C.c_int32, C.c_int32, _f32p, _f32p, _f32p, C.c_void_p]),
    "hippie_encoder_forward": (C.c_int, [_H, C.c_int32, _f32p, C.c_int32, C.c_int32, _f32p, C.c_void_p]),
    "hippie_decoder_forward": (C.c_int, [_H, C.c_int32, _f32p, C.c_int32, C.c_int32, _f32p, C.c_void_p]),
    "hippie_encode": (C.c_int, [_H, _f32p, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, _f32p, _f32p, _f32p, C.c_void_p]),
    "hippie_decode": (C.c_int, [_H, _f32p, _f32p, _f32p, C.c_int32, C.c_int32, _f32p, _f32p, C.c_void_p]),
    "hippie_slice_wait": (C.c_int, [_H, C.c_int32, C.c_void_p]),
    "hippie_device_flags": (C.c_int, [_H, C.POINTER(C.c_uint32), C.c_int32, C.c_void_p]),
    "hippie_params_changed": (C.c_int, [_H]),
    "hippie_clip_adamw": (C.c_int, [_H, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_float, C.c_float,
                                    C.c_int32, C.c_int32, C.c_int32, _f32p, C.c_void_p]),
    "hippie_preprocess_batch": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, _i64p, C.c_int32, _f32p, C.c_int32,
                                          _f32p, C.c_int32, C.c_void_p]),
    "hippie_knn_neighbors": (C.c_int, [_f32p, C.c_int64, _f32p, C.c_int64, C.c_int32, C.c_int32, _i64p, C.c_void_p,
                                       C.c_void_p]),
    "hippie_knn_evaluate": (C.c_int, [_i64p, C.c_int64, C.c_int32, _i64p, _i64p, C.c_int32, C.c_int32, C.c_int32, _i64p,
                                      _i64p, C.c_void_p, C.c_void_p]),
    "hippie_last_launch_count": (C.c_int, [_H]),
    "hippie_conv_path_in_use": (C.c_int, [_H]),
    "hippie_profile": (C.c_int, [_H, C.c_int]),
    "hippie_profile_read": (C.c_int, [_H, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int)]),
}


def lib():
    """Loads the shared library once.  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C hippie_b200/csrc` (or "
                "`python -c 'import __graft_entry__ as g; g.build()'`). hippie_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.hippie_abi_version() != ABI_VERSION:
            raise RuntimeError("libhippie_b200.so ABI version mismatch; rebuild")
        _lib = L
    return _lib
