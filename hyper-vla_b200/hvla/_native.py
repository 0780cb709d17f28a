"""ctypes binding of libhvla.so (the C ABI declared in include/hvla.h).

The product path has no CPU fallback: if the library is missing this module raises at import
of the symbols (``lib()``), and every compute call raises ``HvlaError`` on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import os

HVLA_F32, HVLA_BF16, HVLA_BF16X3 = 0, 1, 2
_LIB = None

c_void_p, c_int, c_i64, c_size_t = C.c_void_p, C.c_int, C.c_int64, C.c_size_t

# name -> (restype, argtypes); must list every symbol include/hvla.h declares
SIGNATURES = {
    "hvla_version": (c_int, []),
    "hvla_last_error": (C.c_char_p, []),
    "hvla_hn_blob_elems": (c_i64, []),
    "hvla_generated_elems": (c_i64, []),
    "hvla_generated_row_stride": (c_i64, []),
    "hvla_dino_vec_elems": (c_i64, []),
    "hvla_dino_mat_elems": (c_i64, []),
    "hvla_layout_offset": (c_i64, [C.c_char_p]),
    "hvla_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "hvla_generate": (c_int, [c_void_p] * 9 + [c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_generate_rows": (c_int, [c_void_p] * 9 + [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_dino_forward": (c_int, [c_void_p] * 4 + [c_int, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_base_act": (c_int, [c_void_p] * 4 + [c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_act": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_act_debug": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_discrete_generated_elems": (c_i64, [c_int]),
    "hvla_discrete_row_stride": (c_i64, [c_int]),
    "hvla_generate_n": (c_int, [c_void_p] * 9 + [c_int, c_i64, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_act_discrete": (c_int, [c_void_p] * 6 + [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_act_host": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
    "hvla_gemm_bf16": (c_int, [c_void_p] * 5 + [c_int, c_int, c_int, c_int]),
    "hvla_dino_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "hvla_postprocess_state_floats": (c_i64, []),
    "hvla_postprocess": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int,
                                 C.c_float, c_int, c_int, c_void_p, c_void_p]),
    "hvla_resize_workspace_bytes": (c_size_t, [c_int] * 5),
    "hvla_resize_lanczos3": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                                     c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t]),
    "hvla_t5_blob_elems": (c_i64, []),
    "hvla_t5_workspace_bytes": (c_size_t, [c_int, c_int]),
    "hvla_t5_encode": (c_int, [c_void_p] * 5 + [c_int, c_int, c_void_p, c_void_p, c_size_t]),
    "hvla_t5_mat_elems": (c_i64, []),
    "hvla_t5_tc_workspace_bytes": (c_size_t, [c_int, c_int]),
    "hvla_t5_encode_tc": (c_int, [c_void_p] * 6 + [c_int, c_int, c_void_p, c_void_p, c_size_t]),
    "hvla_launch_count": (c_i64, []),
    "hvla_profile_enable": (c_int, [c_int]),
    "hvla_profile_report": (c_int, [C.c_char_p, c_size_t]),
    "hvla_xla_generate": (None, [c_void_p, c_void_p, C.c_char_p, c_size_t]),
    "hvla_xla_act": (None, [c_void_p, c_void_p, C.c_char_p, c_size_t]),
    "hvla_xla_register_status_setter": (None, [c_void_p]),
    "hvla_xla_generate_status": (None, [c_void_p, c_void_p, C.c_char_p, c_size_t, c_void_p]),
    "hvla_xla_act_status": (None, [c_void_p, c_void_p, C.c_char_p, c_size_t, c_void_p]),
}


class HvlaError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhvla.so")


def lib():
    """Load libhvla.so (building it first if nvcc is present and it is missing/stale)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        from . import _build
        _build.build()
    if not os.path.exists(path):
        raise HvlaError(f"libhvla.so not found at {path}; run `python __graft_entry__.py build` (no CPU fallback exists)")
    L = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _LIB = L
    return L


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().hvla_last_error().decode("utf-8", "replace")
        raise HvlaError(f"{what} failed with status {status}: {msg}")


def layout_offset(name: str) -> int:
    return int(lib().hvla_layout_offset(name.encode()))
