"""ctypes binding of libfa_b200.so (the C ABI in include/fa_b200.h).

There is no CPU fallback: if the CUDA library is missing, import of the compute entry points fails loudly.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_int, c_size_t, c_void_p
from pathlib import Path

from ._build import LIB_PATH

FA_DTYPE_F32, FA_DTYPE_BF16, FA_DTYPE_F16, FA_DTYPE_F64 = 0, 1, 2, 3
FA_OK = 0
ERROR_NAMES = {-1: "FA_ERR_SHAPE", -2: "FA_ERR_DTYPE", -3: "FA_ERR_ALIGN", -4: "FA_ERR_UNSUPPORTED_D",
               -5: "FA_ERR_CUDA", -6: "FA_ERR_WORKSPACE"}

# every symbol include/fa_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "fa_last_error": (c_char_p, []),
    "fa_device_sm_count": (c_int, []),
    "fa_v1_forward": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "fa_v1_forward_ex": (c_int, [c_void_p] * 5 + [c_int] * 5 + [ctypes.c_uint, c_void_p]),
    "fa_v1_forward_varlen": (c_int, [c_void_p] * 6 + [c_int] * 6 + [ctypes.c_uint, c_void_p]),
    "fa_partial_forward": (c_int, [c_void_p] * 5 + [c_int] * 6 + [ctypes.c_longlong] * 3 + [ctypes.c_uint, c_void_p]),
    "fa_v1_tiled_d_forward": (c_int, [c_void_p] * 4 + [c_int] * 7 + [c_void_p]),
    "fa_v1_tiled_d_pair_forward": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "fa_v2_num_splits": (c_int, [c_int, c_int]),
    "fa_v2_workspace_bytes": (c_size_t, [c_int] * 5),
    "fa_v2_splitkv_forward": (c_int, [c_void_p] * 5 + [c_int] * 6 + [c_void_p]),
    "fa_v2_combine": (c_int, [c_void_p] * 3 + [c_int] * 6 + [c_void_p]),
    "fa_v2_forward": (c_int, [c_void_p] * 4 + [c_int] * 6 + [c_void_p, c_size_t, c_void_p]),
    "fa_v1_backward_workspace_bytes": (c_size_t, [c_int] * 3),
    "fa_v1_backward": (c_int, [c_void_p] * 9 + [c_int] * 5 + [ctypes.c_uint, c_void_p, c_size_t, c_void_p]),
    "fa_naive_attention_workspace_bytes": (c_size_t, [c_int] * 4),
    "fa_naive_attention": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p, c_size_t, c_void_p]),
    "fa_forward_host": (c_int, [c_int] + [c_void_p] * 4 + [c_int] * 6),
    "fa_copy_2d_async": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "fa_copy_2d_multi_async": (c_int, [c_int, c_void_p, c_size_t, c_void_p, c_size_t, c_size_t, c_size_t, c_void_p]),
    "fa_copy_multi_async": (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fa_release_host_staging": (None, []),
    "fa_debug_map_cache_stats": (None, [c_void_p, c_void_p]),
}


class FlashAttentionError(RuntimeError):
    """Non-zero status from libfa_b200.so (the reference aborts via assert(); we raise)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


_lib = None


def load() -> ctypes.CDLL:
    """Load the library (no compute is triggered). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(LIB_PATH)
    if not path.exists():
        # Not a fallback: the same CUDA library, compiled in-tree on first use when a checkout has no built artefact yet.
        try:
            from ._build import build
            build()
        except Exception as e:  # noqa: BLE001
            raise RuntimeError(
                f"{path} is missing and could not be built ({e}); build it with "
                "`python -m exploring_flash_attention_b200._build` (or __graft_entry__.build()). "
                "This package has no CPU fallback.") from e
    lib = ctypes.CDLL(str(path))
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != FA_OK:
        raise FlashAttentionError(code, load().fa_last_error().decode())
