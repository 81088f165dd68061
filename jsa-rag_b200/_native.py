"""ctypes binding of libjsa_mips.so (C ABI declared in include/jsa_mips.h).

There is deliberately no fallback: if the shared library has not been built, or no sm_100 device
is present, every entry point raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``python jsa-rag_b200/build.py --force -v``).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjsa_mips.so")

MIPS_DTYPE_F16, MIPS_DTYPE_BF16, MIPS_DTYPE_F32 = 0, 1, 2
MIPS_OK, MIPS_EINVAL, MIPS_EKRANGE, MIPS_ECUDA, MIPS_ENOTBOUND, MIPS_EWORKSPACE, MIPS_EUNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
MIPS_ETIMEOUT = -7

# every symbol include/jsa_mips.h declares: (restype, argtypes)
SYMBOLS = {
    "mips_abi_version": (c_int, []),
    "mips_max_k": (c_int, []),
    "mips_max_dim": (c_int, []),
    "mips_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int]),
    "mips_destroy": (None, [c_void_p]),
    "mips_last_error": (c_char_p, [c_void_p]),
    "mips_bind_index": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64]),
    "mips_bind_index_layout": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64]),
    "mips_workspace_bytes": (c_int, [c_void_p, c_int, c_int, POINTER(c_size_t)]),
    "mips_workspace_pin": (c_int, [c_void_p, c_int]),
    "mips_search_local": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_void_p, c_size_t, c_void_p]),
    "mips_merge_topk": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mips_merge_topk_strided": (c_int, [c_int, c_void_p, c_void_p, c_int, c_int64, c_int64, c_int, c_int, c_int, c_void_p,
                                        c_void_p, c_void_p]),
    "mips_gather_rows": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "mips_rerank": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p]),
    "mips_max_rerank_candidates": (c_int, []),
    "mips_xchg_handle_bytes": (c_int, []),
    "mips_xchg_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_size_t]),
    "mips_xchg_export": (c_int, [c_void_p, c_void_p]),
    "mips_xchg_connect": (c_int, [c_void_p, c_void_p]),
    "mips_xchg_capacity": (c_size_t, [c_void_p]),
    "mips_xchg_merge": (c_int, [c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mips_xchg_gather": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "mips_xchg_push": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "mips_xchg_merge_wait": (c_int, [c_void_p, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mips_xchg_gather_wait": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "mips_xchg_set_timeout_ms": (c_int, [c_void_p, c_int64]),
    "mips_xchg_status": (c_int, [c_void_p]),
    "mips_xchg_connect_local": (c_int, [c_void_p, POINTER(c_void_p), c_int]),
    "mips_xchg_last_error": (c_char_p, [c_void_p]),
    "mips_xchg_destroy": (c_int, [c_void_p]),
    "mips_search_host": (c_int, [c_void_p, POINTER(c_float), c_int, c_int, c_int, POINTER(c_float), POINTER(c_int64),
                                 c_void_p]),
    "mips_search_host_async": (c_int, [c_void_p, POINTER(c_float), c_int, c_int, c_int, POINTER(c_float), POINTER(c_int64),
                                       c_void_p]),
    "mips_parse_float_list": (c_int64, [c_char_p, c_size_t, POINTER(c_float), c_int64]),
    "mips_last_launch_count": (c_int, [c_void_p]),
    "mips_debug_config": (c_int, [c_void_p, c_int, c_void_p]),
    "mips_debug_num_stats": (c_int, []),
    "mips_scan_times_ms": (c_int, [c_void_p, POINTER(c_float), c_int, POINTER(c_int)]),
    "mips_num_sms": (c_int, [c_void_p]),
}

_lib = None


class NativeLibraryMissing(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Loads the CUDA extension; raises NativeLibraryMissing (never falls back) if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: the B200 MIPS engine has no CPU/PyTorch fallback. "
            "Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.mips_abi_version() != 1:
        raise NativeLibraryMissing(f"{LIB_PATH}: ABI version {lib.mips_abi_version()} != 1; rebuild")
    _lib = lib
    return lib


def is_built() -> bool:
    return os.path.isfile(LIB_PATH)


def last_error(handle=None) -> str:
    msg = load().mips_last_error(handle)
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, handle=None, what: str = "") -> None:
    """Maps C error codes to the exception types the reference raises at the same points."""
    if rc == MIPS_OK:
        return
    msg = last_error(handle) or f"error {rc}"
    if rc == MIPS_EKRANGE:
        # torch.topk in the reference (src/index.py:119) raises RuntimeError("selected index k out of range")
        raise RuntimeError(msg)
    if rc == MIPS_EINVAL:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")
