"""ctypes binding of libb200_bridge.so (C ABI declared in include/b200_bridge.h).

There is no CPU fallback: if the shared library is missing or a call fails, a RuntimeError is
raised. `torch` is only used by callers for device memory and streams; no torch type crosses the
ABI (plain pointers and sizes).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "csrc", "libb200_bridge.so")

# enum b200b_epilogue
EPI_BF16_BIAS = 0
EPI_BF16_BIAS_GELU = 1
EPI_F32_BIAS_RESID = 2
EPI_BF16_DGELU = 3
EPI_F32 = 4


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p),
        ("b", C.c_void_p),
        ("a_major", C.c_int32),
        ("b_major", C.c_int32),
        ("m", C.c_int32),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("lda", C.c_int64),
        ("ldb", C.c_int64),
        ("epilogue", C.c_int32),
        ("block_n", C.c_int32),
        ("out", C.c_void_p),
        ("ldo", C.c_int64),
        ("aux", C.c_void_p),
        ("ldaux", C.c_int64),
        ("bias", C.c_void_p),
        ("resid", C.c_void_p),
        ("ldr", C.c_int64),
        ("beta", C.c_float),
        ("dropout_p", C.c_float),
        ("seed", C.c_uint64),
        ("dropout_stream", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


_lock = threading.Lock()
_lib = None


def _declare(lib) -> None:
    lib.b200b_abi_version.restype = C.c_int
    lib.b200b_abi_version.argtypes = []
    lib.b200b_last_error.restype = C.c_char_p
    lib.b200b_last_error.argtypes = []
    lib.b200b_launch_count.restype = C.c_uint64
    lib.b200b_launch_count.argtypes = []
    lib.b200b_gemm.restype = C.c_int
    lib.b200b_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]


def lib():
    """Load (once) and return the shared library; raise if it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m vlm_bridge_b200.build` "
                        "(there is no CPU or PyTorch fallback for the bridge kernels)"
                    )
                handle = C.CDLL(LIB_PATH)
                _declare(handle)
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b200b_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"b200b {what} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().b200b_launch_count())
