"""The C-ABI shared library builds without a GPU, loads, and exports every function that
include/b200_bridge.h declares (no compute calls here: those are the `-m gpu` tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200_bridge.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)          # drop comments
    src = re.sub(r"typedef struct.*?}\s*\w+;", "", src, flags=re.S)
    src = re.sub(r"enum\s+\w+\s*{.*?};", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200b_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from vlm_bridge_b200 import build

    return ctypes.CDLL(build.build())


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("b200b_gemm", "b200b_layernorm_fwd", "b200b_layernorm_bwd", "b200b_colsum", "b200b_cast_bf16",
                 "b200b_attention_fwd", "b200b_attention_bwd", "b200b_bridge_kv_project",
                 "b200b_bridge_block_forward", "b200b_bridge_block_backward", "b200b_bridge_kv_backward",
                 "b200b_last_error", "b200b_abi_version"):
        assert must in names


def test_every_declared_symbol_is_exported(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_and_error_string(lib):
    assert lib.b200b_abi_version() == 3
    lib.b200b_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.b200b_last_error(), bytes)


def test_argument_validation_needs_no_gpu(lib):
    """null / bad-shape arguments are rejected on the host before any CUDA call"""
    lib.b200b_last_error.restype = ctypes.c_char_p
    assert lib.b200b_gemm(None, None) == -3                                  # B200B_ERR_ARG
    assert b"null" in lib.b200b_last_error()
    lib.b200b_layernorm_fwd.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                                               ctypes.c_void_p]
    assert lib.b200b_layernorm_fwd(None, None, None, None, None, None, 4, 8, 1e-5, None) == -3
    lib.b200b_bridge_block_saved_bytes.restype = ctypes.c_size_t
    assert lib.b200b_bridge_block_saved_bytes(None) == 0
    # fused cross-entropy: null pointers, a dtype that is neither f32 nor bf16, a shift length that does not divide the rows
    i64, vp = ctypes.c_int64, ctypes.c_void_p
    lib.b200b_cross_entropy_fwd.argtypes = [vp, ctypes.c_int, i64, vp, i64, i64, i64, i64, vp, vp, vp, vp]
    assert lib.b200b_cross_entropy_fwd(None, 0, 8, None, 4, 8, -100, 0, None, None, None, None) == -3
    buf = ctypes.create_string_buffer(64)                   # a non-null pointer that is never dereferenced on the host
    addr = ctypes.addressof(buf)
    assert lib.b200b_cross_entropy_fwd(addr, 2, 8, addr, 4, 8, -100, 0, addr, addr, addr, None) == -3
    assert b"dtype" in lib.b200b_last_error()
    assert lib.b200b_cross_entropy_fwd(addr, 0, 8, addr, 4, 8, -100, 3, addr, addr, addr, None) == -1    # B200B_ERR_SHAPE
    assert lib.b200b_cross_entropy_fwd(addr, 0, 4, addr, 4, 8, -100, 0, addr, addr, addr, None) == -3    # ld < vocab


def test_product_fails_loudly_without_cuda():
    import torch

    from vlm_bridge_b200 import BridgeLite

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = BridgeLite(vision_dim=32, language_dim=64, num_heads_cross=1, num_heads_self=1)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 4, 32), torch.randn(1, 3, 64))
