"""Thin torch-tensor wrappers over the C ABI (one function per exported kernel family).

torch supplies device memory and the current stream only; all arithmetic happens inside
libb200_bridge.so. Every wrapper validates devices/dtypes, then passes raw pointers.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_BF16_BIAS, EPI_BF16_BIAS_GELU, EPI_BF16_DGELU, EPI_F32,  # noqa: F401
                   EPI_F32_BIAS_RESID)


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: torch.Tensor | None) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("b200 bridge kernels need CUDA tensors (there is no CPU fallback)")


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_major: int = 0, b_major: int = 0,
         epilogue: int = EPI_BF16_BIAS, out: torch.Tensor | None = None,
         bias: torch.Tensor | None = None, aux: torch.Tensor | None = None,
         resid: torch.Tensor | None = None, beta: float = 0.0, dropout_p: float = 0.0,
         seed: int = 0, dropout_stream: int = 0, block_n: int = 0) -> torch.Tensor:
    """acc[m,n] = sum_k A(m,k) B(n,k) on tcgen05 tensor cores with a fused epilogue.

    a_major=0: a is [M,K]; a_major=1: a is [K,M]. b_major=0: b is [N,K]; b_major=1: b is [K,N].
    """
    _need_cuda(a, b, out, bias, aux, resid)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise RuntimeError("gemm operands must be bf16")
    if a.dim() != 2 or b.dim() != 2 or a.stride(1) != 1 or b.stride(1) != 1:
        raise RuntimeError("gemm operands must be 2-D with unit inner stride")
    m, k = (a.shape[1], a.shape[0]) if a_major else (a.shape[0], a.shape[1])
    n, kb = (b.shape[1], b.shape[0]) if b_major else (b.shape[0], b.shape[1])
    if k != kb:
        raise RuntimeError(f"gemm: contraction mismatch {k} vs {kb}")
    out_f32 = epilogue in (EPI_F32_BIAS_RESID, EPI_F32)
    if out is None:
        out = torch.empty((m, n), device=a.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    if out.dtype != (torch.float32 if out_f32 else torch.bfloat16) or out.shape != (m, n) or out.stride(1) != 1:
        raise RuntimeError("gemm: bad `out` tensor")
    if epilogue == EPI_BF16_BIAS_GELU and aux is None:
        aux = torch.empty((m, n), device=a.device, dtype=torch.bfloat16)
    args = _lib.GemmArgs(
        a=a.data_ptr(), b=b.data_ptr(), a_major=a_major, b_major=b_major, m=m, n=n, k=k,
        lda=a.stride(0), ldb=b.stride(0), epilogue=epilogue, block_n=block_n,
        out=out.data_ptr(), ldo=out.stride(0),
        aux=_ptr(aux), ldaux=0 if aux is None else aux.stride(0),
        bias=_ptr(bias), resid=_ptr(resid), ldr=0 if resid is None else resid.stride(0),
        beta=beta, dropout_p=dropout_p, seed=seed, dropout_stream=dropout_stream, reserved=0)
    _lib.check(_lib.lib().b200b_gemm(C.byref(args), _stream_ptr()), "gemm")
    if epilogue == EPI_BF16_BIAS_GELU:
        return out, aux
    return out
