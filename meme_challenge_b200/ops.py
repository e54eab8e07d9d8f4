"""Thin Python launchers over the C-ABI (one function per kernel entry point).

These do no arithmetic of their own: they validate dtypes/shapes, allocate outputs with torch
(PyTorch owns every buffer) and call libb200u on the current CUDA stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ADD, EPI_ATOMIC_F32, EPI_BIAS_DROP_RES, EPI_BIAS_GELU, EPI_DGELU, EPI_STORE,
                   EPI_STORE_F32)


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor"
    return t.stride(0)


def gemm(a, b, *, a_mn=False, b_mn=False, epilogue=EPI_STORE, out=None, out2=None, bias=None,
         res=None, drop=None, splits=0, block_n=0, impl=0):
    """acc[M,N] = sum_k A(m,k) B(n,k) with a fused epilogue (see include/b200u.h, K3).

    a: [M,K] (or [K,M] when a_mn), b: [N,K] (or [K,N] when b_mn); bf16, row-major.
    """
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, "inner dimensions differ: %d vs %d" % (K, Kb)
    f32_out = epilogue in (EPI_ATOMIC_F32, EPI_STORE_F32)
    if out is None:
        assert epilogue != EPI_ATOMIC_F32, "EPI_ATOMIC_F32 accumulates into an existing buffer"
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    assert out.dtype == (torch.float32 if f32_out else torch.bfloat16) and tuple(out.shape) == (M, N)
    if epilogue == EPI_BIAS_GELU and out2 is None:
        out2 = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    g = _lib.GemmT()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn_major = a.data_ptr(), _ld(a), int(a_mn)
    g.B, g.ldb, g.b_mn_major = b.data_ptr(), _ld(b), int(b_mn)
    g.epilogue = epilogue
    g.C, g.ldc = out.data_ptr(), _ld(out)
    if out2 is not None:
        g.C2, g.ldc2 = out2.data_ptr(), _ld(out2)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N and bias.is_contiguous()
        g.bias = bias.data_ptr()
    if res is not None:
        assert res.dtype == torch.bfloat16 and tuple(res.shape) == (M, N)
        g.R, g.ldr = res.data_ptr(), _ld(res)
    if drop is not None:
        g.drop = drop
    g.splits, g.block_n, g.impl = splits, block_n, impl
    _lib.check(_lib.lib().b200u_gemm(C.byref(g), _lib.stream_ptr()), "b200u_gemm")
    return (out, out2) if epilogue == EPI_BIAS_GELU else out
