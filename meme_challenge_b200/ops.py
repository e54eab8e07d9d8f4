"""Thin Python launchers over the C-ABI (one function per kernel entry point).

These do no arithmetic of their own: they validate dtypes/shapes, allocate outputs with torch
(PyTorch owns every buffer) and call libb200u on the current CUDA stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ADD, EPI_ATOMIC_F32, EPI_BIAS_DROP_RES, EPI_BIAS_DROP_RES_LN, EPI_BIAS_GELU,
                   EPI_BIAS_GELU_DG, EPI_DGELU, EPI_DUAL, EPI_MUL, EPI_STORE, EPI_STORE_F32)


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, "expected a row-major 2-D tensor"
    return t.stride(0)


def gemm(a, b, *, a_mn=False, b_mn=False, epilogue=EPI_STORE, out=None, out2=None, bias=None,
         res=None, drop=None, splits=0, block_n=0, impl=0, cluster=0, colsum=None, ln=None, ce=None):
    """acc[M,N] = sum_k A(m,k) B(n,k) with a fused epilogue (see include/b200u.h, K3).

    a: [M,K] (or [K,M] when a_mn), b: [N,K] (or [K,N] when b_mn); bf16, row-major.
    colsum (EPI_MUL): f32 [N], += column sums of the output. ln (EPI_BIAS_DROP_RES_LN): tuple
    (gamma f32 [N], beta f32 [N], eps, mean f32 [M] or None, rstd f32 [M] or None); out = pre-LayerNorm
    values, out2 = LayerNorm output. Dual-output epilogues return (out, out2).
    ce (EPI_CE_STATS / EPI_CE_GRAD): dict with `target` (i64 [M]) and, for STATS, `partial` (f32 [M, ceil(N/128), 2])
    and `tlogit` (f32 [M]) -- no output matrix, returns None --, for GRAD `lse` and `scale` (f32 [M]).
    """
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    assert K == Kb, "inner dimensions differ: %d vs %d" % (K, Kb)
    f32_out = epilogue in (EPI_ATOMIC_F32, EPI_STORE_F32)
    stats_only = epilogue == _lib.EPI_CE_STATS
    if out is None and not stats_only:
        assert epilogue != EPI_ATOMIC_F32, "EPI_ATOMIC_F32 accumulates into an existing buffer"
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if f32_out else torch.bfloat16)
    assert stats_only or (out.dtype == (torch.float32 if f32_out else torch.bfloat16) and tuple(out.shape) == (M, N))
    if epilogue in EPI_DUAL and out2 is None:
        out2 = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    g = _lib.GemmT()
    g.M, g.N, g.K = M, N, K
    g.A, g.lda, g.a_mn_major = a.data_ptr(), _ld(a), int(a_mn)
    g.B, g.ldb, g.b_mn_major = b.data_ptr(), _ld(b), int(b_mn)
    g.epilogue = epilogue
    if not stats_only:
        g.C, g.ldc = out.data_ptr(), _ld(out)
    if ce is not None:
        tgt = ce["target"]
        assert tgt.dtype == torch.int64 and tgt.numel() == M and tgt.is_contiguous()
        g.ce_target = tgt.data_ptr()
        for k in ("partial", "tlogit", "lse", "scale"):
            t = ce.get(k)
            if t is not None:
                assert t.dtype == torch.float32 and t.is_contiguous()
                assert t.numel() == (M * ((N + 127) // 128) * 2 if k == "partial" else M)
                setattr(g, "ce_" + k, t.data_ptr())
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
    g.splits, g.block_n, g.impl, g.cluster = splits, block_n, impl, cluster
    if colsum is not None:
        assert colsum.dtype == torch.float32 and colsum.numel() == N and colsum.is_contiguous()
        g.colsum = colsum.data_ptr()
    if ln is not None:
        gamma, beta, eps, mean, rstd = ln
        assert gamma.dtype == torch.float32 and beta.dtype == torch.float32 and gamma.numel() == N == beta.numel()
        g.ln_gamma, g.ln_beta, g.ln_eps = gamma.data_ptr(), beta.data_ptr(), float(eps)
        for t in (mean, rstd):
            assert t is None or (t.dtype == torch.float32 and t.numel() == M and t.is_contiguous())
        g.ln_mean = mean.data_ptr() if mean is not None else None
        g.ln_rstd = rstd.data_ptr() if rstd is not None else None
    _lib.check(_lib.lib().b200u_gemm(C.byref(g), _lib.stream_ptr()), "b200u_gemm")
    return (out, out2) if epilogue in EPI_DUAL else out


def ce_finish(partial, tlogit):
    """(lse [M], loss [M]) from the EPI_CE_STATS partials."""
    M, nt = partial.shape[0], partial.shape[1]
    lse = torch.empty(M, device=partial.device, dtype=torch.float32)
    loss = torch.empty(M, device=partial.device, dtype=torch.float32)
    _call("b200u_ce_finish", P(partial), P(tlogit), P(lse), P(loss), M, nt)
    return lse, loss


# --------------------------------------------------------------------------------------------
def _dt(t):
    if t.dtype == torch.bfloat16:
        return 0
    if t.dtype == torch.float32:
        return 1
    raise _lib.B200UError("unsupported dtype %s" % t.dtype)


def _drop_ref(drop):
    return C.byref(drop) if drop is not None else None


def _call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args, _lib.stream_ptr()), name)


P = _lib.ptr


def layernorm_fwd(x, gamma, beta, eps, out_dtype=None, drop=None, want_stats=True):
    H = x.shape[-1]
    M = x.numel() // H
    assert x.is_contiguous()
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    _call("b200u_layernorm_fwd", P(x), _dt(x), P(gamma), P(beta), P(y), _dt(y), P(mean), P(rstd), M, H,
          float(eps), _drop_ref(drop))
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, *, want_dx=True, dz=False, dbias=None,
                  drop=None, drop_on_input=False):
    H = x.shape[-1]
    M = x.numel() // H
    assert dy.dtype == torch.bfloat16 and dy.is_contiguous() and x.is_contiguous()
    dx = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if want_dx else None
    dzt = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16) if dz else None
    _call("b200u_layernorm_bwd", P(dy), P(x), _dt(x), P(mean), P(rstd), P(gamma), P(dx), P(dzt),
          P(dgamma), P(dbeta), P(dbias), M, H, _drop_ref(drop), int(drop_on_input))
    return dx, dzt


def colsum_accum(x, out):
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and out.dtype == torch.float32
    _call("b200u_colsum_accum", P(x), x.stride(0), P(out), x.shape[0], x.shape[1])


def cast_f32_to_bf16(x, y):
    assert x.dtype == torch.float32 and y.dtype == torch.bfloat16 and x.numel() == y.numel()
    assert x.is_contiguous() and y.is_contiguous()
    _call("b200u_cast_f32_to_bf16", P(x), P(y), C.c_size_t(x.numel()))
    return y


def gather_rows(txt, img, gather_index):
    B, T, H = txt.shape
    R = img.shape[1]
    L = gather_index.shape[1]
    assert gather_index.dtype == torch.int64 and gather_index.is_contiguous()
    out = torch.empty(B, L, H, device=txt.device, dtype=torch.bfloat16)
    _call("b200u_gather_rows", P(txt), P(img), P(gather_index), P(out), B, T, R, L, H)
    return out


def gather_rows_bwd(dout, gather_index, T, R):
    B, L, H = dout.shape
    dtxt = torch.empty(B, T, H, device=dout.device, dtype=torch.bfloat16)
    dimg = torch.empty(B, R, H, device=dout.device, dtype=torch.bfloat16)
    _call("b200u_gather_rows_bwd", P(dout), P(gather_index), P(dtxt), P(dimg), B, T, R, L, H)
    return dtxt, dimg


def attention_fwd(qkv, mask, B, L, heads, H, drop=None, want_lse=True):
    ctx = torch.empty(B * L, H, device=qkv.device, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, L, device=qkv.device, dtype=torch.float32) if want_lse else None
    _call("b200u_attention_fwd", P(qkv), P(mask), P(ctx), P(lse), B, L, heads, H, _drop_ref(drop))
    return ctx, lse


def attention_bwd(qkv, mask, ctx, dctx, lse, B, L, heads, H, drop=None, dbias_qkv=None):
    """dqkv of the fused attention; dbias_qkv (f32 [3H], optional) += column sums of dqkv."""
    dqkv = torch.empty_like(qkv)
    nbytes = _lib.lib().b200u_attention_bwd_scratch_bytes(B, L, heads)
    scratch = torch.empty(nbytes, device=qkv.device, dtype=torch.uint8)
    if dbias_qkv is not None:
        assert dbias_qkv.dtype == torch.float32 and dbias_qkv.numel() == 3 * H and dbias_qkv.is_contiguous()
    _call("b200u_attention_bwd", P(qkv), P(mask), P(ctx), P(dctx), P(lse), P(dqkv), P(scratch), P(dbias_qkv),
          B, L, heads, H, _drop_ref(drop))
    return dqkv


def counter_add(counter, inc):
    assert counter.dtype == torch.int64
    _call("b200u_counter_add", P(counter), C.c_ulonglong(inc))
