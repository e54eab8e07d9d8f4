"""ctypes binding of libb200u.so (the C-ABI in include/b200u.h).

There is deliberately no CPU or PyTorch-eager fallback: if the shared library is missing or a
tensor is not a CUDA tensor on an sm_100 device the call raises (BASELINE.json north_star:
"no CPU fallback").
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200u.so")

_lib = None


class B200UError(RuntimeError):
    pass


class DropoutT(C.Structure):
    _fields_ = [("seed_ptr", C.c_void_p), ("stream", C.c_uint32), ("p", C.c_float)]


class GemmT(C.Structure):
    _fields_ = [
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("A", C.c_void_p), ("lda", C.c_int), ("a_mn_major", C.c_int),
        ("B", C.c_void_p), ("ldb", C.c_int), ("b_mn_major", C.c_int),
        ("epilogue", C.c_int),
        ("C", C.c_void_p), ("ldc", C.c_int),
        ("C2", C.c_void_p), ("ldc2", C.c_int),
        ("bias", C.c_void_p),
        ("R", C.c_void_p), ("ldr", C.c_int),
        ("drop", DropoutT),
        ("splits", C.c_int), ("block_n", C.c_int), ("impl", C.c_int), ("cluster", C.c_int),
        ("colsum", C.c_void_p),
        ("ln_gamma", C.c_void_p), ("ln_beta", C.c_void_p), ("ln_mean", C.c_void_p), ("ln_rstd", C.c_void_p),
        ("ln_eps", C.c_float),
        ("ce_target", C.c_void_p), ("ce_partial", C.c_void_p), ("ce_tlogit", C.c_void_p),
        ("ce_lse", C.c_void_p), ("ce_scale", C.c_void_p),
    ]


(EPI_STORE, EPI_BIAS_GELU, EPI_BIAS_DROP_RES, EPI_ADD, EPI_DGELU, EPI_ATOMIC_F32, EPI_STORE_F32,
 EPI_BIAS_GELU_DG, EPI_MUL, EPI_BIAS_DROP_RES_LN, EPI_CE_STATS, EPI_CE_GRAD) = range(12)
EPI_COUNT = 12
EPI_HAS_BIAS = (EPI_STORE, EPI_BIAS_GELU, EPI_BIAS_DROP_RES, EPI_STORE_F32, EPI_BIAS_GELU_DG, EPI_BIAS_DROP_RES_LN)
EPI_HAS_RES = (EPI_BIAS_DROP_RES, EPI_ADD, EPI_DGELU, EPI_MUL, EPI_BIAS_DROP_RES_LN)
EPI_DUAL = (EPI_BIAS_GELU, EPI_BIAS_GELU_DG, EPI_BIAS_DROP_RES_LN)


def lib():
    """Load libb200u.so (built in-tree by meme_challenge_b200.build); raise loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200UError(
                "libb200u.so not found at %s — run `python -m meme_challenge_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.b200u_last_error_string.restype = C.c_char_p
        _declare(_lib)
    return _lib


def _declare(L):
    from . import _abi
    for name, (restype, argtypes) in _abi.SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = restype
        fn.argtypes = argtypes


def check(rc, what=""):
    if rc != 0:
        msg = lib().b200u_last_error_string().decode("utf-8", "replace")
        raise B200UError("%s failed (%d): %s" % (what or "b200u call", rc, msg))


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL). Refuses non-CUDA tensors: no CPU path."""
    if t is None:
        return None
    if not t.is_cuda:
        raise B200UError("b200u kernels need CUDA tensors (got %s); there is no CPU fallback" % t.device)
    return C.c_void_p(t.data_ptr())


def dropout_t(seed, stream_id, p):
    d = DropoutT()
    d.seed_ptr = seed.data_ptr() if (seed is not None and p > 0.0) else None
    d.stream = int(stream_id) & 0xFFFFFFFF
    d.p = float(p)
    return d
