"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports exactly the symbols
include/b200u.h declares, with the arities the ctypes table binds. No compute calls here."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _header_functions():
    hdr = open(os.path.join(ROOT, "include", "b200u.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|long long|size_t|const char\*)\s+(b200u_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", hdr, re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


def test_header_and_ctypes_table_agree():
    from meme_challenge_b200 import _abi
    fns = _header_functions()
    assert len(fns) >= 27
    assert set(fns) == set(_abi.SIGNATURES), set(fns) ^ set(_abi.SIGNATURES)
    for name, nargs in fns.items():
        assert len(_abi.SIGNATURES[name][1]) == nargs, name


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in _header_functions():
        assert hasattr(lib, name), name
    lib.b200u_version.restype = ctypes.c_int
    assert lib.b200u_version() == 100
    lib.b200u_last_error_string.restype = ctypes.c_char_p
    assert isinstance(lib.b200u_last_error_string(), bytes)


def test_library_is_sm100a_native(built_lib):
    """SASS carries tcgen05 MMAs (UTC*MMA), TMEM loads (LDTM) and TMA loads (UTMALDG)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    r = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True)
    assert "sm_100a" in r.stdout
    sass = subprocess.run([cuobjdump, "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass


def test_argument_errors_do_not_need_a_gpu(built_lib):
    """Validation failures return a negative code + message, never raise across the ABI."""
    from meme_challenge_b200 import _lib
    L = _lib.lib()
    g = _lib.GemmT()
    rc = L.b200u_gemm(ctypes.byref(g), None)
    assert rc < 0
    assert b"bad shape" in L.b200u_last_error_string()
    rc = L.b200u_attention_fwd(1, 1, 1, None, 2, 300, 12, 768, None, None)
    assert rc < 0 and b"sequence length" in L.b200u_last_error_string()
