"""ctypes signatures of every symbol include/b200u.h declares (kept in the same order).

tests/test_abi.py checks this table against the header and against the built library, so a
symbol added to one and not the others fails the CPU suite.
"""
import ctypes as C

_p = C.c_void_p
_i = C.c_int
_f = C.c_float

SIGNATURES = {
    "b200u_last_error_string": (C.c_char_p, []),
    "b200u_version": (_i, []),
    "b200u_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "b200u_gemm": (_i, [_p, _p]),
}
