"""ctypes binding of tests/gen/psb_gen.c -- the Appendix-F generator in threaded C (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "gen", "psb_gen.c")
_OUT = os.path.join(_HERE, "gen", "libpsb_gen.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
            subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-pthread", "-o", _OUT, _SRC], check=True)
        _lib = C.CDLL(_OUT)
    return _lib


def _threads():
    return os.cpu_count() or 1


def random(seed, stream, n, protein):
    out = np.empty(n, dtype=np.uint8)
    lib().psbg_random(C.c_uint64(seed), C.c_uint64(stream), C.c_int64(n), C.c_int(int(protein)), out.ctypes.data_as(C.c_void_p), _threads())
    return out


def substitute(src, seed, stream, rate, protein):
    src = np.ascontiguousarray(src, dtype=np.uint8)
    out = np.empty_like(src)
    lib().psbg_substitute(src.ctypes.data_as(C.c_void_p), C.c_int64(len(src)), C.c_uint64(seed), C.c_uint64(stream),
                          C.c_int(int(round(rate * 10000))), C.c_int(int(protein)), out.ctypes.data_as(C.c_void_p), _threads())
    return out


def starts(seed, stream, n, mod):
    out = np.empty(n, dtype=np.int64)
    lib().psbg_starts(C.c_uint64(seed), C.c_uint64(stream), C.c_int64(n), C.c_uint64(mod), out.ctypes.data_as(C.c_void_p))
    return out


def gather(src, npairs, win_len, st, read_len):
    out = np.empty(npairs * read_len, dtype=np.uint8)
    st = np.ascontiguousarray(st, dtype=np.int64)
    lib().psbg_gather(src.ctypes.data_as(C.c_void_p), C.c_int64(npairs), C.c_int64(win_len), st.ctypes.data_as(C.c_void_p),
                      C.c_int64(read_len), out.ctypes.data_as(C.c_void_p), _threads())
    return out
