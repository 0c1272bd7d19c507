"""TEST INFRASTRUCTURE ONLY (oracle/): builds and binds oracle/inflate_model.cpp, the scalar
instantiation of the DEFLATE decoder the CUDA kernel runs (falcon_unzip_b200/csrc/
fuz_inflate_core.h).  zlib is the oracle of the format; this model lets the table construction
and block parsing of the shared core be checked against it on the CPU.  Never imported by
falcon_unzip_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "inflate_model.cpp")
_CORE = os.path.join(_HERE, "..", "falcon_unzip_b200", "csrc", "fuz_inflate_core.h")
_SO = os.path.join(_HERE, "_build", "libinflate_model.so")
_lib = None


def build(force: bool = False, defines=(), out: str | None = None) -> str:
    so = out or _SO
    newest = max(os.path.getmtime(_SRC), os.path.getmtime(_CORE))
    if force or not os.path.exists(so) or os.path.getmtime(so) < newest:
        os.makedirs(os.path.dirname(so), exist_ok=True)
        tmp = so + ".tmp%d" % os.getpid()
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", *("-D" + d for d in defines), "-o", tmp, _SRC])
        os.replace(tmp, so)
    return so


def load(path: str | None = None):
    lib = C.CDLL(path or build())
    lib.inflate_model.restype = C.c_int
    lib.inflate_model.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
    return lib


def inflate(data: bytes, off: int, n: int, size: int, lib=None):
    """-> (FUZ_INF_* code, bytes produced) for the raw deflate stream data[off:off+n]."""
    global _lib
    if lib is None:
        if _lib is None:
            _lib = load()
        lib = _lib
    out = C.create_string_buffer(size + 8)
    n_out = C.c_int64(0)
    rc = lib.inflate_model(data, len(data), off, n, out, size, C.byref(n_out))
    return rc, out.raw[:n_out.value]
