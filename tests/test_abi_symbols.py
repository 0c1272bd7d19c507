"""The C-ABI library loads without a GPU and exports every symbol include/fuz.h declares (and
nothing is bound in _lib.py that the header does not declare).  No compute calls here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fuz.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fuz_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from falcon_unzip_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libfuz.so does not export %s" % n


def test_python_binding_matches_header():
    from falcon_unzip_b200 import _lib
    bound = sorted(n for n, _r, _a in _lib.SYMBOLS)
    assert bound == header_symbols()
    assert _lib.lib().fuz_version() == 1 and _lib.lib().fuz_tile_size() == 8192


def test_host_helpers_work_without_gpu():
    import numpy as np
    from falcon_unzip_b200 import _lib
    text = b"000000011 000000001 -7000 99.0 0 0 7000 9000 0 100 7100 8000 overlap\n\n000000012 000000001 -5 98.5 0 0 5 9000 1 100 105 2499 overlap"
    q, t, ln, tl = (np.zeros(4, np.int32) for _ in range(4))
    n = _lib.lib().fuz_host_parse_la4falcon(text, len(text), 4, q.ctypes.data, t.ctypes.data, ln.ctypes.data, tl.ctypes.data)
    assert n == 2 and q[:2].tolist() == [11, 12] and ln[:2].tolist() == [7000, 5] and tl[:2].tolist() == [8000, 2499]
    assert _lib.lib().fuz_host_parse_la4falcon(b"1 2 x\n", 6, 4, q.ctypes.data, t.ctypes.data, ln.ctypes.data, tl.ctypes.data) == -1


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    from falcon_unzip_b200 import engine
    with pytest.raises(RuntimeError):
        engine.Engine(0)
    ctx = ctypes.c_void_p()
    from falcon_unzip_b200 import _lib
    assert _lib.lib().fuz_ctx_create(0, ctypes.byref(ctx)) != 0
    assert b"no CPU fallback" in _lib.lib().fuz_last_error(None) or b"CUDA" in _lib.lib().fuz_last_error(None)
