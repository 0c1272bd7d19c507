"""The DEFLATE decoder core shared with the CUDA kernel (fuz_inflate_core.h), instantiated with a
scalar IO policy (oracle/inflate_model.cpp), against zlib: every block type, table shape and
strategy, byte offsets 0..3 of the stream start, a small lookup table to force the canonical
search, and corrupt streams (must fail or at least terminate inside the buffers)."""
import os
import zlib

import pytest

from oracle import inflate_model
import deflate_cases


@pytest.fixture(scope="module")
def libs():
    small = os.path.join(os.path.dirname(inflate_model._SO), "libinflate_model_l4.so")
    return inflate_model.load(), inflate_model.load(inflate_model.build(defines=("FUZ_INF_LBITS=4",), out=small))


def test_streams_match_zlib(libs):
    n = 0
    for name, data, comp in deflate_cases.streams():
        assert zlib.decompress(comp, -15) == data
        for lib in libs:
            for pre in (0, 1, 2, 3):
                buf = b"\x5a" * pre + comp + b"\xa5" * 3
                rc, got = inflate_model.inflate(buf, pre, len(comp), len(data), lib)
                assert rc == 0 and got == data, (name, pre)
                n += 1
    assert n > 1000


def test_corrupt_streams_terminate(libs):
    for name, comp, size in deflate_cases.corrupt_streams():
        try:
            want = zlib.decompress(comp, -15)
        except zlib.error:
            want = None
        for lib in libs:
            rc, got = inflate_model.inflate(comp, 0, len(comp), size, lib)
            if want is not None and len(want) == size:
                assert rc == 0 and got == want, name      # a bit flip zlib also decodes (e.g. inside a literal)
            else:
                assert rc != 0 or len(got) != size, name
