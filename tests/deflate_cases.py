"""Raw deflate streams covering every block type and table shape (shared by the CPU model test
and the GPU kernel test): (name, payload, zlib level, zlib strategy)."""
import os
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def payloads():
    rng = np.random.default_rng(7)
    text = open(os.path.join(ROOT, "SURVEY.md"), "rb").read()
    skew = np.array([2.0 ** -(i / 8) for i in range(256)]); skew /= skew.sum()
    steep = np.array([2.0 ** -(i / 2) for i in range(40)]); steep /= steep.sum()
    return [
        ("empty", b""),
        ("one", b"a"),
        ("ff_run", b"\xff" * 65280),                                            # QUAL of a PacBio BAM: distance-1 matches
        ("random", bytes(rng.integers(0, 256, 65280, dtype=np.uint8))),          # packed SEQ: literals only
        ("acgt2", bytes(rng.integers(0, 4, 65280, dtype=np.uint8))),
        ("repeat", (b"ACGT" * 1000 + bytes(rng.integers(0, 256, 3000, dtype=np.uint8))) * 5),
        ("period200", bytes(rng.integers(0, 256, 200, dtype=np.uint8)) * 300),
        ("text", text[:65280]),
        ("skew", bytes(rng.choice(256, 60000, p=skew).astype(np.uint8))),        # code lengths up to 15
        ("steep", bytes(rng.choice(40, 60000, p=steep).astype(np.uint8))),
        ("short", text[1000:1037]),
    ]


def streams():
    out = []
    for name, data in payloads():
        for level in (0, 1, 6, 9):
            for strat in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
                c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strat)
                out.append(("%s-l%d-s%d" % (name, level, strat), data, c.compress(data) + c.flush()))
    return out


def corrupt_streams():
    """Streams a decoder must reject (or at least survive): (name, stream, expected size)."""
    rng = np.random.default_rng(11)
    good = zlib.compressobj(6, zlib.DEFLATED, -15)
    data = open(os.path.join(ROOT, "SURVEY.md"), "rb").read()[:30000]
    comp = good.compress(data) + good.flush()
    out = [("truncated", comp[:len(comp) // 2], len(data)),
           ("reserved_type", bytes([0x07]) + comp[1:], len(data)),
           ("stored_len_mismatch", bytes([0x01, 0x05, 0x00, 0x00, 0x00]) + b"hello", 5),
           ("distance_too_far", zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED).compress(b"") + bytes([0x73, 0x04, 0x02, 0x00]), 20),
           ("wrong_size", comp, len(data) - 7)]
    for k in range(8):
        b = bytearray(comp)
        for _ in range(4):
            b[int(rng.integers(8, len(b)))] ^= 1 << int(rng.integers(0, 8))
        out.append(("bitflips%d" % k, bytes(b), len(data)))
    for k in range(4):
        out.append(("noise%d" % k, bytes(rng.integers(0, 256, 4000, dtype=np.uint8)), 30000))
    return out
