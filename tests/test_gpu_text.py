"""LA4Falcon text parsed on the device (fuz_parse_la4falcon, through the C ABI) against the host parser of the same
library and plain Python: every column, line offsets, blank lines, mixed whitespace, a last line without newline,
lines crossing the 128-byte chunks in every phase, identity notations the kernel hands to the host, rejected lines."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blobs(seed=21, n=700):
    from falcon_unzip_b200 import synth_rr
    s = synth_rr.generate_ovlp(n_reads=n, seed=seed)
    return [("\n".join(v) + "\n").encode() for v in s.las_lines.values()]


def _same(DL, HL):
    assert DL.n == HL.n
    for k in ("q", "t", "len", "qs", "qe", "ql", "ts", "te", "tl", "flags", "off", "llen"):
        assert np.array_equal(DL.a[k], HL.a[k]), k
    assert np.array_equal(DL.file, HL.file)


def test_columns_match_host_parser(eng):
    from falcon_unzip_b200 import la4falcon, ovlp_filter_with_phase as ofp
    blobs = _blobs()
    blobs[0] = b"\n  \n\t\n" + blobs[0].replace(b" ", b"\t ", 7)          # blank lines, mixed whitespace
    blobs[1] = blobs[1][:-1]                                              # (normalised: newline added back)
    _same(la4falcon.DeviceLines(blobs, True), ofp.Lines(blobs))
    # every alignment of line ends relative to the 128-byte chunks
    one = b"000000001 000000002 -3000 99.0 0 0 3000 9000 0 100 3100 8000 overlap\n"
    for pad in range(0, 130, 7):
        blob = b" " * pad + b"\n" + one * 5 + b"\n\n" + one.replace(b"overlap", b"contains") * 3
        _same(la4falcon.DeviceLines([blob], True), ofp.Lines([blob]))
    # empty input, only blank lines
    assert la4falcon.DeviceLines([b""], True).n == 0
    assert la4falcon.DeviceLines([b"\n \n\t\n"], True).n == 0


def test_identity_column_and_tags(eng):
    from falcon_unzip_b200 import la4falcon
    L = lambda idt, tag="overlap": ("000000001 000000002 -3000 %s 0 0 3000 9000 0 100 3100 8000 %s" % (idt, tag)).encode()
    cases = [("99.0", 1), ("89.99", 0), ("90", 1), ("90.00", 1), ("+90.5", 1), ("-95.0", 0), ("0089.9", 0), (".5", 0), ("100.", 1),
             ("9e1", 1), ("8.99e1", 0), ("inf", 1), ("nan", 1), ("-inf", 0),
             ("89.99999999999999999999", 1),            # rounds to 90.0 in double: not < 90
             ("89.9999999999999", 0), ("90.0000000000000000001", 1)]
    blob = b"\n".join(L(c) for c, _ in cases) + b"\n"
    DL = la4falcon.DeviceLines([blob], True)
    assert DL.n == len(cases)
    for (c, want), f in zip(cases, DL.a["flags"].tolist()):
        assert (f & 1) == want and f < 128, c
        assert (0 if float(c) < 90 else 1) == want
    tags = [("overlap", 1), ("contains", 2), ("contained", 3), ("none", 0), ("overlaps", 0), ("contain", 0), ("x", 0)]
    DL = la4falcon.DeviceLines([b"\n".join(L("99.0", t) for t, _ in tags)], True)
    assert [(f >> 1) & 3 for f in DL.a["flags"].tolist()] == [w for _t, w in tags]
    # more than 13 columns: the LAST token is the tag (l[-1])
    DL = la4falcon.DeviceLines([L("99.0", "overlap contained")], True)
    assert (DL.a["flags"][0] >> 1) & 3 == 3


def test_rejected_lines(eng):
    from falcon_unzip_b200 import la4falcon
    from falcon_unzip_b200._lib import FuzError
    good = "000000001 000000002 -3000 99.0 0 0 3000 9000 0 100 3100 8000 overlap"
    for bad in (good.replace("3100", "31x0"), " ".join(good.split()[:11]), good.replace("99.0", "9x"), good.replace("-3000", "-"),
                good.replace("9000", "99999999999")):
        with pytest.raises(ValueError):
            la4falcon.DeviceLines([[good, bad, good]], True)
    with pytest.raises(FuzError):
        la4falcon.DeviceLines([[good.replace("000000001", "1", 1)]], True)
    assert la4falcon.DeviceLines([[good.replace("000000001", "1", 1)]], False).a["q"][0] == 1      # -m path: ids as ints


def test_large_text(eng):
    """~20 MB of text: the multi-CTA scan of the chunk counts."""
    from falcon_unzip_b200 import la4falcon, ovlp_filter_with_phase as ofp
    blobs = _blobs(seed=8, n=2500)
    assert sum(len(b) for b in blobs) > 15_000_000
    _same(la4falcon.DeviceLines(blobs, True), ofp.Lines(blobs))


def test_parser_fuzz(eng):
    """Random lines: signs, leading zeros, 1-4 blanks of several kinds between tokens, extra columns, CRLF, random
    identity notations -- the device columns must equal the host parser's (which is checked against Python)."""
    from falcon_unzip_b200 import la4falcon, ovlp_filter_with_phase as ofp
    rng = np.random.default_rng(99)
    blanks = [" ", "  ", "\t", " \t ", "\v", "\f "]
    tags = ["overlap", "contains", "contained", "none", "overlapx", "c"]
    lines = []
    for _ in range(20000):
        num = lambda lo, hi: ("%+d" if rng.random() < 0.05 else "%d") % int(rng.integers(lo, hi))
        idt = rng.choice(["%.2f" % (80 + 20 * rng.random()), "%d" % int(rng.integers(85, 101)), "0%.3f" % (89.5 + rng.random()),
                          "9.%de1" % int(rng.integers(0, 10)), "%.17f" % (89.9999 + 0.0002 * rng.random())])
        tok = ["%09d" % int(rng.integers(0, 5000)), "%09d" % int(rng.integers(0, 5000)), num(-30000, 1), idt, num(0, 2), num(0, 100),
               num(0, 20000), num(0, 20000), num(0, 2), num(0, 100), num(0, 20000), num(0, 20000), str(rng.choice(tags))]
        if rng.random() < 0.05:
            tok += ["extra", str(rng.choice(tags))]
        line = (str(rng.choice(blanks)) if rng.random() < 0.1 else "") + "".join(t + str(rng.choice(blanks)) for t in tok[:-1]) + tok[-1]
        lines.append(line + ("\r" if rng.random() < 0.05 else "") + ("\n\n" if rng.random() < 0.03 else "\n"))
    blobs = ["".join(lines[:9000]).encode("ascii"), "".join(lines[9000:]).encode("ascii")]
    DL, HL = la4falcon.DeviceLines(blobs, True), ofp.Lines(blobs)
    assert DL.n == HL.n >= 20000
    for k in ("q", "t", "len", "qs", "qe", "ql", "ts", "te", "tl", "flags", "off", "llen"):
        assert np.array_equal(DL.a[k], HL.a[k]), k


def test_capacity_retry_and_empty_filter_input(eng, monkeypatch):
    """Short lines exceed the line-count estimate (capacity retry inside DeviceLines); an overlap filter call whose
    files are all empty returns no text."""
    from falcon_unzip_b200 import la4falcon, ovlp_filter_with_phase as ofp
    blob = b"1 2 -3 9 0 0 1 1 0 0 1 1 x\n" * 5000
    DL = la4falcon.DeviceLines([blob], False)
    assert DL.n == 5000 and DL.a["q"].tolist() == [1] * 5000 and DL.a["len"][0] == 3 and DL.a["flags"][4999] == 0
    monkeypatch.setattr(ofp, "read_las_lines", lambda db, fn: b"")
    ofp.arid2phase.clear()
    ofp.arid2phase.update({"000000001": ("c", "1", "0")})
    assert ofp.run_ovlp_filter(["a.las", "b.las"], "db", 120, 120, 1, 2500, 10) == b""
    assert ofp.filter_stage1(("db", "a.las", 120, 120, 1, 2500)) == ("a.las", [])
