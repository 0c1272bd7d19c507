"""Host helpers of the overlap filter (no GPU): the LA4Falcon -mo parser and the output formatter of
libfuz.so against plain Python on the same text; the restated oracle against the committed golden
fixture (made from the reference's own source by tests/test_ovlp_oracle_vs_reference.py)."""
import os

import numpy as np
import pytest


def _set(seed=21, n=700):
    from falcon_unzip_b200 import synth_rr
    return synth_rr.generate_ovlp(n_reads=n, seed=seed)


def test_parser_matches_python_split():
    from falcon_unzip_b200 import ovlp_filter_with_phase as ofp
    s = _set()
    blobs = ["\n".join(v).encode() + b"\n" for v in s.las_lines.values()]
    blobs[0] = b"\n  \n" + blobs[0].replace(b" ", b"\t ", 7)          # blank lines, mixed whitespace
    L = ofp.Lines(blobs)
    rows = [l.split() for b in blobs for l in b.decode().split("\n") if l.strip()]
    assert L.n == len(rows)
    tag = {"overlap": 1, "contains": 2, "contained": 3}
    for i in list(range(0, L.n, 97)) + [0, L.n - 1]:
        r = rows[i]
        assert (L.a["q"][i], L.a["t"][i], L.a["len"][i]) == (int(r[0]), int(r[1]), -int(r[2]))
        assert [L.a[k][i] for k in ("qs", "qe", "ql", "ts", "te", "tl")] == [int(r[k]) for k in (5, 6, 7, 9, 10, 11)]
        assert L.a["flags"][i] == (0 if float(r[3]) < 90 else 1) | (tag.get(r[-1], 0) << 1)
        assert L.tokens(i) == r
    counts = np.cumsum([sum(1 for l in b.decode().split("\n") if l.strip()) for b in blobs])
    assert np.array_equal(L.file, np.searchsorted(counts, np.arange(L.n), side="right"))


def test_parser_rejects_what_the_reference_rejects():
    from falcon_unzip_b200 import ovlp_filter_with_phase as ofp
    from falcon_unzip_b200._lib import FuzError
    good = "000000001 000000002 -3000 99.0 0 0 3000 9000 0 100 3100 8000 overlap"
    assert ofp.Lines([[good]]).n == 1
    with pytest.raises(ValueError):
        ofp.Lines([[good.replace("99.0", "9x")]])
    with pytest.raises(ValueError):
        ofp.Lines([[" ".join(good.split()[:11])]])
    with pytest.raises(FuzError):
        ofp.Lines([[good.replace("000000001", "1", 1)]])


def test_formatter_matches_python_join():
    from falcon_unzip_b200 import ovlp_filter_with_phase as ofp
    s = _set()
    L = ofp.Lines(["\n".join(v).encode() + b"\n" for v in s.las_lines.values()])
    a2p = {r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows}
    tab = ofp.PhaseTable(a2p, n_reads=s.n_reads)
    sel = np.asarray([i for i in range(0, L.n, 3) if "%09d" % L.a["q"][i] in a2p and "%09d" % L.a["t"][i] in a2p][:5000], np.int64)
    got = ofp._format(L, tab, sel).decode()
    want = "".join(" ".join(L.tokens(i) + [".".join(a2p["%09d" % L.a["q"][i]]), ".".join(a2p["%09d" % L.a["t"][i]])]) + "\n" for i in sel)
    assert got == want


def test_oracle_matches_golden():
    from oracle import ovlp_oracle
    s = _set()
    a2p = {r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows}
    got = ovlp_oracle.run_filter(list(s.las_lines.items()), a2p, max_diff=5, max_cov=14, min_cov=2, min_len=2500, bestn=3)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ovlp_small.txt")
    assert got == open(path).read() and len(got) > 1000
