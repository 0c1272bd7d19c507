"""Overlap filter with phase on the device, through the C ABI (fuz_ovlp_filter): the fused
call and the three per-file stage functions against the restated oracle (oracle/ovlp_oracle.py, itself pinned
to the reference's source) and against the committed golden fixture; quirk cases; the CLI."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PARAMS = [dict(max_diff=120, max_cov=120, min_cov=1, min_len=2500, bestn=10),
          dict(max_diff=5, max_cov=14, min_cov=2, min_len=2500, bestn=3),
          dict(max_diff=1000, max_cov=1000, min_cov=0, min_len=600, bestn=0)]


def _a2p(rows):
    return {r.split()[0]: tuple(r.split()[1:4]) for r in rows}


@pytest.fixture
def ofp(eng, monkeypatch):
    from falcon_unzip_b200 import ovlp_filter_with_phase as m
    store = {}
    monkeypatch.setattr(m, "read_las_lines", lambda db, fn: ("\n".join(store[fn]) + "\n").encode() if store[fn] else b"")
    m.arid2phase.clear()
    m._store = store
    return m


def _load(ofp, las, rows):
    ofp._store.clear(); ofp._store.update(las)
    ofp.arid2phase.clear(); ofp.arid2phase.update(_a2p(rows))


@pytest.mark.parametrize("seed", [5, 6])
@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_fused_filter_matches_oracle(ofp, seed, pi):
    from falcon_unzip_b200 import synth_rr
    from oracle import ovlp_oracle
    s = synth_rr.generate_ovlp(n_reads=900, seed=seed)
    _load(ofp, s.las_lines, s.rid_phase_rows)
    p = PARAMS[pi]
    want = ovlp_oracle.run_filter(list(s.las_lines.items()), _a2p(s.rid_phase_rows), **p)
    got = ofp.run_ovlp_filter(list(s.las_lines), "db", p["max_diff"], p["max_cov"], p["min_cov"], p["min_len"], p["bestn"]).decode()
    assert got == want and len(want) > 1000


def test_fused_filter_large(ofp):
    """~280 k lines: the multi-CTA scan (k_scan_wide) and reads with ~100 candidates per end."""
    from falcon_unzip_b200 import synth_rr
    from oracle import ovlp_oracle
    s = synth_rr.generate_ovlp(n_reads=2500, seed=8)
    assert sum(len(v) for v in s.las_lines.values()) > 200000
    _load(ofp, s.las_lines, s.rid_phase_rows)
    for p in (PARAMS[0], dict(max_diff=30, max_cov=70, min_cov=3, min_len=2500, bestn=5)):
        want = ovlp_oracle.run_filter(list(s.las_lines.items()), _a2p(s.rid_phase_rows), **p)
        got = ofp.run_ovlp_filter(list(s.las_lines), "db", p["max_diff"], p["max_cov"], p["min_cov"], p["min_len"], p["bestn"]).decode()
        assert got == want and len(want) > 10000, p


def test_golden_fixture(ofp):
    from falcon_unzip_b200 import synth_rr
    s = synth_rr.generate_ovlp(n_reads=700, seed=21)
    _load(ofp, s.las_lines, s.rid_phase_rows)
    got = ofp.run_ovlp_filter(list(s.las_lines), "db", 5, 14, 2, 2500, 3).decode()
    want = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ovlp_small.txt")).read()
    assert got == want


def test_stage_functions_match_oracle(ofp):
    from falcon_unzip_b200 import synth_rr
    from oracle import ovlp_oracle
    s = synth_rr.generate_ovlp(n_reads=600, seed=9)
    _load(ofp, s.las_lines, s.rid_phase_rows)
    a2p = _a2p(s.rid_phase_rows)
    for fn, lines in s.las_lines.items():
        want1 = ovlp_oracle.stage1(lines, a2p, 8, 20, 2, 2500)
        assert ofp.filter_stage1(("db", fn, 8, 20, 2, 2500)) == (fn, want1)
        ig = set(want1)
        want2 = ovlp_oracle.stage2(lines, a2p, 2500, ig)
        assert ofp.filter_stage2(("db", fn, 8, 20, 2, 2500, ig)) == (fn, want2)
        want3 = ovlp_oracle.stage3(lines, a2p, 2500, ig, want2, 4)
        assert ofp.filter_stage3(("db", fn, 8, 20, 2, 2500, ig, want2, 4)) == (fn, want3)
    # min_cov 0: the run of `None` passes the verdict and is not reported
    fn, lines = next(iter(s.las_lines.items()))
    assert ofp.filter_stage1(("db", fn, 1000, 1000, 0, 2500)) == (fn, ovlp_oracle.stage1(lines, a2p, 1000, 1000, 0, 2500))


def test_quirk_cases(ofp):
    from oracle import ovlp_oracle
    rows = ["000000001 c 1 0", "000000002 c 1 1", "000000003 c 1 0", "000000004 c -1 0", "000000005 d 1 0", "000000007 c 2 1"]
    L = lambda q, t, ln, idt, qs, qe, ql, ts, te, tl, tag: "%09d %09d %d %s 0 %d %d %d 0 %d %d %d %s" % (
        q, t, -ln, idt, qs, qe, ql, ts, te, tl, tag)
    lines = [L(1, 2, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(1, 3, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(1, 4, 3000, "89.99", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(1, 4, 3000, "90", 6000, 9000, 9000, 0, 3000, 8000, "overlap"),
             L(1, 5, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(1, 6, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(1, 7, 3000, "99.0", 0, 9000, 9000, 100, 9100, 9900, "contained"),
             L(3, 1, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(3, 4, 2500, "99.0", 6500, 9000, 9000, 0, 2500, 2499, "overlap"),
             L(4, 1, 3000, "99.0", 0, 3000, 8000, 6000, 9000, 9000, "contains"),
             L(4, 3, 3000, "99.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap"),
             L(4, 3, 3000, "99.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap"),
             L(4, 7, 3000, "98.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap"),
             L(4, 7, 3000, "97.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap")]
    las = {"a.las": lines, "b.las": [lines[7], lines[1]], "c.las": []}
    _load(ofp, las, rows)
    for p in PARAMS + [dict(max_diff=0, max_cov=5, min_cov=1, min_len=2500, bestn=1)]:
        want = ovlp_oracle.run_filter(list(las.items()), _a2p(rows), **p)
        got = ofp.run_ovlp_filter(list(las), "db", p["max_diff"], p["max_cov"], p["min_cov"], p["min_len"], p["bestn"]).decode()
        assert got == want, p


def test_cli(ofp, tmp_path, capsys):
    from falcon_unzip_b200 import synth_rr
    from oracle import ovlp_oracle
    s = synth_rr.generate_ovlp(n_reads=500, seed=3)
    _load(ofp, s.las_lines, s.rid_phase_rows)
    (tmp_path / "las.fofn").write_text("\n".join(s.las_lines) + "\n")
    (tmp_path / "rid_to_phase.all").write_text("\n".join(s.rid_phase_rows) + "\n")
    ofp.main(["fc_ovlp_filter_with_phase.py", "--fofn", str(tmp_path / "las.fofn"), "--max_diff", "120", "--max_cov", "120",
              "--min_cov", "1", "--n_core", "12", "--min_len", "2500", "--db", "preads.db", "--rid_phase_map",
              str(tmp_path / "rid_to_phase.all")])
    want = ovlp_oracle.run_filter(list(s.las_lines.items()), _a2p(s.rid_phase_rows), 120, 120, 1, 2500, 10)
    assert capsys.readouterr().out == want
