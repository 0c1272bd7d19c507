"""oracle/ovlp_oracle.py (restatement of ovlp_filter_with_phase.py) against the reference's own
source run under Python 3 (oracle/ref_exec.py), on synthetic LA4Falcon -mo sets and hand-made quirk
cases.  Also writes / checks the committed golden fixture used on the GPU box."""
import os

import pytest

from oracle import ovlp_oracle, ref_exec

pytestmark = pytest.mark.skipif(not ref_exec.available(), reason="reference tree not present")

PARAMS = [dict(max_diff=120, max_cov=120, min_cov=1, min_len=2500, bestn=10),
          dict(max_diff=5, max_cov=14, min_cov=2, min_len=2500, bestn=3),
          dict(max_diff=1000, max_cov=1000, min_cov=0, min_len=600, bestn=0)]


def _a2p(rows):
    return {r.split()[0]: tuple(r.split()[1:4]) for r in rows}


@pytest.mark.parametrize("seed", [5, 6])
@pytest.mark.parametrize("pi", range(len(PARAMS)))
def test_oracle_matches_reference(seed, pi):
    from falcon_unzip_b200 import synth_rr
    s = synth_rr.generate_ovlp(n_reads=900, seed=seed)
    p = PARAMS[pi]
    want = ref_exec.run_ovlp_filter(s.las_lines, s.rid_phase_rows, **p)
    got = ovlp_oracle.run_filter(list(s.las_lines.items()), _a2p(s.rid_phase_rows), **p)
    assert got == want
    assert len(want) > 1000


def test_stage_lists_match_reference():
    from falcon_unzip_b200 import synth_rr
    s = synth_rr.generate_ovlp(n_reads=600, seed=9)
    mod = ref_exec.load_ovlp_filter(s.las_lines)
    mod.arid2phase.update(_a2p(s.rid_phase_rows))
    a2p = _a2p(s.rid_phase_rows)
    for fn, lines in s.las_lines.items():
        want1 = mod.filter_stage1(("db", fn, 8, 20, 2, 2500))[1]
        assert ovlp_oracle.stage1(lines, a2p, 8, 20, 2, 2500) == want1
        assert want1[0] is None                       # the run of `None` judged on (0, 0)
        ig = set(want1)
        want2 = mod.filter_stage2(("db", fn, 8, 20, 2, 2500, ig))[1]
        assert ovlp_oracle.stage2(lines, a2p, 2500, ig) == want2
        want3 = mod.filter_stage3(("db", fn, 8, 20, 2, 2500, ig, want2, 4))[1]
        assert ovlp_oracle.stage3(lines, a2p, 2500, ig, want2, 4) == want3


def test_quirk_cases():
    a2p_rows = ["000000001 c 1 0", "000000002 c 1 1", "000000003 c 1 0", "000000004 c -1 0", "000000005 d 1 0",
                "000000007 c 2 1"]
    L = lambda q, t, ln, idt, qs, qe, ql, ts, te, tl, tag: "%09d %09d %d %s 0 %d %d %d 0 %d %d %d %s" % (
        q, t, -ln, idt, qs, qe, ql, ts, te, tl, tag)
    lines = [L(1, 2, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),        # same block, other phase: dropped
             L(1, 3, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(1, 4, 3000, "89.99", 0, 3000, 9000, 100, 3100, 8000, "overlap"),       # idt < 90
             L(1, 4, 3000, "90", 6000, 9000, 9000, 0, 3000, 8000, "overlap"),
             L(1, 5, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),        # other contig
             L(1, 6, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),        # t not in the map
             L(1, 7, 3000, "99.0", 0, 9000, 9000, 100, 9100, 9900, "contained"),      # both ends: 5' only in stage 3
             L(3, 1, 3000, "99.0", 0, 3000, 9000, 100, 3100, 8000, "overlap"),
             L(3, 4, 2500, "99.0", 6500, 9000, 9000, 0, 2500, 2499, "overlap"),       # t shorter than min_len
             L(4, 1, 3000, "99.0", 0, 3000, 8000, 6000, 9000, 9000, "contains"),
             L(4, 3, 3000, "99.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap"),
             L(4, 3, 3000, "99.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap"),        # identical line twice
             L(4, 7, 3000, "98.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap"),
             L(4, 7, 3000, "97.0", 5000, 8000, 8000, 0, 3000, 9000, "overlap")]        # key tie, different text
    las = {"a.las": lines, "b.las": [lines[7], lines[1]]}
    for p in PARAMS + [dict(max_diff=0, max_cov=5, min_cov=1, min_len=2500, bestn=1)]:
        want = ref_exec.run_ovlp_filter(las, a2p_rows, **p)
        assert ovlp_oracle.run_filter(list(las.items()), _a2p(a2p_rows), **p) == want


def test_golden_fixture_is_current():
    """tests/golden/ovlp_small.txt = what the reference prints for generate_ovlp(seed=21) (made here,
    checked on the GPU box where the reference tree does not exist)."""
    from falcon_unzip_b200 import synth_rr
    s = synth_rr.generate_ovlp(n_reads=700, seed=21)
    want = ref_exec.run_ovlp_filter(s.las_lines, s.rid_phase_rows, **PARAMS[1])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ovlp_small.txt")
    if not os.path.exists(path):
        with open(path, "w") as f:
            f.write(want)
    assert open(path).read() == want
