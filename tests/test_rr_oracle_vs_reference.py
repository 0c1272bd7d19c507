"""Pins oracle/rr_oracle.py against the reference's own rr_hctg_track.py source executed
under Python 3 (oracle/ref_exec.load_rr_hctg_track: one .items() patch + the CPython-2
dict / set order emulators of SURVEY.md B.4).  Build container only."""
import os

import pytest

from oracle import ref_exec, rr_oracle

pytestmark = pytest.mark.skipif(not ref_exec.available(), reason="reference tree not mounted")


def write_inputs(rr, d):
    os.makedirs(d, exist_ok=True)
    p = {k: os.path.join(d, k) for k in ("phased", "r2c", "ids", "out")}
    with open(p["phased"], "w") as f:
        f.write("".join(l + "\n" for l in rr.phased_reads))
    with open(p["r2c"], "w") as f:
        f.write("".join(l + "\n" for l in rr.read_to_contig_map))
    with open(p["ids"], "w") as f:
        f.write(rr.rawread_ids)
    return p


@pytest.mark.parametrize("seed,bestn,n_files", [(4, 40, 3), (5, 5, 2), (6, 2, 4), (7, 40, 1)])
def test_rr_oracle_matches_reference(seed, bestn, n_files, tmp_path):
    from falcon_unzip_b200 import synth_rr
    rr = synth_rr.generate_rr(n_reads=1200, n_ctg=3, ctg_len=100_000, n_files=n_files, seed=seed)
    p = write_inputs(rr, str(tmp_path))
    ref_exec.run_rr_track(rr.las_lines, p["phased"], p["r2c"], p["ids"], p["out"], min_len=2500, bestn=bestn)
    want = open(p["out"]).read()
    got = rr_oracle.run_track_reads(rr.las_lines, rr.phased_reads, rr.read_to_contig_map, rr.rawread_ids, 2500, bestn)
    assert len(want.splitlines()) > 500
    # order-free content first: rows without the rank column (ties in score are ranked by CPython-2 dict order)
    strip = lambda t: sorted(" ".join(l.split()[:3] + l.split()[4:]) for l in t.splitlines())
    assert strip(want) == strip(got)
    assert want == got


def test_e19_e20_known_answers(tmp_path):
    """SURVEY.md Appendix E19-E21 style cases, by hand."""
    names = ["r%d" % i for i in range(12)]
    r2c = ["%09d %09d %s %s" % (i, rid, names[rid], c) for i, (rid, c) in enumerate(
        [(1, "000000F"), (2, "000000F"), (2, "000000F_001"), (3, "000000F_001"), (4, "000000F"), (5, "000000F_001")])]
    phased = ["4 000000F 1 0 5 0 r4", "5 000000F 1 1 0 5 r5", "6 000000F 1 0 5 0 r6", "7 000000F -1 0 5 0 r7"]

    def line(q, t, ln, tl=8000):
        return "%09d %09d %d 99.0 0 0 %d 9000 0 100 %d %d overlap" % (q, t, -ln, ln, 100 + ln, tl)
    las = {"a.las": [line(1, 0, 5000), line(2, 0, 7000), line(3, 0, 7000),      # bestn 2 keeps the two 7000s
                     line(1, 8, 6000, 2499), line(1, 9, 6000, 2500),             # t_l 2499 dropped / 2500 kept
                     line(10, 9, 9000),                                          # q not in read_to_contig_map
                     line(5, 4, 3000),                                           # same ctg+block, other phase: dropped
                     line(5, 6, 3000), line(4, 6, 3100),                         # t=6 phase 0: q=5 dropped, q=4 kept
                     line(5, 7, 3200),                                           # t block -1: kept
                     line(5, 11, 3300)],                                         # t unphased: kept
           "b.las": [line(4, 0, 6500)]}                                          # second file: merged, still two 7000s
    p = write_inputs(type("RR", (), dict(phased_reads=phased, read_to_contig_map=r2c,
                                         rawread_ids="\n".join(names) + "\n"))(), str(tmp_path))
    ref_exec.run_rr_track(las, p["phased"], p["r2c"], p["ids"], p["out"], min_len=2500, bestn=2)
    want = open(p["out"]).read()
    got = rr_oracle.run_track_reads(las, phased, r2c, "\n".join(names) + "\n", 2500, 2)
    assert want == got
    rows = {tuple(l.split()[:2]): l.split() for l in got.splitlines()}
    assert rows[("000000000", "000000F_001")][2:5] == ["2", "0", "-14000"]      # E20 verified output
    assert rows[("000000000", "000000F")][2:5] == ["1", "1", "-7000"]
    assert not any(l.startswith("000000008") for l in got.splitlines())
    assert any(l.startswith("000000009 000000F 1 0 -6000") for l in got.splitlines())
    assert not any(l.startswith("000000004") for l in got.splitlines())
    assert [l for l in got.splitlines() if l.startswith("000000006")] == ["000000006 000000F 1 0 -3100 0"]
