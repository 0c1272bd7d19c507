"""GPU parity of the raw-read -> haplotig tracking (fuz_rr_track through the reference-signature
functions of falcon_unzip_b200.rr_hctg_track) against the CPU oracle: byte-identical
rawread_to_contigs, identical heapq arrays from tr_stage1."""
import os

import pytest

pytestmark = pytest.mark.gpu


def _write(rr, d):
    os.makedirs(d, exist_ok=True)
    p = {k: os.path.join(d, k) for k in ("phased", "r2c", "ids", "out")}
    with open(p["phased"], "w") as f:
        f.write("".join(l + "\n" for l in rr.phased_reads))
    with open(p["r2c"], "w") as f:
        f.write("".join(l + "\n" for l in rr.read_to_contig_map))
    with open(p["ids"], "w") as f:
        f.write(rr.rawread_ids)
    return p


@pytest.mark.parametrize("seed,bestn,n_files,n_reads", [(4, 40, 3, 1500), (5, 5, 2, 1200), (6, 2, 4, 1200), (7, 40, 1, 800),
                                                        (8, 1, 3, 600), (9, 40, 5, 4000)])
def test_run_track_reads_bytes_match_oracle(eng, seed, bestn, n_files, n_reads, tmp_path, monkeypatch):
    from falcon_unzip_b200 import rr_hctg_track, synth_rr
    from oracle import rr_oracle
    rr = synth_rr.generate_rr(n_reads=n_reads, n_ctg=3, ctg_len=100_000, n_files=n_files, seed=seed)
    p = _write(rr, str(tmp_path))
    # the caller's file order is kept (the reference walks file_list as given, rr_hctg_track.py:88-98): reversed names here
    order = list(reversed(sorted(rr.las_lines)))
    want = rr_oracle.run_track_reads(rr.las_lines, rr.phased_reads, rr.read_to_contig_map, rr.rawread_ids, 2500, bestn,
                                     file_order=order)
    monkeypatch.setattr(rr_hctg_track, "read_las_lines", lambda db_fn, fn: iter(rr.las_lines[fn]))
    rr_hctg_track.run_track_reads(None, p["phased"], p["r2c"], p["ids"], order, 2500, bestn, "raw_reads.db", p["out"])
    got = open(p["out"]).read()
    assert len(want.splitlines()) > 200
    assert want == got


def test_tr_stage1_heaps_match_oracle(eng):
    from falcon_unzip_b200 import rr_hctg_track, synth_rr
    from oracle import rr_oracle
    rr = synth_rr.generate_rr(n_reads=900, n_ctg=2, ctg_len=80_000, n_files=1, seed=12)
    lines = next(iter(rr.las_lines.values()))
    rid_to_ctg_o = rr_oracle.get_rid_to_ctg(rr.read_to_contig_map)
    rid_to_phase = rr_oracle.phase_table(rr.phased_reads, rr.rawread_ids)
    want = rr_oracle.tr_stage1(lines, 2500, 7, rid_to_ctg_o, rid_to_phase)
    got = rr_hctg_track.tr_stage1(lambda: iter(lines), 2500, 7, rid_to_ctg_o, rid_to_phase)
    assert list(want) == list(got)                    # same targets, same first-appearance order
    assert want == got                                # same heapq ARRAY order


def test_cli_and_known_answers(eng, tmp_path, monkeypatch):
    """fc_rr_hctg_track.py command line + the Appendix E19 / E20 cases."""
    from falcon_unzip_b200 import rr_hctg_track
    from oracle import rr_oracle
    names = ["r%d" % i for i in range(12)]
    r2c = ["%09d %09d %s %s" % (i, rid, names[rid], c) for i, (rid, c) in enumerate(
        [(1, "000000F"), (2, "000000F"), (2, "000000F_001"), (3, "000000F_001"), (4, "000000F"), (5, "000000F_001")])]
    phased = ["4 000000F 1 0 5 0 r4", "5 000000F 1 1 0 5 r5", "6 000000F 1 0 5 0 r6", "7 000000F -1 0 5 0 r7"]

    def line(q, t, ln, tl=8000):
        return "%09d %09d %d 99.0 0 0 %d 9000 0 100 %d %d overlap" % (q, t, -ln, ln, 100 + ln, tl)
    las = {"0-rawreads/m1/raw_reads.1.las": [line(1, 0, 5000), line(2, 0, 7000), line(3, 0, 7000), line(1, 8, 6000, 2499),
                                             line(1, 9, 6000, 2500), line(10, 9, 9000), line(5, 4, 3000), line(5, 6, 3000),
                                             line(4, 6, 3100), line(5, 7, 3200), line(5, 11, 3300)],
           "0-rawreads/m2/raw_reads.2.las": [line(4, 0, 6500)]}
    rr = type("RR", (), dict(phased_reads=phased, read_to_contig_map=r2c, rawread_ids="\n".join(names) + "\n"))()
    p = _write(rr, str(tmp_path))
    want = rr_oracle.run_track_reads(las, phased, r2c, rr.rawread_ids, 2500, 2)
    monkeypatch.chdir(tmp_path)
    for fn in las:
        os.makedirs(os.path.dirname(fn), exist_ok=True)
        open(fn, "w").close()
    monkeypatch.setattr(rr_hctg_track, "read_las_lines",
                        lambda db_fn, fn: iter(las[os.path.relpath(fn, str(tmp_path))]))
    rr_hctg_track.main(["fc_rr_hctg_track.py", "--phased-read-file", p["phased"], "--read-to-contig-map", p["r2c"],
                        "--rawread-ids", p["ids"], "--output", p["out"], "--bestn", "2", "--n-core", "4"])
    got = open(p["out"]).read()
    assert want == got
    assert "000000000 000000F_001 2 0 -14000 0\n" in got and "000000000 000000F 1 1 -7000 0\n" in got


def test_target_voting_for_more_contigs_than_the_fast_table(eng, tmp_path, monkeypatch):
    """A repeat read whose 40 kept a-reads map to 3 contigs each (120 distinct contigs): the per-thread vote table of
    k_rr_vote holds 64; the unbounded path must give the reference's rows (it has no such limit, rr_hctg_track.py:113-138)."""
    from falcon_unzip_b200 import rr_hctg_track
    from oracle import rr_oracle
    n = 42
    names = ["r%d" % i for i in range(n)]
    r2c, pid = [], 0
    for q in range(1, 41):
        for k in range(3):
            r2c.append("%09d %09d %s %06dF" % (pid, q, names[q], 3 * q + k))
            pid += 1
    r2c.append("%09d %09d %s %06dF" % (pid, 0, names[0], 5))
    las = {"0-rawreads/m1/raw_reads.1.las": ["%09d %09d %d 99.0 0 0 %d 9000 0 100 %d 8000 overlap" % (q, 0, -(3000 + 10 * q), 3000, 3100)
                                             for q in range(1, 41)] +
                                            ["%09d %09d %d 99.0 0 0 %d 9000 0 100 %d 8000 overlap" % (q, 41, -4000, 3000, 3100) for q in (3, 4)]}
    rr = type("RR", (), dict(phased_reads=[], read_to_contig_map=r2c, rawread_ids="\n".join(names) + "\n"))()
    p = _write(rr, str(tmp_path))
    want = rr_oracle.run_track_reads(las, [], r2c, rr.rawread_ids, 2500, 40)
    monkeypatch.setattr(rr_hctg_track, "read_las_lines", lambda db_fn, fn: iter(las[fn]))
    rr_hctg_track.run_track_reads(None, p["phased"], p["r2c"], p["ids"], list(las), 2500, 40, "raw_reads.db", p["out"])
    got = open(p["out"]).read()
    assert len([l for l in want.splitlines() if l.startswith("000000000 ")]) == 120
    assert got == want
