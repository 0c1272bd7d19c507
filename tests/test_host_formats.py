"""CPU tests of the host layer: the device data model (include/fuz.h) filled from ORACLE
arrays must render to the oracle's files byte for byte.  Covers formats.py, the C++ host
helpers (record index, q_id assignment, py2 dict order) and prepare_batch -- no kernels."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import synth_set

LETTER_IDX = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
ACTG_RANK = {0: 0, 1: 1, 3: 2, 2: 3}


def model_from_oracle(recs, ctg_id, ref_seq):
    """Oracle arrays of one contig -> the struct-of-arrays the kernels produce."""
    from oracle import c_oracle
    off = c_oracle.index_records(recs)
    qid, names = c_oracle.assign_qids(c_oracle.record_names(recs, off))
    h = c_oracle.het_call(recs, off, qid)
    n_sites = len(h["site_pos"])
    site_pos = h["site_pos"] + 1
    cnt = np.zeros((n_sites, 4), np.int32)
    for k in range(4):
        for i in range(n_sites):
            cnt[i, LETTER_IDX[h["site_base"][i, k]]] = h["site_count"][i, k]
    top = np.vectorize(LETTER_IDX.get)(h["site_base"][:, :2]).astype(np.uint8) if n_sites else np.zeros((0, 2), np.uint8)
    al = np.array([sorted(t, key=ACTG_RANK.get) for t in top.tolist()], np.uint8).reshape(-1, 2)
    vm_site = np.searchsorted(site_pos, h["vm_pos"] + 1).astype(np.int32)
    vm_base = np.vectorize(LETTER_IDX.get)(h["vm_allele"]).astype(np.uint8) if len(vm_site) else np.zeros(0, np.uint8)
    t = c_oracle.association_table(h["vm_pos"] + 1, h["vm_allele"], h["vm_qid"])
    at_s1 = np.searchsorted(site_pos, t["pos1"]).astype(np.int32)
    at_s2 = np.searchsorted(site_pos, t["pos2"]).astype(np.int32)
    b = c_oracle.phased_blocks(t["pos1"], t["pos2"], t["b"], t["ct"])
    ph_block = np.zeros(n_sites, np.int32); ph_state = np.full(n_sites, 255, np.uint8)
    ph = {k: np.zeros(n_sites, np.int32) for k in ("lext", "rext", "lscore", "rscore")}
    for i in range(len(b["pid"])):
        s = int(np.searchsorted(site_pos, b["pos"][i]))
        ph_block[s] = b["pid"][i]
        ph_state[s] = 0 if LETTER_IDX[b["h"][i, 0]] == al[s, 0] else 1
        for k in ph:
            ph[k][s] = b[k][i]
    r = c_oracle.phased_reads(h["vm_pos"] + 1, h["vm_allele"], h["vm_qid"], b["pid"], b["pos"], b["h"])
    order = np.lexsort((r["pid"], r["qid"]))
    res = SimpleNamespace(
        site_ctg=np.zeros(n_sites, np.int32), site_pos=site_pos.astype(np.int32), site_cnt=cnt, site_al=al,
        site_top=top, vm_site=vm_site, vm_base=vm_base, vm_qid=h["vm_qid"], at_s1=at_s1, at_s2=at_s2,
        at_ct=t["ct"], ph_state=ph_state, ph_block=ph_block, ph_lext=ph["lext"], ph_rext=ph["rext"],
        ph_lscore=ph["lscore"], ph_rscore=ph["rscore"], pr_ctg=np.zeros(len(order), np.int32),
        pr_qid=r["qid"][order], pr_block=r["pid"][order], pr_phase=r["phase"][order], pr_n0=r["n0"][order],
        pr_n1=r["n1"][order])
    return res, names


@pytest.mark.parametrize("name_rows", [False, True])
@pytest.mark.parametrize("cfg", ["tiny", "quirks"])
def test_formats_render_oracle_model_to_oracle_bytes(cfg, name_rows, tmp_path):
    """name_rows: the QNAMEs as the fixed-width byte rows a device batch keeps (formatted in libfuz without str objects)."""
    from falcon_unzip_b200 import formats, phasing
    from oracle import c_oracle
    sset = synth_set(cfg)
    for c, (name, _l) in enumerate(sset.refs):
        recs = sset.contig_records(c)
        want = c_oracle.run_phasing_stages(recs, name, sset.ref_seqs[c], str(tmp_path / "oracle"))
        res, names = model_from_oracle(recs, name, sset.ref_seqs[c])
        if name_rows:
            names = np.array([n.encode("latin-1") for n in names], dtype="S")
        sl = formats.contig_slices(res, 1)
        got = phasing.write_contig_files(res, sl, 0, name, sset.ref_seqs[c], names, str(tmp_path / "fmt"))
        for k in want:
            assert open(want[k]).read() == open(got[k]).read(), (name, k)
        # the Python rendering of phased_variants (kept for the file-level stage, which looks bases up in a dict) agrees
        n_sites = len(res.site_pos)
        assert formats.phased_variants_bytes(res, 0, n_sites, sset.ref_seqs[c]).decode() == \
            formats.phased_variants_text(res, 0, n_sites, sset.ref_seqs[c])


def test_native_formatters_edge_cases():
    """Python 2 float text of the P rows (integral quotient -> '.0', 12 significant digits), ties in variant_pos (T > G > C > A
    on equal counts), empty inputs, a position beyond the reference (IndexError like the reference, phasing.py:123)."""
    from falcon_unzip_b200 import formats
    z = np.zeros
    res = SimpleNamespace(site_pos=np.array([3, 10, 17, 20, 1000], np.int32), site_al=np.array([[0, 1], [1, 3], [0, 2], [3, 2], [0, 1]], np.uint8),
                          site_cnt=np.array([[5, 5, 5, 5], [0, 7, 0, 7], [9, 1, 2, 3], [1, 1, 8, 8], [4, 3, 2, 1]], np.int32),
                          ph_block=np.array([1, 1, 1, 2, 0], np.int32), ph_state=np.array([0, 1, 0, 1, 255], np.uint8),
                          ph_lext=np.array([3, 3, 10, 20, 0], np.int32), ph_rext=np.array([17, 17, 17, 20, 0], np.int32),
                          ph_lscore=np.array([0, 12, 30, 0, 0], np.int32), ph_rscore=np.array([14, 11, 0, 0, 0], np.int32))
    ref = "ACGTACGTACGTACGTACGTACGT"
    assert formats.variant_pos_bytes(res, 0, 4, ref) == (b"3 G 20 T 5 G 5 C 5 A 5\n10 C 14 T 7 C 7 G 0 A 0\n"
                                                         b"17 A 15 A 9 T 3 G 2 C 1\n20 T 18 T 8 G 8 C 1 A 1\n")
    want = ("P 1 3 17 14 3 4.66666666667\nV 1 3 3_G_A 3_G_C 3 17 0 14\nV 1 10 10_C_T 10_C_C 3 17 12 11\nV 1 17 17_A_A 17_A_G 10 17 30 0\n"
            "P 2 20 20 0 1 0.0\nV 2 20 20_T_G 20_T_T 20 20 0 0\n")
    assert formats.phased_variants_bytes(res, 0, 5, ref).decode() == want == formats.phased_variants_text(res, 0, 5, ref)
    assert formats.phased_variants_bytes(res, 0, 0, ref) == b"" and formats.variant_pos_bytes(res, 2, 2, ref) == b""
    with pytest.raises(IndexError):
        formats.variant_pos_bytes(res, 0, 5, ref)
    rows = np.array([b"m1/1/0_9", b"", b"x" * 12], dtype="S12")
    assert formats.q_id_map_bytes(rows) == b"0 m1/1/0_9\n1 \n2 xxxxxxxxxxxx\n" == formats.q_id_map_bytes(["m1/1/0_9", "", "x" * 12])
    assert formats.q_id_map_bytes(np.zeros(0, dtype="S8")) == b""


def test_prepare_batch_matches_oracle_qids():
    from falcon_unzip_b200 import engine
    from oracle import c_oracle
    sset = synth_set("quirks")
    pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
    assert np.array_equal(pb.rec_off, sset.rec_off)
    assert np.array_equal(np.diff(pb.ctg_rec_off), np.bincount(sset.rec_ctg, minlength=len(sset.refs)))
    for c in range(pb.n_ctg):
        recs = sset.contig_records(c)
        off = c_oracle.index_records(recs)
        qid, names = c_oracle.assign_qids(c_oracle.record_names(recs, off))
        assert np.array_equal(pb.rec_qid[pb.ctg_rec_off[c]:pb.ctg_rec_off[c + 1]], qid)
        assert pb.qnames(c) == names
    t = pb.goff()
    assert np.all(t % 8192 == 0) and np.all(np.diff(t) >= pb.ctg_len)


def test_py27_int_dict_order_vectors_and_emulator():
    """SURVEY.md Appendix E18 hand-derived vectors + agreement with the oracle's emulator."""
    from falcon_unzip_b200 import formats
    from oracle import py2emu
    assert formats.py27_int_dict_order([1, 9, 17, 2]).tolist() == [1, 2, 17, 9]
    assert formats.py27_int_dict_order([8, 0, 16, 1]).tolist() == [8, 0, 16, 1]
    assert formats.py27_int_dict_order([5, 13, 21, 3, 11]).tolist() == [11, 3, 21, 5, 13]
    rng = np.random.default_rng(5)
    for n in (0, 1, 5, 6, 21, 22, 85, 86, 1000, 60000):
        keys = rng.permutation(rng.choice(4 * n + 10, size=n, replace=False))
        assert formats.py27_int_dict_order(keys).tolist() == py2emu.py27_int_dict_order(keys.tolist())
    assert formats.py27_float_str(183848 / 183) == "1004.63387978"
    assert formats.py27_float_str(1000 / 4) == "250.0"


def test_read_fasta_bytes_agrees_with_read_fasta(tmp_path):
    """The bytes reader of phase_bam keeps the record and line rules of the FastaReader stand-in (reference phasing.py:490-494)."""
    from falcon_unzip_b200 import bam
    fn = str(tmp_path / "a.fa")
    with open(fn, "w", newline="") as f:
        f.write("junk before the first header\n>c1 some description\nACGT\nacgt \n\n>c2\nTTTT\n>c3\n>c4\r\nAA\r\nCC\r\n>c5\nGATTACA")
    a, b = list(bam.read_fasta(fn)), list(bam.read_fasta_bytes(fn))
    assert [(h, s.encode()) for h, s in a] == b
    assert [h for h, _ in b] == ["c1 some description", "c2", "c3", "c4", "c5"] and b[0][1] == b"ACGTacgt" and b[4][1] == b"GATTACA"
