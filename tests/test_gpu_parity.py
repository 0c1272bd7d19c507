"""GPU parity: the CUDA path (through the C ABI of libfuz.so) against the CPU oracle on the
same seeded inputs -- bit-exact arrays and byte-exact files."""
import os

import numpy as np
import pytest

from conftest import synth_set

pytestmark = pytest.mark.gpu


def _oracle_files(sset, tmp):
    from oracle import c_oracle
    out = {}
    for c, (name, _l) in enumerate(sset.refs):
        out[name] = c_oracle.run_phasing_stages(sset.contig_records(c), name, sset.ref_seqs[c], str(tmp))
    return out


def _assert_same_files(a, b):
    for ctg in a:
        for k in a[ctg]:
            ta, tb = open(a[ctg][k]).read(), open(b[ctg][k]).read()
            # order-free content first (does not depend on the CPython-2 container-order emulation), then the bytes
            assert sorted(ta.splitlines()) == sorted(tb.splitlines()), "%s/%s: row CONTENT differs (%d vs %d rows)" % (
                ctg, k, len(ta.splitlines()), len(tb.splitlines()))
            if ta != tb:
                la, lb = ta.splitlines(), tb.splitlines()
                for i, (x, y) in enumerate(zip(la, lb)):
                    if x != y:
                        raise AssertionError("%s/%s differs at line %d: oracle %r | gpu %r (%d vs %d lines)"
                                             % (ctg, k, i, x, y, len(la), len(lb)))
                raise AssertionError("%s/%s: %d vs %d lines" % (ctg, k, len(la), len(lb)))


@pytest.mark.parametrize("impl", [0, 1, 2, 3])
@pytest.mark.parametrize("cfg", ["tiny", "quirks", "noisy", "noisy_m"])
def test_pileup_counts_match_oracle(eng, cfg, impl):
    from falcon_unzip_b200 import engine
    from oracle import c_oracle
    sset = synth_set(cfg)
    pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
    eng.set_option("pileup_impl", impl)
    try:
        res = eng.phase_device(pb, want_counts=True, stage="het")
    finally:
        eng.set_option("pileup_impl", 0)
    goff = res.arrays["goff"]
    for c, (name, L) in enumerate(sset.refs):
        recs = sset.contig_records(c)
        want = c_oracle.pileup_counts(recs, c_oracle.index_records(recs), L)
        got = res.arrays["counts"][goff[c]:goff[c] + L]
        bad = np.flatnonzero((want != got).any(axis=1))
        assert len(bad) == 0, "contig %s: %d positions differ, first %d: want %s got %s" % (
            name, len(bad), bad[0], want[bad[0]], got[bad[0]])


@pytest.mark.parametrize("impl", [0, 1, 2, 3])
@pytest.mark.parametrize("cfg", ["tiny", "quirks", "noisy", "noisy_m"])
def test_het_call_arrays_match_oracle(eng, cfg, impl):
    from falcon_unzip_b200 import engine
    from oracle import c_oracle
    sset = synth_set(cfg)
    pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
    eng.set_option("pileup_impl", impl)
    try:
        res = eng.phase_device(pb, stage="het")
    finally:
        eng.set_option("pileup_impl", 0)
    site_off = np.searchsorted(res.site_ctg, np.arange(pb.n_ctg + 1))
    vm_off = np.searchsorted(res.vm_site, site_off)
    aligned = accepted = 0
    for c in range(pb.n_ctg):
        recs = sset.contig_records(c)
        off = c_oracle.index_records(recs)
        qid, _names = c_oracle.assign_qids(c_oracle.record_names(recs, off))
        h = c_oracle.het_call(recs, off, qid)
        aligned += h["aligned_bases"]; accepted += h["n_accepted"]
        s0, s1 = site_off[c], site_off[c + 1]
        assert np.array_equal(res.site_pos[s0:s1], h["site_pos"] + 1)
        # counts in A,C,G,T order vs the oracle's sorted order
        letters = np.frombuffer(b"ACGT", dtype=np.uint8)
        for k in range(4):
            idx = np.argmax(h["site_base"] == letters[k], axis=1)
            assert np.array_equal(res.site_cnt[s0:s1, k], np.take_along_axis(h["site_count"], idx[:, None], 1)[:, 0])
        v0, v1 = vm_off[c], vm_off[c + 1]
        assert np.array_equal(res.site_pos[res.vm_site[v0:v1]], h["vm_pos"] + 1)
        assert np.array_equal(letters[res.vm_base[v0:v1]], h["vm_allele"])
        assert np.array_equal(res.vm_qid[v0:v1], h["vm_qid"])
    assert res.aligned_bases == aligned and res.n_accepted == accepted


@pytest.mark.parametrize("host_path", [True, False])
@pytest.mark.parametrize("cfg", ["tiny", "quirks", "noisy", "noisy_m", "long"])
def test_fused_batch_files_match_oracle(eng, cfg, host_path, tmp_path):
    from falcon_unzip_b200 import phasing
    sset = synth_set(cfg)
    want = _oracle_files(sset, tmp_path / "oracle")
    _res, got = phasing.phase_contigs(sset.records, [r[0] for r in sset.refs], sset.ref_seqs,
                                      str(tmp_path / "gpu"), host_path=host_path)
    _assert_same_files(want, got)


@pytest.mark.parametrize("cfg", ["quirks", "noisy", "long", "c5_slice"])
def test_segment_tma_pileup_files_match_oracle(eng, cfg, tmp_path):
    """pileup_impl 2 (match segments + cp.async.bulk staged SEQ slices, csrc/fuz_pileup_seg.cuh): the six files of every
    contig against the oracle.  noisy: every read takes the global-memory route (more segments per tile than a stage slot);
    long: 100 kb reads over many tiles; c5_slice: the bench shape (15 kb reads, 60x), several tiles per CTA."""
    import dataclasses
    from falcon_unzip_b200 import phasing, synth
    if cfg == "c5_slice":
        sset = synth.generate(dataclasses.replace(synth.CONFIGS["c5"], n_contigs=2, contig_len=150_000))
    else:
        sset = synth_set(cfg)
    eng.set_option("pileup_impl", 2)
    try:
        _res, got = phasing.phase_contigs(sset.records, [r[0] for r in sset.refs], sset.ref_seqs, str(tmp_path / "gpu"))
    finally:
        eng.set_option("pileup_impl", 0)
    _assert_same_files(_oracle_files(sset, tmp_path / "oracle"), got)


@pytest.mark.parametrize("opts", [dict(gather_tma=0), dict(grid_sig=4, grid_assoc=4, grid_reads=4), dict(grid_sig=1, grid_assoc=1, grid_reads=1),
                                  dict(project_ctas=37), dict(sweep_passes=0)])
@pytest.mark.parametrize("cfg", ["quirks", "c5_slice"])
def test_launch_options_do_not_change_the_files(eng, cfg, opts, tmp_path):
    """The tuning options of the default path (the plain tile-per-CTA gather instead of the TMA-fed one, the CTAs per SM of
    the grid-stride kernels, the grid of k_project, the sequential sweep) change launch shapes only: same six files."""
    import dataclasses
    from falcon_unzip_b200 import phasing, synth
    if cfg == "c5_slice":
        sset = synth.generate(dataclasses.replace(synth.CONFIGS["c5"], n_contigs=2, contig_len=150_000))
    else:
        sset = synth_set(cfg)
    defaults = dict(gather_tma=1, grid_sig=8, grid_assoc=6, grid_reads=8, project_ctas=148 * 5, sweep_passes=64)
    for k, v in opts.items():
        eng.set_option(k, v)
    try:
        _res, got = phasing.phase_contigs(sset.records, [r[0] for r in sset.refs], sset.ref_seqs, str(tmp_path / "gpu"))
    finally:
        for k in opts:
            eng.set_option(k, defaults[k])
    _assert_same_files(_oracle_files(sset, tmp_path / "oracle"), got)


@pytest.mark.parametrize("opts", [dict(), dict(gather_tma=0), dict(pileup_impl=2)])
def test_deep_coverage_spills_the_bit_plane_counters(eng, opts, tmp_path):
    """600x over a 24 kb contig: more than 255 reads cover every pileup tile, so the 8-bit vertical counters of the register
    pileup are spilled into the 16-bit counters twice per tile (global scratch in the TMA-fed gather, shared memory in the
    plain one).  Counts and files against the oracle."""
    import dataclasses
    from falcon_unzip_b200 import engine, phasing, synth
    from oracle import c_oracle
    cfg = dataclasses.replace(synth.CONFIGS["tiny"], name="deep", n_contigs=1, contig_len=24_000, coverage=600.0, mean_read_len=4_000,
                              min_read_len=2_500, seed=23)
    sset = synth.generate(cfg)
    defaults = dict(gather_tma=1, pileup_impl=0)
    for k, v in opts.items():
        eng.set_option(k, v)
    try:
        pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
        res = eng.phase_device(pb, want_counts=True, stage="het")
        _res, got = phasing.phase_contigs(sset.records, [r[0] for r in sset.refs], sset.ref_seqs, str(tmp_path / "gpu"))
    finally:
        for k in opts:
            eng.set_option(k, defaults[k])
    recs = sset.contig_records(0)
    want = c_oracle.pileup_counts(recs, c_oracle.index_records(recs), sset.refs[0][1])
    assert want.sum(axis=1).max() > 300
    goff = res.arrays["goff"]
    assert np.array_equal(res.arrays["counts"][goff[0]:goff[0] + sset.refs[0][1]], want)
    _assert_same_files(_oracle_files(sset, tmp_path / "oracle"), got)


def test_reference_cli_per_stage_files_match_oracle(eng, tmp_path):
    """fc_phasing-style run: BAM + FASTA on disk, the four stage functions chained through
    files exactly like reference phasing.py:482-553."""
    from falcon_unzip_b200 import bam, phasing, synth
    sset = synth_set("quirks")
    bam_fn, fa_fn = str(tmp_path / "in.bam"), str(tmp_path / "ref.fa")
    bam.write_bam(bam_fn, sset.refs, sset.records.tobytes())
    synth.write_fasta(fa_fn, sset)
    want = _oracle_files(sset, tmp_path / "oracle")
    got = {}
    for name, _l in sset.refs:
        phasing.main(["fc_phasing.py", "--bam", bam_fn, "--fasta", fa_fn, "--ctg_id", name,
                      "--base_dir", str(tmp_path / "gpu"), "--samtools", "/nonexistent/samtools"])
        base = tmp_path / "gpu" / name
        got[name] = dict(variant_map=str(base / "het_call" / "variant_map"),
                         variant_pos=str(base / "het_call" / "variant_pos"),
                         q_id_map=str(base / "het_call" / "q_id_map"),
                         atable=str(base / "g_atable" / "atable"),
                         phased_variants=str(base / "get_phased_blocks" / "phased_variants"),
                         phased_reads=str(base / "phased_reads"))
    _assert_same_files(want, got)


def test_m_style_cigar_and_single_contig(eng, tmp_path):
    from falcon_unzip_b200 import phasing
    sset = synth_set("tiny", cigar_style="M", n_contigs=1, seed=99)
    want = _oracle_files(sset, tmp_path / "oracle")
    _res, got = phasing.phase_contigs(sset.records, [r[0] for r in sset.refs], sset.ref_seqs, str(tmp_path / "gpu"))
    _assert_same_files(want, got)


@pytest.mark.parametrize("cfg", ["tiny", "quirks"])
def test_selective_fetch_of_pinned_records_equals_full_copy(eng, cfg):
    """Host entry with page-locked records: only header/name/CIGAR/SEQ cross PCIe
    (k_fetch_records); results equal the whole-buffer copy and the pageable-memory path."""
    from falcon_unzip_b200 import engine

    def run(s, fetch, pin):
        eng.set_option("host_fetch", fetch)
        pb = engine.prepare_batch(s.records, [r[0] for r in s.refs], [r[1] for r in s.refs], pin=pin)
        return eng.phase_host(pb)

    other, sset = synth_set("quirks" if cfg == "tiny" else "tiny"), synth_set(cfg)
    try:
        run(other, 1, True)                 # leaves another batch's bytes in the device staging
        got = run(sset, 1, True)
        full = run(sset, 0, True)
        run(other, 1, True)
        pageable = run(sset, 1, False)
    finally:
        eng.set_option("host_fetch", 1)
    assert got.h2d_bytes < 0.6 * full.h2d_bytes and pageable.h2d_bytes == full.h2d_bytes
    for res in (got, pageable):
        assert (res.n_sites, res.n_vmap, res.n_atable, res.n_reads, res.aligned_bases) == \
               (full.n_sites, full.n_vmap, full.n_atable, full.n_reads, full.aligned_bases)
        for k, a in full.arrays.items():
            assert np.array_equal(res.arrays[k], a), k
    assert full.n_sites > 0 and full.n_reads > 0


@pytest.mark.parametrize("cfg", ["tiny", "quirks"])
def test_device_qid_assignment_equals_host_helper(eng, cfg):
    """QNAME -> q_id in first-seen order per contig (phasing.py:47-54): hash-table kernels against
    the host helper (itself pinned to the oracle in test_host_formats), duplicated names included."""
    from falcon_unzip_b200 import engine
    sset = synth_set(cfg)
    pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
    qid, nq, first = eng.assign_qids(eng.upload(pb))
    assert np.array_equal(qid, pb.rec_qid) and np.array_equal(nq, pb.ctg_nq) and np.array_equal(first, pb.name_first)
    if cfg == "quirks":
        assert int(nq.sum()) < pb.n_rec                      # the set carries duplicated QNAMEs


def test_same_qname_in_two_contigs_and_library_assigned_qids(eng):
    """A name seen in two contigs gets an id in each; fuz_phase_batch with d_rec_qid = NULL gives the
    same rows as with q_ids assigned on the host."""
    from falcon_unzip_b200 import bam, engine
    sset = synth_set("tiny")
    off = engine.index_records(sset.records)
    rec = sset.records.copy()
    pb0 = engine.prepare_batch(rec, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
    # rebuild one record of contig 1 with the name of the first record of contig 0
    a, r1 = int(off[0]), int(pb0.ctg_rec_off[1]) + 3
    name = rec[a + 36:a + 36 + int(rec[a + 12])].tobytes()
    o0, o1 = int(off[r1]), int(off[r1 + 1])
    body = bytearray(rec[o0:o1].tobytes())
    body[36:36 + body[12]] = name
    body[12] = len(name)
    body[0:4] = (len(body) - 4).to_bytes(4, "little")
    rec = np.frombuffer(rec[:o0].tobytes() + bytes(body) + rec[o1:].tobytes(), dtype=np.uint8).copy()
    want = engine.prepare_batch(rec, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
    qid, nq, first = eng.assign_qids(eng.upload(want))
    assert np.array_equal(qid, want.rec_qid) and np.array_equal(nq, want.ctg_nq) and np.array_equal(first, want.name_first)
    host_q = eng.phase_device(want)
    pb = engine.prepare_batch(rec, [r[0] for r in sset.refs], [r[1] for r in sset.refs], assign_qids=False)
    dev_q = eng.phase_device(pb)
    assert np.array_equal(pb.ctg_nq, want.ctg_nq) and np.array_equal(pb.name_first, want.name_first)
    for k, v in host_q.arrays.items():
        assert np.array_equal(dev_q.arrays[k], v), k
    assert host_q.n_reads > 0
