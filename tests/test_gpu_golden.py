"""GPU parity against the committed golden vectors and the crafted Appendix-E cases: the
reference-signature functions of falcon_unzip_b200.phasing (through the C ABI) must write
byte-identical files."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

import cases
from test_golden_cpu import CASES, FILES, GOLD, STAGE_CASES, load_case

pytestmark = pytest.mark.gpu


def _paths(base, ctg):
    b = os.path.join(base, ctg)
    return dict(variant_map=os.path.join(b, "het_call", "variant_map"), variant_pos=os.path.join(b, "het_call", "variant_pos"),
                q_id_map=os.path.join(b, "het_call", "q_id_map"), atable=os.path.join(b, "g_atable", "atable"),
                phased_variants=os.path.join(b, "get_phased_blocks", "phased_variants"),
                phased_reads=os.path.join(b, "phased_reads"))


@pytest.mark.parametrize("case", CASES)
def test_cli_on_golden_inputs(eng, case, tmp_path):
    """fc_phasing.py command line on the golden BAM + FASTA (four stages through files)."""
    from falcon_unzip_b200 import phasing
    d, _refs, _records, ctg, _ref = load_case(case)
    phasing.main(["fc_phasing.py", "--bam", os.path.join(d, "in.bam"), "--fasta", os.path.join(d, "ref.fa"),
                  "--ctg_id", ctg, "--base_dir", str(tmp_path)])
    got = _paths(str(tmp_path), ctg)
    for k in FILES:
        assert open(os.path.join(d, k)).read() == open(got[k]).read(), (case, k)


@pytest.mark.parametrize("case", CASES)
def test_fused_batch_on_golden_inputs(eng, case, tmp_path):
    from falcon_unzip_b200 import phasing
    d, _refs, records, ctg, ref = load_case(case)
    _res, got = phasing.phase_contigs(records, [ctg], [ref], str(tmp_path))
    for k in FILES:
        assert open(os.path.join(d, k)).read() == open(got[ctg][k]).read(), (case, k)


@pytest.mark.parametrize("case", STAGE_CASES)
def test_stage_functions_on_golden_inputs(eng, case, tmp_path):
    """generate_association_table / get_phased_blocks / get_phased_reads called one by one
    with the reference's PypeTask `self` convention, each from the golden input files."""
    from falcon_unzip_b200 import phasing
    d = os.path.join(GOLD, case)
    p = lambda k: os.path.join(d, k)
    o = lambda k: str(tmp_path / k)
    phasing.generate_association_table(SimpleNamespace(vmap_file=p("variant_map"), atable_file=o("atable"),
                                                       parameters=dict(ctg_id="c", base_dir=".")))
    phasing.get_phased_blocks(SimpleNamespace(vmap_file=p("variant_map"), atable_file=p("atable"),
                                              phased_variant_file=o("phased_variants"), parameters={}))
    phasing.get_phased_reads(SimpleNamespace(vmap_file=p("variant_map"), q_id_map_file=p("q_id_map"),
                                             phased_variant_file=p("phased_variants"), phased_read_file=o("phased_reads"),
                                             parameters=dict(ctg_id="c")))
    for k in ("atable", "phased_variants", "phased_reads"):
        assert open(p(k)).read() == open(o(k)).read(), (case, k)


@pytest.mark.parametrize("gap", [65536, 65537])
def test_window_edge_vs_oracle(eng, gap, tmp_path):
    from falcon_unzip_b200 import phasing
    from oracle import c_oracle
    recs, ref = cases.window_case(gap)
    records, _refs = cases.build(recs, len(ref))
    want = c_oracle.run_phasing_stages(records, cases.CTG, ref, str(tmp_path / "oracle"))
    _res, got = phasing.phase_contigs(records, [cases.CTG], [ref], str(tmp_path / "gpu"))
    for k in FILES:
        assert open(want[k]).read() == open(got[cases.CTG][k]).read(), k
    assert len(open(want["atable"]).read().splitlines()) == (1 if gap == 65536 else 0)


@pytest.mark.parametrize("seed", range(6, 30))
def test_stage_fuzz_vs_oracle(eng, seed, tmp_path):
    """Random variant_map inputs through the three later stage functions, GPU vs oracle."""
    from falcon_unzip_b200 import phasing
    from oracle import c_oracle
    rng = np.random.default_rng(1000 + seed)
    n_sites = int(rng.integers(5, 400))
    _p, _r, rows = cases.random_vmap(rng, n_sites, int(rng.integers(8, 40)), int(rng.integers(30, 600)),
                                     dup_rate=float(rng.choice([0.0, 0.1, 0.3])))
    d_g, d_o = tmp_path / "gpu", tmp_path / "oracle"
    for d in (d_g, d_o):
        os.makedirs(d)
        with open(d / "variant_map", "w") as f:
            f.write("".join("%d %s %s %d\n" % r for r in rows))
        with open(d / "q_id_map", "w") as f:
            f.write("".join("%d read%d\n" % (q, q) for q in range(max(r[3] for r in rows) + 1)))
    g, o = (lambda k: str(d_g / k)), (lambda k: str(d_o / k))
    c_oracle.generate_association_table_files(o("variant_map"), o("atable"))
    c_oracle.get_phased_blocks_files(o("variant_map"), o("atable"), o("phased_variants"))
    c_oracle.get_phased_reads_files(o("variant_map"), o("q_id_map"), o("phased_variants"), "c", o("phased_reads"))
    phasing.generate_association_table(SimpleNamespace(vmap_file=g("variant_map"), atable_file=g("atable"),
                                                       parameters=dict(ctg_id="c", base_dir=".")))
    phasing.get_phased_blocks(SimpleNamespace(vmap_file=g("variant_map"), atable_file=g("atable"),
                                              phased_variant_file=g("phased_variants"), parameters={}))
    phasing.get_phased_reads(SimpleNamespace(vmap_file=g("variant_map"), q_id_map_file=g("q_id_map"),
                                             phased_variant_file=g("phased_variants"), phased_read_file=g("phased_reads"),
                                             parameters=dict(ctg_id="c")))
    for k in ("atable", "phased_variants", "phased_reads"):
        assert open(o(k)).read() == open(g(k)).read(), k


def test_c1_scale_contig_vs_oracle(eng, tmp_path):
    """A 300 kb / 30x contig (C1 shape at reduced length so the oracle finishes in seconds)."""
    from conftest import synth_set
    from falcon_unzip_b200 import phasing
    from oracle import c_oracle
    sset = synth_set("c1", contig_len=300_000)
    want = c_oracle.run_phasing_stages(sset.contig_records(0), sset.refs[0][0], sset.ref_seqs[0], str(tmp_path / "oracle"))
    _res, got = phasing.phase_contigs(sset.records, [sset.refs[0][0]], sset.ref_seqs, str(tmp_path / "gpu"))
    for k in FILES:
        assert open(want[k]).read() == open(got[sset.refs[0][0]][k]).read(), k
    assert len(open(want["phased_reads"]).read().splitlines()) > 500


@pytest.mark.parametrize("staging", [1, 2])
def test_phase_staging_tiers_vs_oracle(eng, staging, tmp_path):
    """k_ctg_phase keeps a contig in shared memory in full, for the sweep only, or not at all
    (large contigs); the two smaller tiers are forced here on a contig that would fit in full."""
    from conftest import synth_set
    from falcon_unzip_b200 import phasing
    from oracle import c_oracle
    sset = synth_set("c1", contig_len=300_000)
    name = sset.refs[0][0]
    want = c_oracle.run_phasing_stages(sset.contig_records(0), name, sset.ref_seqs[0], str(tmp_path / "oracle"))
    eng.set_option("phase_staging", staging)
    try:
        _res, got = phasing.phase_contigs(sset.records, [name], sset.ref_seqs, str(tmp_path / "gpu"))
    finally:
        eng.set_option("phase_staging", 0)
    for k in FILES:
        assert open(want[k]).read() == open(got[name][k]).read(), k


@pytest.mark.parametrize("passes,staging", [(0, 0), (1, 0), (2, 2), (64, 2)])
def test_sweep_fixed_point_and_sequential_walk_agree(eng, passes, staging, tmp_path):
    """The pass-2 sweep (phasing.py:311-344) runs as parallel fixed-point passes with the sequential walk as the fallback
    (option sweep_passes): no pass at all, one or two passes (the fallback takes over after them unless they already
    converged) and the default give the oracle's files, from shared and from global memory."""
    from conftest import synth_set
    from falcon_unzip_b200 import phasing
    from oracle import c_oracle
    sset = synth_set("c1", contig_len=300_000)
    name = sset.refs[0][0]
    want = c_oracle.run_phasing_stages(sset.contig_records(0), name, sset.ref_seqs[0], str(tmp_path / "oracle"))
    eng.set_option("sweep_passes", passes)
    eng.set_option("phase_staging", staging)
    try:
        _res, got = phasing.phase_contigs(sset.records, [name], sset.ref_seqs, str(tmp_path / "gpu"))
    finally:
        eng.set_option("sweep_passes", 64)
        eng.set_option("phase_staging", 0)
    for k in FILES:
        assert open(want[k]).read() == open(got[name][k]).read(), k


def test_c1_full_size_contig_vs_oracle(eng, tmp_path):
    """BASELINE.json config 1 at full size: 1 Mb contig, 0.1 % het, 30x 10 kb reads, 1 % error."""
    from falcon_unzip_b200 import phasing, synth
    from oracle import c_oracle
    sset = synth.generate(synth.CONFIGS["c1"])
    name = sset.refs[0][0]
    want = c_oracle.run_phasing_stages(sset.contig_records(0), name, sset.ref_seqs[0], str(tmp_path / "oracle"))
    res, got = phasing.phase_contigs(sset.records, [name], sset.ref_seqs, str(tmp_path / "gpu"))
    for k in FILES:
        assert open(want[k]).read() == open(got[name][k]).read(), k
    assert res.n_sites > 800 and res.n_reads > 2500


def test_c2_shape_batch_vs_oracle_and_batch_independence(eng, tmp_path):
    """Six contigs of the C2 shape (250 kb, 40x) in ONE fused call: files equal the oracle's, and
    equal what the same contigs give when phased alone (contigs are independent units)."""
    import dataclasses
    import numpy as np
    from falcon_unzip_b200 import phasing, synth
    from oracle import c_oracle
    sset = synth.generate_parallel(dataclasses.replace(synth.CONFIGS["c2"], n_contigs=6))
    names = [r[0] for r in sset.refs]
    _res, got = phasing.phase_contigs(sset.records, names, sset.ref_seqs, str(tmp_path / "gpu"), host_path=False)
    for c, name in enumerate(names):
        want = c_oracle.run_phasing_stages(sset.contig_records(c), name, sset.ref_seqs[c], str(tmp_path / "oracle"))
        for k in FILES:
            assert open(want[k]).read() == open(got[name][k]).read(), (name, k)
    idx = np.flatnonzero(sset.rec_ctg == 3)
    sub = sset.records[sset.rec_off[idx[0]]:sset.rec_off[idx[-1] + 1]]
    _r, alone = phasing.phase_contigs(sub, [names[3]], [sset.ref_seqs[3]], str(tmp_path / "alone"),
                                      ctg_rec_off=np.asarray([0, len(idx)], np.int32))
    for k in FILES:
        assert open(alone[names[3]][k]).read() == open(got[names[3]][k]).read(), k


def test_many_small_contigs_c3_shape(eng, tmp_path):
    """C3 shape: many short contigs (67.5 kb, 50x) batched into one launch; 40 of them here."""
    import dataclasses
    from falcon_unzip_b200 import phasing, synth
    from oracle import c_oracle
    sset = synth.generate_parallel(dataclasses.replace(synth.CONFIGS["c3"], n_contigs=40))
    names = [r[0] for r in sset.refs]
    _res, got = phasing.phase_contigs(sset.records, names, sset.ref_seqs, str(tmp_path / "gpu"))
    for c in (0, 17, 39):
        want = c_oracle.run_phasing_stages(sset.contig_records(c), names[c], sset.ref_seqs[c], str(tmp_path / "oracle"))
        for k in FILES:
            assert open(want[k]).read() == open(got[names[c]][k]).read(), (names[c], k)


def test_c5_shape_contigs_vs_oracle_and_batch_splitting(eng, tmp_path):
    """BASELINE.json config 5 shape (60x, 15 kb reads, 1 % error, 0.1 % het) from the generator bench.py uses
    (libfuz_synth.so): three contigs in one batch equal the oracle's files, and the same contigs forced into one batch
    per contig (engine.plan_batches with a small byte limit) give the same files."""
    import dataclasses
    import numpy as np
    from falcon_unzip_b200 import engine, phasing, synth
    from oracle import c_oracle
    cfg = dataclasses.replace(synth.CONFIGS["c5"], contig_len=400_000)
    parts = [synth.generate_contig_fast(cfg, ci) for ci in (0, 1, 2)]
    names = [p.refs[0][0] for p in parts]
    seqs = [p.ref_seqs[0] for p in parts]
    pb = engine.build_batch([(p.records, p.rec_off) for p in parts], names, [cfg.contig_len] * 3)
    cro = np.asarray(pb.ctg_rec_off, np.int32)
    _res, got = phasing.phase_contigs(pb.records, names, seqs, str(tmp_path / "gpu"), ctg_rec_off=cro)
    res_split, got_split = phasing.phase_contigs(pb.records, names, seqs, str(tmp_path / "split"), ctg_rec_off=cro,
                                                 max_batch_bytes=len(parts[0].records) + 1000)
    assert isinstance(res_split, list) and len(res_split) == 3
    for c, name in enumerate(names):
        want = c_oracle.run_phasing_stages(parts[c].records.tobytes(), name, seqs[c], str(tmp_path / "oracle"))
        for k in FILES:
            ref = open(want[k]).read()
            assert ref == open(got[name][k]).read(), (name, k)
            assert ref == open(got_split[name][k]).read(), (name, k, "split")
    assert len(open(got[names[0]]["phased_reads"]).read().splitlines()) > 1000
