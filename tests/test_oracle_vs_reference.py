"""Pins the CPU oracle (oracle/phasing_oracle.c) against the reference's OWN source text
executed under Python 3 (oracle/ref_exec.py).  Only runs where the
reference tree is mounted (the build container); the GPU box uses the committed goldens."""
import os

import numpy as np
import pytest

import cases
from conftest import synth_set
from oracle import c_oracle, ref_exec

pytestmark = pytest.mark.skipif(not ref_exec.available(), reason="reference tree not mounted")

_mod = None


def ref_mod():
    global _mod
    if _mod is None:
        _mod = ref_exec.load_phasing()
    return _mod


def run_both(records, refs, ref, tmp, ctg=cases.CTG):
    from falcon_unzip_b200 import bam
    sam = os.path.join(str(tmp), "in.sam")
    os.makedirs(str(tmp), exist_ok=True)
    with open(sam, "w") as f:
        f.write("\n".join(bam.sam_lines_from_records(records, refs)) + "\n")
    want = ref_exec.run_phasing_stages(sam, ctg, ref, os.path.join(str(tmp), "ref"), mod=ref_mod())
    got = c_oracle.run_phasing_stages(records, ctg, ref, os.path.join(str(tmp), "oracle"))
    return want, got


def assert_same(want, got):
    """Two assertions per file: the order-free content (multiset of rows; independent of the emulated CPython-2 container
    orders of SURVEY.md B.3 / B.4) and then the bytes (row order included)."""
    for k in want:
        a, b = open(want[k]).read(), open(got[k]).read()
        assert sorted(a.splitlines()) == sorted(b.splitlines()), "%s: CONTENT differs:\nreference:\n%s\noracle:\n%s" % (k, a[:600], b[:600])
        assert a == b, "%s: same rows, different ORDER:\nreference:\n%s\noracle:\n%s" % (k, a[:600], b[:600])


@pytest.mark.parametrize("name", sorted(cases.all_cases()))
def test_appendix_e_cases(name, tmp_path):
    recs, ref = cases.all_cases()[name]
    records, refs = cases.build(recs, len(ref))
    want, got = run_both(records, refs, ref, tmp_path)
    assert_same(want, got)


@pytest.mark.parametrize("gap", [65536, 65537])
def test_window_edge(gap, tmp_path):
    recs, ref = cases.window_case(gap)
    records, refs = cases.build(recs, len(ref))
    want, got = run_both(records, refs, ref, tmp_path)
    assert_same(want, got)
    n_rows = len(open(want["atable"]).read().splitlines())
    assert n_rows == (1 if gap == 65536 else 0)


@pytest.mark.parametrize("cfg", ["tiny", "quirks", "noisy", "noisy_m"])
def test_synthetic_sets(cfg, tmp_path):
    sset = synth_set(cfg)
    for c, (name, _l) in enumerate(sset.refs):
        want, got = run_both(sset.contig_records(c), sset.refs, sset.ref_seqs[c], tmp_path / name, ctg=name)
        assert_same(want, got)


def write_stage_inputs(d, rows):
    os.makedirs(d)
    with open(os.path.join(d, "variant_map"), "w") as f:
        f.write("".join("%d %s %s %d\n" % r for r in rows))
    with open(os.path.join(d, "q_id_map"), "w") as f:
        f.write("".join("%d read%d\n" % (q, q) for q in range(max(r[3] for r in rows) + 1)))


def reference_stages_2_to_4(d):
    m = ref_mod()
    T = ref_exec.TaskSelf
    p = lambda k: os.path.join(d, k)
    m.generate_association_table(T(dict(ctg_id="c", base_dir="."), vmap_file=p("variant_map"), atable_file=p("atable")))
    m.get_phased_blocks(T({}, vmap_file=p("variant_map"), atable_file=p("atable"),
                          phased_variant_file=p("phased_variants")))
    m.get_phased_reads(T(dict(ctg_id="c"), vmap_file=p("variant_map"), q_id_map_file=p("q_id_map"),
                         phased_variant_file=p("phased_variants"), phased_read_file=p("phased_reads")))


def oracle_stages_2_to_4(d):
    p = lambda k: os.path.join(d, k)
    c_oracle.generate_association_table_files(p("variant_map"), p("atable"))
    c_oracle.get_phased_blocks_files(p("variant_map"), p("atable"), p("phased_variants"))
    c_oracle.get_phased_reads_files(p("variant_map"), p("q_id_map"), p("phased_variants"), "c", p("phased_reads"))


@pytest.mark.parametrize("seed", range(12))
def test_stage_fuzz_atable_blocks_reads(seed, tmp_path):
    """Random variant_map files (duplicates, noisy reads) through stages 2-4 of both."""
    rng = np.random.default_rng(1000 + seed)
    n_sites = int(rng.integers(5, 120))
    _pos, _refs, rows = cases.random_vmap(rng, n_sites, int(rng.integers(8, 30)), int(rng.integers(30, 200)),
                                          dup_rate=float(rng.choice([0.0, 0.1, 0.3])))
    d_ref, d_or = str(tmp_path / "ref"), str(tmp_path / "oracle")
    write_stage_inputs(d_ref, rows)
    write_stage_inputs(d_or, rows)
    reference_stages_2_to_4(d_ref)
    oracle_stages_2_to_4(d_or)
    for k in ("atable", "phased_variants", "phased_reads"):
        assert open(os.path.join(d_ref, k)).read() == open(os.path.join(d_or, k)).read(), k
