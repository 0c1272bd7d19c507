"""The CPU oracle against the committed golden vectors (tests/golden/, produced from the
reference's own source by scripts/make_golden.py) -- runs everywhere, no GPU, no reference."""
import glob
import os

import pytest

from oracle import c_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FILES = ("variant_pos", "variant_map", "q_id_map", "atable", "phased_variants", "phased_reads")
CASES = sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(GOLD, "*", "in.bam")))
STAGE_CASES = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "stage_fuzz_*")))


def load_case(case):
    from falcon_unzip_b200 import bam
    d = os.path.join(GOLD, case)
    _text, refs, records = bam.read_bam(os.path.join(d, "in.bam"))
    name, seq = next(iter(bam.read_fasta(os.path.join(d, "ref.fa"))))
    return d, refs, records, name, seq


def test_golden_inventory():
    assert len(CASES) >= 15 and len(STAGE_CASES) >= 6


@pytest.mark.parametrize("case", CASES)
def test_oracle_reproduces_golden(case, tmp_path):
    d, _refs, records, ctg, ref = load_case(case)
    got = c_oracle.run_phasing_stages(records, ctg, ref, str(tmp_path))
    for k in FILES:
        assert open(os.path.join(d, k)).read() == open(got[k]).read(), (case, k)


@pytest.mark.parametrize("case", STAGE_CASES)
def test_oracle_stage_level_golden(case, tmp_path):
    d = os.path.join(GOLD, case)
    p = lambda k: os.path.join(d, k)
    o = lambda k: str(tmp_path / k)
    c_oracle.generate_association_table_files(p("variant_map"), o("atable"))
    c_oracle.get_phased_blocks_files(p("variant_map"), p("atable"), o("phased_variants"))
    c_oracle.get_phased_reads_files(p("variant_map"), p("q_id_map"), p("phased_variants"), "c", o("phased_reads"))
    for k in ("atable", "phased_variants", "phased_reads"):
        assert open(p(k)).read() == open(o(k)).read(), (case, k)
