"""The reference's own entry scripts import `falcon_unzip.<module>` (src/py_scripts/fc_phasing.py:2 etc.) and pypeFLOW
looks task functions up by module path (unzip.py:304): the alias package must expose the same names."""
import importlib

import pytest


@pytest.mark.parametrize("mod,names", [
    ("phasing", ["make_het_call", "generate_association_table", "get_score", "get_phased_blocks", "get_phased_reads",
                 "phasing", "parse_args", "main"]),
    ("rr_hctg_track", ["get_rid_to_ctg", "run_tr_stage1", "tr_stage1", "run_track_reads", "try_run_track_reads",
                       "track_reads", "parse_args", "main"]),
    ("ovlp_filter_with_phase", ["filter_stage1", "filter_stage2", "filter_stage3", "parse_args", "main"]),
    ("select_reads_from_bam", ["select_reads_from_bam", "parse_args", "main"]),
    ("phasing_readmap", ["get_phasing_readmap", "parse_args", "main"]),
    ("get_read_hctg_map", ["generate_read_to_hctg_map", "get_read_hctg_map", "parse_args", "main"]),
])
def test_alias_module_exposes_reference_names(mod, names):
    m = importlib.import_module("falcon_unzip." + mod)
    missing = [n for n in names if not hasattr(m, n)]
    assert not missing, "falcon_unzip.%s lacks %s" % (mod, missing)


def test_reference_shim_line_imports_unchanged():
    ns = {}
    exec("from falcon_unzip.phasing import main", ns)          # the line of reference src/py_scripts/fc_phasing.py:2
    from falcon_unzip_b200 import phasing
    assert ns["main"] is phasing.main
