"""tests/golden/select_seed7.json (made from the reference's own select_reads_from_bam.py by scripts/make_golden_select.py)
against the CPU oracle -- runs everywhere -- and, on a GPU, against the device path."""
import hashlib
import json
import os
import sys

import pytest

import select_cases

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "scripts"))
GOLD = json.load(open(os.path.join(HERE, "golden", "select_seed7.json")))


def _inputs(tmp_path):
    import make_golden_select
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=GOLD["seed"])
    if make_golden_select.input_digests(fofn, r2c, ids) != GOLD["inputs"]:
        pytest.skip("the synthetic generator gives other inputs here than where the golden was made")
    return fofn, r2c, ids


def _check(streams, header):
    assert sorted(streams) == sorted(GOLD["contigs"])
    for ctg, want in GOLD["contigs"].items():
        data = streams[ctg]
        assert len(data) == want["bytes"] and hashlib.sha256(data).hexdigest() == want["sha256"], ctg
    assert header == GOLD["header"]


def test_oracle_reproduces_select_golden(tmp_path):
    from oracle import select_oracle
    fofn, r2c, ids = _inputs(tmp_path)
    header, out = select_oracle.select(fofn, r2c, ids)
    _check({c: b"".join(r) for c, r in out.items()}, header)


@pytest.mark.gpu
def test_device_path_reproduces_select_golden(tmp_path):
    from falcon_unzip_b200 import bam, select_reads_from_bam as srb
    from oracle import select_oracle
    fofn, r2c, ids = _inputs(tmp_path)
    sam_dir = str(tmp_path / "reads")
    made = srb.select_reads_from_bam(fofn, r2c, ids, sam_dir, level=1)
    streams, header = {}, None
    for ctg in made:
        text, _refs, recs = bam.read_bam(os.path.join(sam_dir, "%s.bam" % ctg))
        streams[ctg] = bytes(recs)
        header = select_oracle.parse_header(text)
    _check(streams, header)
