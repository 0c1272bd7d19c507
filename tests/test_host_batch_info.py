"""Host-side pieces of the BAM batch path that need no GPU: the QNAME table of a device batch (fixed-width byte rows,
str made per contig) and the best-effort CPU binding of a rank."""
import os

import numpy as np


def test_bam_batch_info_names_from_byte_rows_and_lists():
    from falcon_unzip_b200.engine import BamBatchInfo
    names = ["m1/10/0_100", "m1/11/5_9000", "x", "m2/7/0_12345678"]
    width = max(len(n) for n in names) + 3
    rows = np.zeros((len(names), width), np.uint8)
    for i, n in enumerate(names):
        rows[i, :len(n)] = np.frombuffer(n.encode("ascii"), np.uint8)          # NUL and everything behind it already zero
    s = rows.view("S%d" % width).ravel()
    info = BamBatchInfo(["c0", "c1", "c2"], [100, 200, 300], [1, 0, 3], s)
    assert info.n_ctg == 3
    assert info.qnames(0) == names[:1] and info.qnames(1) == [] and info.qnames(2) == names[1:]
    assert all(isinstance(n, str) for n in info.qnames(2))
    as_list = BamBatchInfo(["c0", "c1", "c2"], [100, 200, 300], [1, 0, 3], list(names))
    assert [as_list.qnames(c) for c in range(3)] == [info.qnames(c) for c in range(3)]
    # bytes above 127 survive (the writers encode latin-1)
    rows[2, 0] = 0xE9
    info2 = BamBatchInfo(["c"], [1], [4], rows.view("S%d" % width).ravel())
    assert info2.qnames(0)[2] == "\xe9"


def test_bind_to_gpu_cpus_is_best_effort():
    from falcon_unzip_b200 import shard
    before = os.sched_getaffinity(0)
    r = shard.bind_to_gpu_cpus(0)
    assert isinstance(r, dict) and "bound" in r
    after = os.sched_getaffinity(0)
    assert after and after <= before                   # never widens the set, never leaves the rank without a CPU
    if not r["bound"]:
        assert after == before and r.get("why")
    os.sched_setaffinity(0, before)


def test_bam_header_from_file_prefix(tmp_path):
    """bam.read_bam_header_of_file: a header spread over several BGZF blocks, read from growing prefixes of the file."""
    import pytest
    from falcon_unzip_b200 import bam
    refs = [("ctg%05d" % i, 1000 + i) for i in range(6000)]
    text = "@HD\tVN:1.5\n" + "".join("@RG\tID:r%d\n" % i for i in range(3000))
    hdr = bam.bam_header_bytes(text, refs)
    rec = bam.encode_record(0, 5, "r", 0, 254, [(4, "=")], "ACGT")
    fn = str(tmp_path / "h.bam")
    with open(fn, "wb") as f:
        for o in range(0, len(hdr), 40000):
            f.write(bam._bgzf_block(hdr[o:o + 40000], 1))
        f.write(bam._bgzf_block(rec * 1000, 1))
        f.write(bam._BGZF_EOF)
    want_text, want_refs, _recs = bam.read_bam(fn)
    for first in (100, 5000, 1 << 20):
        got_text, got_refs = bam.read_bam_header_of_file(fn, first=first)
        assert got_text == want_text and list(got_refs) == list(want_refs), first
    cut = str(tmp_path / "cut.bam")
    with open(cut, "wb") as f:
        f.write(open(fn, "rb").read()[:3000])
    with pytest.raises(ValueError):
        bam.read_bam_header_of_file(cut)
    with open(cut, "wb") as f:
        f.write(b"not a bam file at all, just text" * 10)
    with pytest.raises(ValueError):
        bam.read_bam_header_of_file(cut)
