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
