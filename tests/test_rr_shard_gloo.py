"""Multi-process raw-read tracking (run_track_reads_sharded) at world size 2 over gloo on CPU.
The device call is replaced by a host stand-in with the semantics of fuz_rr_track (filter, per
file heaps, merge, vote), so this covers the host logic of the N > 1 path: dealing LAS files to
ranks, the all-gather of kept lines, the target shard, the collection of rows and their order.
The result must equal the oracle's single-process output byte for byte."""
import os
from heapq import heappush, heappushpop

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from falcon_unzip_b200 import py2compat


def _fake_track_device(q, t, ln, tl, file_idx, tab, min_len, bestn, filter_only=False):
    n = tab.n_reads
    keep = np.zeros(len(q), np.uint8)
    for i in range(len(q)):
        qi, ti = int(q[i]), int(t[i])
        if tl[i] < min_len or not tab.in_map[qi]:
            continue
        if tab.ph_ctg[ti] >= 0 and tab.ph_block[ti] != -1 and tab.ph_ctg[qi] >= 0 and tab.ph_ctg[qi] == tab.ph_ctg[ti] \
                and tab.ph_block[qi] == tab.ph_block[ti] and tab.ph_phase[qi] != tab.ph_phase[ti]:
            continue
        keep[i] = 1
    empty = np.zeros(0, np.int32)
    if filter_only:
        return keep, None, None, None, np.zeros(n + 1, np.int32), empty, empty, np.zeros(0, np.int64)

    def offer(h, item):
        if len(h) < bestn:
            heappush(h, item)
        else:
            heappushpop(h, item)
    per_file = {}
    for i in np.flatnonzero(keep):
        offer(per_file.setdefault(int(file_idx[i]), {}).setdefault(int(t[i]), []), (int(ln[i]), "%09d" % int(q[i])))
    merged = {}
    for f in sorted(per_file):
        for ks in py2compat.str_dict_order(["%09d" % k for k in per_file[f]]):
            for item in per_file[f][int(ks)]:
                offer(merged.setdefault(int(ks), []), item)
    vt_off, vt_ctg, vt_count, vt_score = np.zeros(n + 1, np.int32), [], [], []
    for tid in range(n):
        score = {}
        for s, rid in merged.get(tid, []):
            r = int(rid)
            for c in tab.rc_ctg[tab.rc_off[r]:tab.rc_off[r + 1]].tolist():
                sc = score.setdefault(c, [0, 0])
                sc[0] += -s
                sc[1] += 1
        for c, (s, cnt) in score.items():
            vt_ctg.append(c); vt_score.append(s); vt_count.append(cnt)
        vt_off[tid + 1] = len(vt_ctg)
    return (keep, None, None, None, vt_off, np.array(vt_ctg, np.int32), np.array(vt_count, np.int32),
            np.array(vt_score, np.int64))


def _worker(rank, world, rr, paths, port, bestn):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from falcon_unzip_b200 import rr_hctg_track
    rr_hctg_track._track_device = _fake_track_device
    rr_hctg_track.read_las_lines = lambda db_fn, fn: iter(rr.las_lines[fn])
    info = rr_hctg_track.run_track_reads_sharded(paths["phased"], paths["r2c"], paths["ids"], list(rr.las_lines), 2500, bestn,
                                                 "raw_reads.db", paths["out"], rank, world)
    assert info["kept_total"] >= info["kept_local"] > 0
    dist.destroy_process_group()


def _write(rr, d):
    p = dict(phased=os.path.join(d, "all_phased_reads"), r2c=os.path.join(d, "read_to_contig_map"),
             ids=os.path.join(d, "rawread_ids"), out=os.path.join(d, "out", "rawread_to_contigs"))
    open(p["phased"], "w").write("".join(l + "\n" for l in rr.phased_reads))
    open(p["r2c"], "w").write("".join(l + "\n" for l in rr.read_to_contig_map))
    open(p["ids"], "w").write(rr.rawread_ids)
    return p


def test_rr_world_size_2_gloo_equals_oracle(tmp_path):
    from falcon_unzip_b200 import synth_rr
    from oracle import rr_oracle
    rr = synth_rr.generate_rr(n_reads=700, n_ctg=3, ctg_len=60_000, n_files=3, seed=31)
    paths = _write(rr, str(tmp_path))
    want = rr_oracle.run_track_reads(rr.las_lines, rr.phased_reads, rr.read_to_contig_map, rr.rawread_ids, 2500, 5)
    port = 31500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, rr, paths, port, 5), nprocs=2, join=True)
    got = open(paths["out"]).read()
    assert len(want.splitlines()) > 100
    assert got == want
