"""Multi-process overlap filter (run_ovlp_filter_sharded) at world size 2 over gloo on CPU.  The device
call is replaced by a host stand-in with the semantics of fuz_ovlp_filter (built from the restated oracle's
stage functions), so this covers the host logic of the N > 1 path: dealing LAS files to the ranks, the two
set unions (all-reduce), per-file text collection and its order.  The result must equal the oracle's
single-process output byte for byte."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp


def _fake_device_filter(L, tab, max_diff, max_ovlp, min_ovlp, min_len, bestn, stage, ignore_in=None, contained_in=None, d=None):
    from oracle import ovlp_oracle
    a2p = {"%09d" % r: (str(tab.ctg[r]), str(tab.blk[r]), str(tab.ph[r])) for r in np.flatnonzero(tab.in_map)}
    lines_by_file = {}
    for i in range(L.n):
        lines_by_file.setdefault(int(L.file[i]), []).append((i, " ".join(L.tokens(i))))
    ids = lambda flags: set("%09d" % r for r in np.flatnonzero(flags))
    ignore = np.zeros(tab.n_reads, np.uint8) if ignore_in is None else np.asarray(ignore_in)
    if ignore_in is None:
        for f in sorted(lines_by_file):
            for x in ovlp_oracle.stage1([t for _i, t in lines_by_file[f]], a2p, max_diff, max_ovlp, min_ovlp, min_len):
                if x is not None:
                    ignore[int(x)] = 1
    out = dict(n_groups=0, ignore=ignore, contained=np.zeros(tab.n_reads, np.uint8), grp_q=np.zeros(0, np.int32),
               grp_line=np.zeros(0, np.int32), grp_ignore=np.zeros(0, np.uint8), grp_tie=np.zeros(0, np.uint8),
               grp_off=np.zeros(1, np.int32), out_line=np.zeros(0, np.int32), cand=None)
    if stage == 1:
        return out
    contained = out["contained"] if contained_in is None else np.asarray(contained_in)
    if contained_in is None:
        for f in sorted(lines_by_file):
            for x in ovlp_oracle.stage2([t for _i, t in lines_by_file[f]], a2p, min_len, ids(ignore)):
                contained[int(x)] = 1
    out["contained"] = contained
    if stage == 2:
        return out
    sel = []
    for f in sorted(lines_by_file):
        index = {}
        for i, t in lines_by_file[f]:
            index.setdefault(t, []).append(i)
        for l in ovlp_oracle.stage3([t for _i, t in lines_by_file[f]], a2p, min_len, ids(ignore), ids(contained), bestn):
            sel.append(index[" ".join(l[:-2])][0])
    out["out_line"] = np.asarray(sel, np.int32)
    return out


def _worker(rank, world, s, port, p, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from falcon_unzip_b200 import ovlp_filter_with_phase as ofp
    ofp._device_filter = _fake_device_filter
    ofp.read_las_lines = lambda db_fn, fn: ("\n".join(s.las_lines[fn]) + "\n").encode() if s.las_lines[fn] else b""
    ofp.arid2phase.clear()
    ofp.arid2phase.update({r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows})
    text = ofp.run_ovlp_filter_sharded(list(s.las_lines), "db", p["max_diff"], p["max_cov"], p["min_cov"], p["min_len"], p["bestn"],
                                       rank, world)
    assert (text is None) == (rank != 0)
    if rank == 0:
        open(out_path, "wb").write(text)
    dist.destroy_process_group()


def test_ovlp_world_size_2_gloo_equals_oracle(tmp_path):
    from falcon_unzip_b200 import synth_rr
    from oracle import ovlp_oracle
    s = synth_rr.generate_ovlp(n_reads=500, n_files=3, seed=41)
    p = dict(max_diff=120, max_cov=120, min_cov=1, min_len=2500, bestn=10)
    a2p = {r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows}
    want = ovlp_oracle.run_filter(list(s.las_lines.items()), a2p, **p)
    port = 34500 + os.getpid() % 2000
    out = str(tmp_path / "out.txt")
    mp.spawn(_worker, args=(2, s, port, p, out), nprocs=2, join=True)
    assert len(want) > 1000
    assert open(out).read() == want
