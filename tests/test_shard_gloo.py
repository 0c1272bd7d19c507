"""Host logic of the contig sharding: LPT assignment and a world-size-2 run over gloo (CPU)
with the device call replaced by a recorder -- every contig is phased exactly once and the
ranks agree on the gathered bookkeeping."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from falcon_unzip_b200 import shard


def test_lpt_assignment_balanced_and_deterministic():
    w = [5, 9, 1, 7, 3, 3, 8, 2]
    a = shard.assign_contigs(w, 3)
    assert sorted(i for part in a for i in part) == list(range(8))
    loads = [sum(w[i] for i in part) for part in a]
    assert max(loads) - min(loads) <= max(w)
    assert a == shard.assign_contigs(w, 3)
    assert shard.assign_contigs([4, 4], 4) == [[0], [1], [], []]


def _worker(rank, world, bam_fn, fa_fn, out_dir, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def fake_phase(records, names, seqs, base_dir):
        from falcon_unzip_b200 import engine
        off = engine.index_records(records)                 # the sub-buffer must be a valid record stream
        for n, s in zip(names, seqs):
            assert len(s) > 0
            os.makedirs(os.path.join(base_dir, n), exist_ok=True)
            with open(os.path.join(base_dir, n, "rank"), "w") as f:
                f.write("%d %d\n" % (rank, len(off) - 1))
        seen.extend(names)
    res = shard.phase_bam_sharded(bam_fn, fa_fn, out_dir, rank, world, phase_fn=fake_phase)
    assert res["mine"] == seen
    assert sorted(res["all"]) == sorted("%06dF" % i for i in range(res["n_contigs"]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    from conftest import synth_set
    from falcon_unzip_b200 import bam, synth
    sset = synth_set("quirks")
    bam_fn, fa_fn = str(tmp_path / "in.bam"), str(tmp_path / "ref.fa")
    bam.write_bam(bam_fn, sset.refs, sset.records.tobytes())
    synth.write_fasta(fa_fn, sset)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, bam_fn, fa_fn, str(tmp_path / "out"), port), nprocs=2, join=True)
    ranks = {n: open(tmp_path / "out" / n / "rank").read().split()[0] for n, _l in sset.refs}
    assert set(ranks.values()) == {"0", "1"}


def _worker_files(rank, world, fns, fa_fn, out_dir, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    seen = []

    def fake_phase(files, fa, base_dir):
        for f in files:
            os.makedirs(base_dir, exist_ok=True)
            with open(os.path.join(base_dir, os.path.basename(f) + ".rank"), "w") as g:
                g.write("%d\n" % rank)
        seen.extend(files)
    res = shard.phase_bam_files_sharded(fns, fa_fn, out_dir, rank, world, phase_fn=fake_phase)
    assert res["mine"] == seen and sorted(res["all"]) == sorted(fns) and res["n_files"] == len(fns)
    dist.barrier()
    dist.destroy_process_group()


def test_files_world_size_2_gloo(tmp_path):
    """One BAM per contig dealt to two ranks by size: every file is phased exactly once, both ranks work."""
    fns = []
    for i, size in enumerate([5000, 100, 3000, 2500, 50, 4000]):
        fn = str(tmp_path / ("%06dF_sorted.bam" % i))
        open(fn, "wb").write(b"\0" * size)
        fns.append(fn)
    port = 30500 + os.getpid() % 2000
    mp.spawn(_worker_files, args=(2, fns, str(tmp_path / "ref.fa"), str(tmp_path / "out"), port), nprocs=2, join=True)
    ranks = [open(tmp_path / "out" / (os.path.basename(f) + ".rank")).read().strip() for f in fns]
    assert set(ranks) == {"0", "1"}
    sizes = [5000, 100, 3000, 2500, 50, 4000]
    loads = [sum(s for s, r in zip(sizes, ranks) if r == k) for k in ("0", "1")]
    assert abs(loads[0] - loads[1]) <= max(sizes)
