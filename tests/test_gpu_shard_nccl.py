"""Contig sharding over two GPUs through the product path (shard.phase_bam_sharded, one process per GPU, NCCL only for
the bookkeeping gather): the files gathered from both ranks equal the files of a single-GPU run byte for byte (contigs
are independent units, reference unzip.py:231-281).  Skipped on a box with fewer than two GPUs."""
import filecmp
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, bam_fn, fa_fn, out_dir, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from falcon_unzip_b200 import shard
    res = shard.phase_bam_sharded(bam_fn, fa_fn, out_dir, rank, world, device=rank)
    assert len(res["mine"]) > 0 and len(res["all"]) == res["n_contigs"]
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_files_equal_single_gpu_files(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import dataclasses
    from falcon_unzip_b200 import bam, phasing, synth
    sset = synth.generate(dataclasses.replace(synth.CONFIGS["c2"], n_contigs=6, contig_len=120_000))
    bam_fn, fa_fn = str(tmp_path / "in.bam"), str(tmp_path / "ref.fa")
    bam.write_bam(bam_fn, sset.refs, sset.records.tobytes())
    synth.write_fasta(fa_fn, sset)
    port = 34500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, bam_fn, fa_fn, str(tmp_path / "two"), port), nprocs=2, join=True)
    _res, one = phasing.phase_bam(bam_fn, fa_fn, str(tmp_path / "one"))
    for name, files in one.items():
        for kind, path in files.items():
            other = path.replace(str(tmp_path / "one"), str(tmp_path / "two"))
            assert os.path.exists(other), other
            assert filecmp.cmp(path, other, shallow=False), (name, kind)
