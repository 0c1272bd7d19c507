"""run_track_reads_sharded on two GPUs over NCCL (skipped with fewer than two devices): the
all-gather of kept overlap lines runs on device tensors, each rank merges and votes its own
targets with fuz_rr_track; rawread_to_contigs must equal the oracle's bytes."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, rr, paths, port, bestn):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from falcon_unzip_b200 import engine, rr_hctg_track
    engine.get_engine(rank)
    rr_hctg_track.read_las_lines = lambda db_fn, fn: iter(rr.las_lines[fn])
    info = rr_hctg_track.run_track_reads_sharded(paths["phased"], paths["r2c"], paths["ids"], list(rr.las_lines), 2500, bestn,
                                                 "raw_reads.db", paths["out"], rank, world)
    assert info["kept_total"] >= info["kept_local"] > 0
    dist.destroy_process_group()


def test_rr_two_gpus_nccl_equals_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from test_rr_shard_gloo import _write
    from falcon_unzip_b200 import synth_rr
    from oracle import rr_oracle
    rr = synth_rr.generate_rr(n_reads=4000, n_ctg=5, ctg_len=150_000, n_files=5, seed=77)
    paths = _write(rr, str(tmp_path))
    want = rr_oracle.run_track_reads(rr.las_lines, rr.phased_reads, rr.read_to_contig_map, rr.rawread_ids, 2500, 40)
    port = 33500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, rr, paths, port, 40), nprocs=2, join=True)
    assert len(want.splitlines()) > 1000
    assert open(paths["out"]).read() == want
