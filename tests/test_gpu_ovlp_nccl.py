"""run_ovlp_filter_sharded on two GPUs over NCCL (skipped with fewer than two devices): the two set unions
are all-reduces of device tensors, every stage runs in fuz_ovlp_filter; the text must equal the oracle's."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, s, port, p, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from falcon_unzip_b200 import engine, ovlp_filter_with_phase as ofp
    engine.get_engine(rank)
    ofp.read_las_lines = lambda db_fn, fn: ("\n".join(s.las_lines[fn]) + "\n").encode() if s.las_lines[fn] else b""
    ofp.arid2phase.clear()
    ofp.arid2phase.update({r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows})
    text = ofp.run_ovlp_filter_sharded(list(s.las_lines), "db", p["max_diff"], p["max_cov"], p["min_cov"], p["min_len"], p["bestn"],
                                       rank, world)
    if rank == 0:
        open(out_path, "wb").write(text)
    dist.destroy_process_group()


def test_ovlp_two_gpus_nccl_equals_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from falcon_unzip_b200 import synth_rr
    from oracle import ovlp_oracle
    s = synth_rr.generate_ovlp(n_reads=2000, n_files=5, seed=43)
    p = dict(max_diff=120, max_cov=120, min_cov=1, min_len=2500, bestn=10)
    a2p = {r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows}
    want = ovlp_oracle.run_filter(list(s.las_lines.items()), a2p, **p)
    port = 35500 + os.getpid() % 2000
    out = str(tmp_path / "out.txt")
    mp.spawn(_worker, args=(2, s, port, p, out), nprocs=2, join=True)
    assert len(want) > 10000 and open(out).read() == want
