"""select_reads_from_bam over two ranks (gloo, CPU): the device call is replaced by a host stand-in with the same
contract (record bytes grouped by contig + byte range of every contig); every selected contig is written by exactly one
rank and holds the oracle's record stream."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import select_cases
from oracle import select_oracle


def host_partition(_eng, image, keys, key_ctg, n_ctg):
    """Engine-free stand-in of select_reads_from_bam.partition_file: host inflate + record walk, same outputs."""
    import tempfile
    from falcon_unzip_b200 import bam
    with tempfile.NamedTemporaryFile(suffix=".bam") as f:
        f.write(image.tobytes())
        f.flush()
        buf = bytes(bam.read_bam(f.name)[2])
    off = bam.index_records(buf)
    table = {k: int(c) for k, c in zip(keys.tolist(), key_ctg.tolist())}
    groups = [[] for _ in range(n_ctg)]
    for i in range(len(off) - 1):
        rec = buf[off[i]:off[i + 1]]
        c = table.get(rec[36:36 + rec[12] - 1])
        if c is not None:
            groups[c].append(rec)
    data = b"".join(b"".join(g) for g in groups)
    bounds = np.concatenate([[0], np.cumsum([sum(len(r) for r in g) for g in groups])]).astype(np.int64)
    return np.frombuffer(data, np.uint8), bounds


def _worker(rank, world, fofn, r2c, ids, sam_dir, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from falcon_unzip_b200 import select_reads_from_bam as srb
    made = srb.select_reads_from_bam(fofn, r2c, ids, sam_dir, level=1, rank=rank, world_size=world, partition_fn=host_partition)
    allmade = [None] * world
    dist.all_gather_object(allmade, made)
    flat = [c for part in allmade for c in part]
    assert len(flat) == len(set(flat)) and sorted(flat) == ["000000F", "000001F", "000004F"]
    assert all(len(part) >= 1 for part in allmade)                      # both ranks write something
    dist.barrier()
    dist.destroy_process_group()


def test_select_world_size_2_gloo(tmp_path):
    from falcon_unzip_b200 import bam
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=7)
    header, want = select_oracle.select(fofn, r2c, ids)
    sam_dir = str(tmp_path / "reads")
    port = 29500 + (os.getpid() + 977) % 2000
    mp.spawn(_worker, args=(2, fofn, r2c, ids, sam_dir, port), nprocs=2, join=True)
    assert sorted(os.listdir(sam_dir)) == ["%s.bam" % c for c in sorted(want)]
    for ctg, recs in want.items():
        text, _refs, got = bam.read_bam(os.path.join(sam_dir, "%s.bam" % ctg))
        assert bytes(got) == b"".join(recs), ctg
        assert select_oracle.parse_header(text) == header


def test_single_rank_with_host_stand_in(tmp_path):
    from falcon_unzip_b200 import bam, select_reads_from_bam as srb
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=8)
    _header, want = select_oracle.select(fofn, r2c, ids)
    sam_dir = str(tmp_path / "reads")
    assert srb.select_reads_from_bam(fofn, r2c, ids, sam_dir, level=1, partition_fn=host_partition) == sorted(want)
    for ctg, recs in want.items():
        assert bytes(bam.read_bam(os.path.join(sam_dir, "%s.bam" % ctg))[2]) == b"".join(recs), ctg
