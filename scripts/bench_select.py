"""Per-contig read selection (select_reads_from_bam.py): raw-read BAMs -> record bytes grouped by
contig.  Device leg: Engine.ingest_bam + name_rows + gather_records (what falcon_unzip_b200.select_reads_from_bam runs per
input file, without the host BGZF compression of the outputs, which both sides share); CPU leg: the per-record loop of
oracle/select_oracle.py on the inflated records.  Usage: bench_select.py [reads] [mean read length]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine  # noqa: E402
from falcon_unzip_b200 import select_reads_from_bam as srb  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    mean = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
    rng = np.random.default_rng(1)
    n_ctg = 40
    names = ["m54006_1/%d/0_%d" % (4000 + i, 9000 + i) for i in range(n)]
    lens = np.clip(rng.normal(mean, 0.2 * mean, n), 500, None).astype(np.int64)
    # records: core + name + 4-bit SEQ + QUAL, built with numpy (no per-base Python)
    parts, sizes = [], []
    for i in range(n):
        l_seq = int(lens[i])
        nb = names[i].encode() + b"\0"
        body = np.zeros(32 + len(nb) + (l_seq + 1) // 2 + l_seq, np.uint8)
        body[0:4] = np.frombuffer(np.int32(-1).tobytes(), np.uint8)
        body[4:8] = np.frombuffer(np.int32(-1).tobytes(), np.uint8)
        body[8] = len(nb)
        body[10:12] = np.frombuffer(np.uint16(4680).tobytes(), np.uint8)
        body[14:16] = np.frombuffer(np.uint16(4).tobytes(), np.uint8)
        body[16:20] = np.frombuffer(np.int32(l_seq).tobytes(), np.uint8)
        body[20:24] = np.frombuffer(np.int32(-1).tobytes(), np.uint8)
        body[24:28] = np.frombuffer(np.int32(-1).tobytes(), np.uint8)
        body[32:32 + len(nb)] = np.frombuffer(nb, np.uint8)
        o = 32 + len(nb)
        body[o:o + (l_seq + 1) // 2] = rng.integers(0, 256, (l_seq + 1) // 2, dtype=np.uint8) & 0x33 | 0x11
        body[o + (l_seq + 1) // 2:] = 0xFF
        parts.append(np.frombuffer(np.int32(len(body)).tobytes(), np.uint8))
        parts.append(body)
        sizes.append(len(body) + 4)
    records = np.concatenate(parts)
    d = tempfile.mkdtemp(prefix="fuz_sel_")
    fn = os.path.join(d, "subreads.bam")
    bam.write_bam(fn, [], records.tobytes(), header_text="@HD\tVN:1.5\n@RG\tID:a\n", level=1)
    image = torch.from_numpy(np.fromfile(fn, dtype=np.uint8)).pin_memory().numpy()
    # 90 % of the reads are assigned to one of n_ctg contigs
    known = rng.random(n) < 0.9
    ctg = rng.integers(0, n_ctg, n)
    by_name = sorted((names[i].encode(), int(ctg[i])) for i in range(n) if known[i])
    keys = np.array([b for b, _c in by_name], dtype="S")
    key_ctg = np.array([c for _b, c in by_name], dtype=np.int64)
    table = {b.decode(): c for b, c in by_name}
    eng = engine.get_engine(0)
    ts = []
    for _ in range(4):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        data, bounds = srb.partition_file(eng, image, keys, key_ctg, n_ctg)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t_dev = min(ts[1:])
    # CPU: host inflate (zlib, all cores) + the per-record loop of the oracle
    t0 = time.perf_counter()
    buf = bytes(bam.read_bam(fn)[2])
    t_inflate = time.perf_counter() - t0
    t0 = time.perf_counter()
    off = bam.index_records(buf)
    out = [[] for _ in range(n_ctg)]
    for i in range(len(off) - 1):
        rec = buf[off[i]:off[i + 1]]
        c = table.get(rec[36:36 + rec[12] - 1].decode("latin-1"))
        if c is not None:
            out[c].append(rec)
    cpu = [b"".join(x) for x in out]
    t_loop = time.perf_counter() - t0
    for c in range(n_ctg):
        assert data[int(bounds[c]):int(bounds[c + 1])].tobytes() == cpu[c], c
    print(json.dumps({"reads": n, "bam_bytes": int(len(image)), "record_bytes": int(len(records)), "selected_bytes": int(len(data)),
                      "contigs": n_ctg, "device_ms": 1e3 * t_dev, "device_GBps_of_records": len(records) / t_dev / 1e9,
                      "cpu_inflate_ms": 1e3 * t_inflate, "cpu_loop_ms": 1e3 * t_loop, "cpu_cores_inflate": os.cpu_count(),
                      "what": "BAM file image in pinned memory -> record bytes grouped by contig on the host (Engine.ingest_bam + name_rows + "
                              "gather_records); CPU: zlib inflate on all cores + the oracle's per-record loop, one thread; outputs compared"}))


if __name__ == "__main__":
    main()
