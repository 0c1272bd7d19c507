"""Synthetic inputs of the per-contig read selection (select_reads_from_bam.py): raw-read BAMs (unaligned records),
rawread_ids, rawread_to_contigs.  Shared by the CPU and the GPU tests."""
import os

import numpy as np


def make_case(root: str, seed: int = 7, n_reads: int = 420, n_files: int = 3):
    """-> (fofn path, rawread_to_contigs path, rawread_ids path).  Covers: contigs with 20 / 21 / few reads (the > 20
    rule), 'NA' rows, rank >= 1 rows, reads with two rank-0 rows (lower score wins, then the contig name), a read whose
    best contig is not selected while its second one is, reads missing from the table, names missing from the BAMs,
    duplicated QNAMEs, an input file without any selected read, a record longer than 64 KiB, three different headers."""
    from falcon_unzip_b200 import bam
    rng = np.random.default_rng(seed)
    os.makedirs(os.path.join(root, "in"), exist_ok=True)
    names = ["m54006_%d/%d/0_%d" % (seed, 1000 + i, 500 + 7 * i) for i in range(n_reads)]
    ids_fn = os.path.join(root, "rawread_ids")
    with open(ids_fn, "w") as f:
        f.write("\n".join(names + ["ghost/1/0_10", "ghost/2/0_10"]) + "\n")
    # contig of every read (rank 0): blocks of reads
    plan = [("000000F", 120), ("000001F", 60), ("000002F_001", 15), ("000003F", 20), ("000004F", 21), ("NA", 30)]
    rows, i = [], 0
    for ctg, cnt in plan:
        for _ in range(cnt):
            score = -int(rng.integers(1000, 90000))
            rows.append("%09d %s %d 0 %d %d" % (i, ctg, int(rng.integers(1, 40)), score, int(rng.integers(0, 2))))
            rows.append("%09d %s %d 1 %d 0" % (i, "000009F", 3, score + 5))                     # rank 1: ignored
            i += 1
    first_free = i
    # two rank-0 rows: lower score wins; equal scores: contig name decides
    rows.append("%09d 000000F 5 0 -500 1" % i); rows.append("%09d 000001F 5 0 -900 1" % i); i += 1
    rows.append("%09d 000001F 5 0 -700 1" % i); rows.append("%09d 000000F 5 0 -700 1" % i); i += 1
    # best contig not selected (000002F_001 has 15 reads), second one is: the read is dropped
    rows.append("%09d 000002F_001 5 0 -9000 1" % i); rows.append("%09d 000000F 5 0 -100 1" % i); i += 1
    # the same row twice: one entry in the set
    rows.append("%09d 000004F 5 0 -800 1" % i); rows.append("%09d 000004F 5 0 -800 1" % i); i += 1
    # names that no BAM holds
    rows.append("%09d 000000F 5 0 -800 1" % n_reads); rows.append("%09d 000001F 5 0 -800 1" % (n_reads + 1))
    order = rng.permutation(len(rows))
    r2c_fn = os.path.join(root, "rawread_to_contigs")
    with open(r2c_fn, "w") as f:
        f.write("".join(rows[k] + "\n" for k in order))
    # BAM files: reads dealt to the files at random, a few twice, a few names the table does not know
    headers = []
    for k in range(n_files):
        h = "@HD\tVN:1.5\tSO:unknown\tpb:3.0.1\n@RG\tID:rg%d\tPL:PACBIO\tDS:READTYPE=SUBREAD\tPU:movie%d\n" % (k, k)
        h += "@PG\tID:bax2bam-%d\tPN:bax2bam\tVN:0.0.8\n" % k
        if k == 0:
            h += "@CO\tfirst file only\n"
        if k == 1:
            h += "@RG\tID:rg1b\tPL:PACBIO\tPU:movie1b\n"
        headers.append(h)
    per_file = [[] for _ in range(n_files)]
    for j, name in enumerate(names):
        if j < first_free + 4 or j % 3 == 0:
            per_file[int(rng.integers(0, n_files - 1))].append(name)            # the last file gets no known read
    for j in range(0, 60, 7):
        per_file[int(rng.integers(0, n_files - 1))].append(names[j])              # duplicated QNAME
    for k in range(n_files):
        per_file[k] += ["unknown/%d/%d_9" % (k, t) for t in range(25)]
    fns = []
    for k in range(n_files):
        lst = [per_file[k][t] for t in rng.permutation(len(per_file[k]))]
        recs = []
        for t, name in enumerate(lst):
            l_seq = 100000 if (k == 0 and t == 5) else int(rng.integers(1, 3000))
            seq = "".join("ACGT"[b] for b in rng.integers(0, 4, l_seq))
            recs.append(bam.encode_record(-1, -1, name, 4, 255, [], seq, aux=b"zmi" + int(t).to_bytes(4, "little")))
        fn = os.path.join(root, "in", "movie%d.subreads.bam" % k)
        bam.write_bam(fn, [], b"".join(recs), header_text=headers[k], level=1)
        fns.append(fn)
    fofn = os.path.join(root, "in", "input_bam.fofn")
    with open(fofn, "w") as f:
        f.write("movie0.subreads.bam\n%s\nmovie2.subreads.bam\n" % fns[1])          # relative and absolute paths
    return fofn, r2c_fn, ids_fn
