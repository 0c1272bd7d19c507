"""Synthetic diploid contigs + truth-aligned reads, emitted directly as BAM records.

Inputs of BASELINE.json's configs: haplotype h0 iid uniform ACGT,
het sites uniform without replacement at ``het_rate`` with a uniform alternative base,
reference = h0, reads drawn from a Bernoulli(1/2) haplotype at uniform starts with
N(mu, 0.2 mu) lengths, iid errors split evenly between substitution / 1-base insertion /
1-base deletion, truth alignment written as a CIGAR with ``=``/``X``/``I``/``D`` (what
``blasr --bam`` emits, reference falcon_unzip/unzip.py:86-88) so blasr is not needed.

Everything is vectorised numpy over all bases of a batch of reads; the output is the
concatenated *uncompressed BAM alignment records* (see bam.py) that both the CUDA path
(verbatim) and the CPU oracle (as ``samtools view`` text or as records) consume.
"""
from __future__ import annotations

import dataclasses
from typing import List, Tuple

import numpy as np

from . import bam

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_NIB = np.array([1, 2, 4, 8], dtype=np.uint8)  # BAM 4-bit codes of A C G T
OP_M, OP_I, OP_D, OP_N, OP_S, OP_H, OP_P, OP_EQ, OP_X = range(9)

_CORE_DT = np.dtype([("block_size", "<i4"), ("refid", "<i4"), ("pos", "<i4"),
                     ("l_name", "u1"), ("mapq", "u1"), ("bin", "<u2"), ("n_cig", "<u2"),
                     ("flag", "<u2"), ("l_seq", "<i4"), ("nref", "<i4"), ("npos", "<i4"),
                     ("tlen", "<i4")])
assert _CORE_DT.itemsize == bam.CORE_BYTES


@dataclasses.dataclass
class SynthConfig:
    name: str = "c1"
    n_contigs: int = 1
    contig_len: int = 1_000_000
    coverage: float = 30.0
    mean_read_len: int = 10_000
    het_rate: float = 1e-3
    error_rate: float = 0.01
    seed: int = 20240601
    cigar_style: str = "=X"        # "=X" (blasr --bam) or "M"
    min_read_len: int = 2500
    frac_softclip: float = 0.0     # reads with soft clips at both ends
    frac_heavy_clip: float = 0.0   # reads clipped > 90 % (dropped by phasing.py:72)
    frac_short: float = 0.0        # reads with total CIGAR length < 2000 (phasing.py:74)
    frac_dup_name: float = 0.0     # reads re-using the previous read's QNAME
    n_base_rate: float = 0.0       # query bases replaced by N
    first_contig: int = 0


# Named configurations of BASELINE.json (seed = 20240601 + config number).
CONFIGS = {
    "c1": SynthConfig("c1", 1, 1_000_000, 30.0, 10_000, seed=20240602),
    "c2": SynthConfig("c2", 20, 250_000, 40.0, 10_000, seed=20240603),
    "c3": SynthConfig("c3", 2000, 67_500, 50.0, 10_000, seed=20240604),
    "c5": SynthConfig("c5", 125, 2_000_000, 60.0, 15_000, seed=20240606),
    # small parity-test shapes
    "tiny": SynthConfig("tiny", 2, 30_000, 24.0, 6_000, seed=7, min_read_len=2500),
    "quirks": SynthConfig("quirks", 3, 40_000, 30.0, 6_000, seed=11, frac_softclip=0.05,
                          frac_heavy_clip=0.02, frac_short=0.03, frac_dup_name=0.03,
                          n_base_rate=0.002),
    # raw-read error rates: CIGARs of ~1000 operations per read, match segments of a few bases
    "noisy": SynthConfig("noisy", 2, 30_000, 30.0, 6_000, seed=13, error_rate=0.12, n_base_rate=0.001),
    "noisy_m": SynthConfig("noisy_m", 2, 30_000, 30.0, 6_000, seed=17, error_rate=0.15, cigar_style="M"),
    # ultra-long, accurate reads: match segments of several kb, reads of ~100 kb
    "long": SynthConfig("long", 1, 400_000, 20.0, 100_000, seed=19, error_rate=0.002, cigar_style="M", min_read_len=20_000),
}


@dataclasses.dataclass
class SynthSet:
    config: SynthConfig
    refs: List[Tuple[str, int]]          # (name, length) per contig, refid order
    ref_seqs: List[str]                  # reference sequence (= haplotype 0)
    het_pos: List[np.ndarray]            # planted het positions (0-based) per contig
    records: np.ndarray                  # uint8, concatenated BAM records, all contigs
    rec_off: np.ndarray                  # int64 [n_rec + 1]
    rec_ctg: np.ndarray                  # int32 [n_rec] contig index of each record

    def contig_records(self, c: int) -> bytes:
        idx = np.flatnonzero(self.rec_ctg == c)
        if len(idx) == 0:
            return b""
        return self.records[self.rec_off[idx[0]]:self.rec_off[idx[-1] + 1]].tobytes()


def contig_name(i: int) -> str:
    return "%06dF" % i


def _reg2bin_vec(beg: np.ndarray, end: np.ndarray) -> np.ndarray:
    end = end - 1
    out = np.zeros(len(beg), dtype=np.int64)
    done = np.zeros(len(beg), dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = ~done & ((beg >> shift) == (end >> shift))
        out[hit] = base + (beg[hit] >> shift)
        done |= hit
    return out


def _ragged_index(dst_start: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Destination indices of a ragged copy: segment k goes to dst_start[k] + [0, lens[k])."""
    total = int(lens.sum())
    seg_off = np.cumsum(lens) - lens
    return np.repeat(dst_start - seg_off, lens) + np.arange(total, dtype=np.int64)


def _gen_batch(rng: np.random.Generator, cfg: SynthConfig, refid: int, haps: np.ndarray,
               starts: np.ndarray, lens: np.ndarray, hap_of: np.ndarray,
               clip_l: np.ndarray, clip_r: np.ndarray, names: List[bytes]):
    """BAM records for one batch of reads of one contig (already in output order)."""
    n = len(starts)
    lens = lens.astype(np.int64)
    N = int(lens.sum())
    roff = np.cumsum(lens) - lens
    read_of = np.repeat(np.arange(n, dtype=np.int32), lens)
    tpos = np.repeat(starts - roff, lens) + np.arange(N, dtype=np.int64)
    tb = haps[hap_of[read_of], tpos]                      # template base codes 0..3
    u = rng.random(N)
    e3 = cfg.error_rate / 3.0
    ev = np.zeros(N, dtype=np.int8)                       # 0 match 1 sub 2 ins 3 del
    ev[u < 3 * e3] = 3
    ev[u < 2 * e3] = 2
    ev[u < e3] = 1
    last = roff + lens - 1
    ev[roff] = 0
    ev[last] = 0
    is_sub, is_ins, is_del = ev == 1, ev == 2, ev == 3

    own_base = tb.copy()
    own_base[is_sub] = (tb[is_sub] + rng.integers(1, 4, int(is_sub.sum()))) % 4

    # query bases: [pre extras][own base unless deleted][post extras]
    pre = is_ins.astype(np.int64)
    pre[roff] += clip_l
    post = np.zeros(N, dtype=np.int64)
    post[last] += clip_r
    own = (~is_del).astype(np.int64)
    qn = pre + own + post
    qoff = np.cumsum(qn) - qn
    Q = int(qn.sum())
    qb = rng.integers(0, 4, Q).astype(np.uint8)           # extras are random bases
    keep = ~is_del
    qb[(qoff + pre)[keep]] = own_base[keep]
    l_seq = np.add.reduceat(qn, roff)

    # CIGAR op units: [pre op (S at read start, else I)] [own op] [post op S]
    if cfg.cigar_style == "M":
        own_op = np.where(is_del, OP_D, OP_M).astype(np.uint8)
    else:
        own_op = np.where(is_del, OP_D, np.where(is_sub, OP_X, OP_EQ)).astype(np.uint8)
    has_pre, has_post = pre > 0, post > 0
    un = has_pre.astype(np.int64) + 1 + has_post.astype(np.int64)
    uoff = np.cumsum(un) - un
    U = int(un.sum())
    u_op = np.empty(U, dtype=np.uint8)
    u_len = np.ones(U, dtype=np.int64)
    u_read = np.repeat(read_of, un)
    pre_op = np.full(N, OP_I, dtype=np.uint8)
    pre_op[roff[clip_l > 0]] = OP_S
    u_op[uoff[has_pre]] = pre_op[has_pre]
    u_len[uoff[has_pre]] = pre[has_pre]
    own_slot = uoff + has_pre
    u_op[own_slot] = own_op
    u_op[(own_slot + 1)[has_post]] = OP_S
    u_len[(own_slot + 1)[has_post]] = post[has_post]
    chg = np.ones(U, dtype=bool)
    chg[1:] = (u_op[1:] != u_op[:-1]) | (u_read[1:] != u_read[:-1])
    run_start = np.flatnonzero(chg)
    run_len = np.add.reduceat(u_len, run_start)
    run_op = u_op[run_start]
    run_read = u_read[run_start]
    n_cig = np.bincount(run_read, minlength=n).astype(np.int64)
    if n_cig.max() > 65535:
        raise ValueError("CIGAR with more than 65535 operations is not representable")
    cig_words = ((run_len << 4) | run_op).astype("<u4")

    # 4-bit packed SEQ, every read starting on a byte boundary
    nib = _NIB[qb]
    if cfg.n_base_rate > 0:
        nib[rng.random(Q) < cfg.n_base_rate] = 15
    seq_bytes = (l_seq + 1) // 2
    seq_off = np.cumsum(seq_bytes) - seq_bytes
    q_read = np.repeat(np.arange(n, dtype=np.int64), l_seq)
    q_first = np.cumsum(l_seq) - l_seq
    k = np.arange(Q, dtype=np.int64) - q_first[q_read]
    dst = seq_off[q_read] + (k >> 1)
    packed = np.zeros(int(seq_bytes.sum()), dtype=np.uint8)
    hi = (k & 1) == 0
    packed[dst[hi]] = nib[hi] << 4
    packed[dst[~hi]] |= nib[~hi]

    # assemble records: core | name\0 | cigar | seq | qual (0xff) ; no aux
    name_len = np.array([len(x) + 1 for x in names], dtype=np.int64)
    body = 32 + name_len + 4 * n_cig + seq_bytes + l_seq
    rec_len = body + 4
    rec_off = np.concatenate([[0], np.cumsum(rec_len)]).astype(np.int64)
    buf = np.full(int(rec_off[-1]), 0xFF, dtype=np.uint8)
    core = np.zeros(n, dtype=_CORE_DT)
    core["block_size"] = body
    core["refid"] = refid
    core["pos"] = starts
    core["l_name"] = name_len
    core["mapq"] = 254
    core["bin"] = _reg2bin_vec(starts.astype(np.int64), starts.astype(np.int64) + lens)
    core["n_cig"] = n_cig
    core["flag"] = np.where(rng.random(n) < 0.5, 0, 16)
    core["l_seq"] = l_seq
    core["nref"] = -1
    core["npos"] = -1
    core["tlen"] = 0
    buf[_ragged_index(rec_off[:-1], np.full(n, bam.CORE_BYTES, np.int64))] = \
        core.view(np.uint8)
    name_cat = np.frombuffer(b"".join(x + b"\0" for x in names), dtype=np.uint8)
    p = rec_off[:-1] + bam.CORE_BYTES
    buf[_ragged_index(p, name_len)] = name_cat
    p = p + name_len
    buf[_ragged_index(p, 4 * n_cig)] = cig_words.view(np.uint8)
    p = p + 4 * n_cig
    buf[_ragged_index(p, seq_bytes)] = packed
    return buf, rec_off


def generate(cfg: SynthConfig, batch_bases: int = 4_000_000) -> SynthSet:
    """Generate every contig of ``cfg`` (records coordinate-sorted within a contig)."""
    refs, ref_seqs, het_all, chunks, ctg_of = [], [], [], [], []
    for ci in range(cfg.first_contig, cfg.first_contig + cfg.n_contigs):
        rng = np.random.Generator(np.random.PCG64([cfg.seed, ci]))
        L = cfg.contig_len
        h0 = rng.integers(0, 4, L).astype(np.uint8)
        n_het = int(round(L * cfg.het_rate))
        het = np.sort(rng.choice(L, size=n_het, replace=False))
        h1 = h0.copy()
        h1[het] = (h0[het] + rng.integers(1, 4, n_het)) % 4
        haps = np.stack([h0, h1])
        n_reads = int(np.ceil(cfg.coverage * L / cfg.mean_read_len))
        lens = np.clip(rng.normal(cfg.mean_read_len, 0.2 * cfg.mean_read_len, n_reads),
                       cfg.min_read_len, L).astype(np.int64)
        kind = rng.random(n_reads)
        short = kind < cfg.frac_short
        lens[short] = rng.integers(300, 1900, int(short.sum()))
        starts = (rng.random(n_reads) * (L - lens + 1)).astype(np.int64)
        order = np.argsort(starts, kind="stable")
        starts, lens, kind = starts[order], lens[order], kind[order]
        hap_of = rng.integers(0, 2, n_reads).astype(np.int8)
        clip_l = np.zeros(n_reads, dtype=np.int64)
        clip_r = np.zeros(n_reads, dtype=np.int64)
        soft = (kind >= cfg.frac_short) & (kind < cfg.frac_short + cfg.frac_softclip)
        clip_l[soft] = rng.integers(1, 400, int(soft.sum()))
        clip_r[soft] = rng.integers(0, 400, int(soft.sum()))
        heavy = (kind >= cfg.frac_short + cfg.frac_softclip) & \
                (kind < cfg.frac_short + cfg.frac_softclip + cfg.frac_heavy_clip)
        # exactly at / just beyond the 90 % clip boundary (phasing.py:72)
        hv = np.flatnonzero(heavy)
        clip_l[hv] = 9 * lens[hv] + rng.integers(0, 3, len(hv)) - 1
        names = []
        read_serial = ci * 1_000_000          # unique across contigs, independent of batch composition
        for i in range(n_reads):
            if i > 0 and cfg.frac_dup_name > 0 and rng.random() < cfg.frac_dup_name:
                names.append(names[-1])
            else:
                names.append(b"m%08d/%d/0_%d" % (read_serial, read_serial, int(lens[i])))
            read_serial += 1
        # batches bound the working set of the vectorised generator
        b0 = 0
        csum = np.cumsum(lens + clip_l + clip_r)
        while b0 < n_reads:
            limit = (csum[b0 - 1] if b0 else 0) + batch_bases
            b1 = max(b0 + 1, int(np.searchsorted(csum, limit, side="right")))
            buf, off = _gen_batch(rng, cfg, ci - cfg.first_contig, haps, starts[b0:b1],
                                  lens[b0:b1], hap_of[b0:b1], clip_l[b0:b1], clip_r[b0:b1],
                                  names[b0:b1])
            chunks.append((buf, off))
            ctg_of.append(np.full(b1 - b0, ci - cfg.first_contig, dtype=np.int32))
            b0 = b1
        refs.append((contig_name(ci), L))
        ref_seqs.append(_ACGT[h0].tobytes().decode("ascii"))
        het_all.append(het)
    records = np.concatenate([c[0] for c in chunks]) if chunks else np.zeros(0, np.uint8)
    offs, base = [np.zeros(1, np.int64)], 0
    for buf, off in chunks:
        offs.append(off[1:] + base)
        base += len(buf)
    return SynthSet(cfg, refs, ref_seqs, het_all, records, np.concatenate(offs),
                    np.concatenate(ctg_of) if ctg_of else np.zeros(0, np.int32))


def write_fasta(path: str, sset: SynthSet, width: int = 80) -> None:
    with open(path, "w") as f:
        for (name, _l), seq in zip(sset.refs, sset.ref_seqs):
            f.write(">%s\n" % name)
            for o in range(0, len(seq), width):
                f.write(seq[o:o + width] + "\n")


def _gen_one(args):
    cfg, ci = args
    return generate(dataclasses.replace(cfg, n_contigs=1, first_contig=ci))


def generate_parallel(cfg: SynthConfig, workers: int = 0) -> SynthSet:
    """generate() with one process per contig (call before CUDA is initialised: fork)."""
    import multiprocessing as mp
    import os
    workers = workers or min(cfg.n_contigs, os.cpu_count() or 1)
    if workers <= 1 or cfg.n_contigs == 1:
        return generate(cfg)
    with mp.get_context("fork").Pool(workers) as pool:
        parts = pool.map(_gen_one, [(cfg, ci) for ci in range(cfg.first_contig, cfg.first_contig + cfg.n_contigs)])
    refs, ref_seqs, het, recs, offs, ctgs, base = [], [], [], [], [np.zeros(1, np.int64)], [], 0
    for c, p in enumerate(parts):
        refs += p.refs; ref_seqs += p.ref_seqs; het += p.het_pos
        buf = p.records.copy()
        # refID of the record = contig index inside the merged set
        idx = p.rec_off[:-1, None] + 4 + np.arange(4)[None, :]
        buf[idx] = np.frombuffer(np.int32(c).tobytes(), dtype=np.uint8)[None, :]
        recs.append(buf)
        offs.append(p.rec_off[1:] + base)
        base += len(buf)
        ctgs.append(np.full(len(p.rec_off) - 1, c, dtype=np.int32))
    return SynthSet(cfg, refs, ref_seqs, het, np.concatenate(recs), np.concatenate(offs), np.concatenate(ctgs))


def generate_contigs(cfg: SynthConfig, contig_ids, workers: int = 0):
    """Yield (contig id, SynthSet of that one contig) for the given GLOBAL contig ids, in the given order, from a pool
    of processes (call before CUDA is initialised: fork).  A contig's content depends on (seed, id) only, so every
    partition of a contig list over ranks sees the same contigs as one rank generating all of them."""
    import multiprocessing as mp
    import os
    ids = [int(i) for i in contig_ids]
    workers = workers or min(len(ids), os.cpu_count() or 1)
    if workers <= 1 or len(ids) <= 1:
        for ci in ids:
            yield ci, _gen_one((cfg, ci))
        return
    with mp.get_context("fork").Pool(workers) as pool:
        for ci, part in zip(ids, pool.imap(_gen_one, [(cfg, ci) for ci in ids])):
            yield ci, part


# --------------------------------------------------------------------------- fast generator (bench workloads)
_synth_lib = None


def _fast_lib():
    """libfuz_synth.so (csrc/fuz_synth.cpp): the same read model with its own random stream, ~100x faster than the
    vectorised numpy generator; used for the full-size bench workloads (15 G aligned bases)."""
    global _synth_lib
    if _synth_lib is None:
        import ctypes as C
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libfuz_synth.so")
        if not os.path.exists(path):
            raise RuntimeError("libfuz_synth.so is not built (make -C falcon_unzip_b200/csrc)")
        lib = C.CDLL(path)
        lib.fuz_synth_bounds.restype = C.c_int64
        lib.fuz_synth_bounds.argtypes = [C.c_int64, C.c_double, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
        lib.fuz_synth_contig.restype = C.c_int64
        lib.fuz_synth_contig.argtypes = [C.c_uint64, C.c_int64, C.c_int32, C.c_int64, C.c_double, C.c_int64, C.c_int64, C.c_double,
                                         C.c_double, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                                         C.c_void_p, C.POINTER(C.c_int64)]
        _synth_lib = lib
    return _synth_lib


def fast_supported(cfg: SynthConfig) -> bool:
    return (cfg.cigar_style == "=X" and cfg.frac_softclip == 0 and cfg.frac_heavy_clip == 0 and cfg.frac_short == 0
            and cfg.frac_dup_name == 0 and cfg.n_base_rate == 0)


def generate_contig_fast(cfg: SynthConfig, ci: int) -> SynthSet:
    """One contig (global id ci, refID 0 in its records) from the native generator."""
    import ctypes as C
    lib = _fast_lib()
    n_reads = C.c_int64(0)
    cap = lib.fuz_synth_bounds(cfg.contig_len, cfg.coverage, cfg.mean_read_len, cfg.min_read_len, C.byref(n_reads))
    buf = np.empty(cap, np.uint8)
    off = np.empty(n_reads.value + 1, np.int64)
    ref = np.empty(cfg.contig_len, np.uint8)
    n_rec, aligned = C.c_int64(0), C.c_int64(0)
    n = lib.fuz_synth_contig(cfg.seed, ci, 0, cfg.contig_len, cfg.coverage, cfg.mean_read_len, cfg.min_read_len, cfg.het_rate,
                             cfg.error_rate, buf.ctypes.data, cap, off.ctypes.data, len(off) - 1, C.byref(n_rec), ref.ctypes.data,
                             C.byref(aligned))
    if n < 0:
        raise RuntimeError("fuz_synth_contig failed (%d)" % n)
    return SynthSet(dataclasses.replace(cfg, n_contigs=1, first_contig=ci), [(contig_name(ci), cfg.contig_len)],
                    [ref.tobytes().decode("ascii")], [np.zeros(0, np.int64)], buf[:n], off[:n_rec.value + 1],
                    np.zeros(n_rec.value, np.int32))


def generate_contigs_fast(cfg: SynthConfig, contig_ids, threads: int = 0):
    """Yield (contig id, SynthSet) in the given order; the contigs are generated on a pool of threads (the native call
    releases the GIL).  The content depends on (seed, id) only."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    ids = [int(i) for i in contig_ids]
    threads = threads or min(len(ids), os.cpu_count() or 1)
    if threads <= 1 or len(ids) <= 1:
        for ci in ids:
            yield ci, generate_contig_fast(cfg, ci)
        return
    with ThreadPoolExecutor(threads) as ex:
        window = 2 * threads                                   # bounded look-ahead: a contig is ~200 MB
        futs = {}
        nxt = 0
        for k, ci in enumerate(ids):
            while nxt < len(ids) and nxt < k + window:
                futs[nxt] = ex.submit(generate_contig_fast, cfg, ids[nxt])
                nxt += 1
            yield ci, futs.pop(k).result()
