"""Drop-in mirror of reference falcon_unzip/phasing.py: same function names, the same
``self`` convention (file attributes + ``parameters`` dict, phasing.py:16-23,139-142,
217-219,425-432), the same CLI (phasing.py:557-570) and byte-identical output files under
``<base_dir>/<ctg_id>/...`` (phasing.py:501-503,520,534,543) -- with every count, index
and vote computed by the CUDA kernels of libfuz.so.  There is no CPU fallback.

Differences that are deliberate and documented in DESIGN.md:
  * the BAM is decoded natively (BGZF inflate + record split); ``--samtools`` is accepted
    and ignored.  A SAM text file is accepted in place of the BAM (the reference's own
    input after the ``samtools view`` pipe, phasing.py:27,42-59).
  * records the reference would crash on (CIGAR ``*``, SEQ ``*``) raise RuntimeError
    instead of ZeroDivisionError / IndexError; unsorted input is rejected.
  * ``phase_contigs`` runs the four stages for many contigs in one fused device call.
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
from types import SimpleNamespace
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import bam, engine, formats
from ._lib import FuzError, FUZ_E_FORMAT

BASE_INDEX = {"A": 0, "C": 1, "G": 2, "T": 3}
_ACTG_RANK = {0: 0, 1: 1, 3: 2, 2: 3}        # order of the string "ACTG" 


def fn(p):
    """pypeflow's fn(): path of a file handle; plain strings pass through."""
    return getattr(p, "path", p)


def makePypeLocalFile(path):
    return SimpleNamespace(path=path)


class _open_out:
    """Output file that only appears under its name when it was written completely: text goes to `<path>.tmp`, which
    replaces `path` when the block ends without an exception (pypeFLOW takes the existence of a file for success; the
    reference writes in place, phasing.py:39-40,132)."""

    def __init__(self, path: str, mode: str = "w"):
        self.path = path
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        self.f = open(path + ".tmp", mode)

    def __enter__(self):
        return self.f

    def __exit__(self, exc_type, exc, tb):
        self.f.close()
        if exc_type is None:
            os.replace(self.path + ".tmp", self.path)
        else:
            try:
                os.unlink(self.path + ".tmp")
            except OSError:
                pass
        return False


# --------------------------------------------------------------------------- inputs
def load_contig_records(bam_fn: str, ctg_id: str, ref_len: int = 0) -> Tuple[np.ndarray, int]:
    """Records `samtools view <bam> <ctg_id>` would print, as raw BAM records."""
    if bam.is_bgzf(bam_fn):
        _text, refs, recs = bam.read_bam(bam_fn)
        names = [r[0] for r in refs]
        if ctg_id not in names:
            return np.zeros(0, np.uint8), ref_len
        rid = names.index(ctg_id)
        records = np.frombuffer(recs, dtype=np.uint8)
        off = engine.index_records(records)
        refid = engine.record_refids(records, off) if len(off) > 1 else np.zeros(0, np.int32)
        sel = np.flatnonzero(refid == rid)
        if len(sel) == 0:
            return np.zeros(0, np.uint8), max(ref_len, refs[rid][1])
        if np.any(np.diff(sel) != 1):
            raise FuzError(5, "records of %s are not contiguous in %s (BAM not sorted)" % (ctg_id, bam_fn))
        sub = records[off[sel[0]]:off[sel[-1] + 1]].copy()
        # refID inside the sub-batch is contig 0
        sub_off = off[sel[0]:sel[-1] + 2] - off[sel[0]]
        idx = sub_off[:-1, None] + 4 + np.arange(4)[None, :]
        sub[idx] = 0
        return sub, max(ref_len, refs[rid][1])
    with open(bam_fn) as f:
        recs = bam.records_from_sam_lines(f, [(ctg_id, ref_len)])
    records = np.frombuffer(recs, dtype=np.uint8).copy()
    if len(records):
        off = engine.index_records(records)
        idx = off[:-1, None] + 4 + np.arange(4)[None, :]
        records[idx] = 0          # every line belongs to the requested contig (like the pipe)
    return records, ref_len


def _contig_len(records: np.ndarray, rec_off: np.ndarray, ref_len: int) -> int:
    if len(rec_off) > 1:
        idx = rec_off[:-1, None] + 8 + np.arange(4)[None, :]
        max_pos = int(records[idx].copy().view("<i4").max())
        ref_len = max(ref_len, max_pos + 1)
    return max(ref_len, 1)


def parse_variant_map(path: str):
    """-> sites (pos, ref, al[2] ACTG-ordered), rows (vm_site, vm_base, vm_qid)."""
    site_pos: List[int] = []
    site_ref: List[str] = []
    site_alleles: List[List[int]] = []
    vm_site, vm_base, vm_qid = [], [], []
    seen: Dict[Tuple[int, str], int] = {}
    last_key = None
    with open(path) as f:
        for line in f:
            l = line.strip().split()
            if not l:
                continue
            key = (int(l[0]), l[1])
            if key != last_key:
                if key in seen:
                    raise FuzError(FUZ_E_FORMAT, "variant_map rows of site %d are not contiguous" % key[0])
                seen[key] = len(site_pos)
                site_pos.append(key[0]); site_ref.append(key[1]); site_alleles.append([])
                last_key = key
            b = BASE_INDEX.get(l[2])
            if b is None:
                raise FuzError(FUZ_E_FORMAT, "variant_map allele %r is not one of ACGT" % l[2])
            if b not in site_alleles[-1]:
                site_alleles[-1].append(b)
            vm_site.append(len(site_pos) - 1); vm_base.append(b); vm_qid.append(int(l[3]))
    for p, al in zip(site_pos, site_alleles):
        if len(al) != 2:
            raise FuzError(FUZ_E_FORMAT, "variant_map site %d carries %d alleles; the reference needs 2" % (p, len(al)))
        al.sort(key=_ACTG_RANK.get)
    pos = np.asarray(site_pos, np.int32)
    if len(pos) > 1 and np.any(np.diff(pos) <= 0):
        raise FuzError(FUZ_E_FORMAT, "variant_map positions must be strictly ascending")
    return (pos, site_ref, np.asarray(site_alleles, np.uint8).reshape(-1, 2), np.asarray(vm_site, np.int32),
            np.asarray(vm_base, np.uint8), np.asarray(vm_qid, np.int32))


def _stage_outputs(eng: engine.Engine, n_sites: int, n_vmap: int, n_atable: int, n_reads: int):
    caps = dict(sites=max(n_sites, 16), vmap=max(n_vmap, 16), atable=max(n_atable, 16), reads=max(n_reads, 16))
    return caps


# --------------------------------------------------------------------------- stage 1
def make_het_call(self):
    """reference phasing.py:14-134."""
    bam_fn = fn(self.bam_file)
    ctg_id = self.parameters["ctg_id"]
    ref_seq = self.parameters["ref_seq"]
    base_dir = self.parameters["base_dir"]
    vmap_fn, vpos_fn, q_id_map_fn = fn(self.vmap_file), fn(self.vpos_file), fn(self.q_id_map_file)
    try:
        os.makedirs("%s/%s" % (base_dir, ctg_id))
    except OSError:
        pass
    records, ref_len = load_contig_records(bam_fn, ctg_id, len(ref_seq))
    rec_off = engine.index_records(records)
    pb = engine.prepare_batch(records, [ctg_id], [_contig_len(records, rec_off, ref_len)], rec_off=rec_off,
                              ctg_rec_off=np.asarray([0, len(rec_off) - 1], np.int32))
    res = engine.get_engine().phase_device(pb, stage="het")
    if res.n_sites and int(res.site_pos.max()) > len(ref_seq):
        raise IndexError("string index out of range (ref_seq shorter than a het position; phasing.py:123)")
    with _open_out(vpos_fn) as f:
        f.write(formats.variant_pos_text(res, 0, res.n_sites, ref_seq))
    with _open_out(vmap_fn) as f:
        f.write(formats.variant_map_text(res, 0, 0, res.n_vmap, ref_seq))
    with _open_out(q_id_map_fn) as f:
        f.write(formats.q_id_map_text(pb.qnames(0)))


# --------------------------------------------------------------------------- stage 2
def generate_association_table(self):
    """reference phasing.py:137-206."""
    vmap_fn, atable_fn = fn(self.vmap_file), fn(self.atable_file)
    pos, _ref, al, vm_site, vm_base, vm_qid = parse_variant_map(vmap_fn)
    eng = engine.get_engine()
    n_sites, n_vmap = len(pos), len(vm_site)

    def run(do):
        do.set("site_ctg", np.zeros(n_sites, np.int32)); do.set("site_pos", pos); do.set("site_al", al)
        do.set("vm_site", vm_site); do.set("vm_base", vm_base); do.set("vm_qid", vm_qid)
        eng._torch.cuda.synchronize(eng.device)
        eng.association_async(1, n_sites, n_vmap, do)
    caps = dict(sites=max(n_sites, 16), vmap=max(n_vmap, 16), atable=max(16, 24 * n_sites), reads=16)
    do, st = eng._retry(caps, 0, run)
    a = do.fetch(st, ("atable",))
    res = SimpleNamespace(site_pos=pos, site_al=al, at_s1=a["at_s1"], at_s2=a["at_s2"], at_ct=a["at_ct"])
    with _open_out(atable_fn) as f:
        f.write(formats.atable_text(res, 0, int(st.n_atable)))


# --------------------------------------------------------------------------- stage 3
def get_score(c_score, pos1, pos2, s1, s2):
    """reference phasing.py:208-214: evidence for the states s1, s2 (allele pairs) of two sites from the table of
    co-occurrence scores.  The device evaluates the same lookup on packed edges (csrc/fuz_phase.cu, k_ctg_phase); this
    host form exists for callers of the module-level name."""
    (lo_pos, lo_state), (hi_pos, hi_state) = sorted(((pos1, s1), (pos2, s2)), key=lambda site: site[0])
    return c_score[(lo_pos, hi_pos)][(lo_state[0] + hi_state[0], lo_state[1] + hi_state[1])]


def get_phased_blocks(self):
    """reference phasing.py:216-421."""
    vmap_fn, atable_fn, p_variant_fn = fn(self.vmap_file), fn(self.atable_file), fn(self.phased_variant_file)
    ref_base: Dict[int, str] = {}
    with open(vmap_fn) as f:
        for line in f:
            l = line.strip().split()
            if l:
                ref_base[int(l[0])] = l[1]                       # phasing.py:230-238
    rows = []
    alleles: Dict[int, Tuple[int, int]] = {}
    with open(atable_fn) as f:
        for line in f:
            l = line.strip().split()
            if not l:
                continue
            p1, p2 = int(l[0]), int(l[3])
            for p, a, b in ((p1, l[1], l[2]), (p2, l[4], l[5])):
                if a not in BASE_INDEX or b not in BASE_INDEX or a == b:
                    raise FuzError(FUZ_E_FORMAT, "atable alleles %r %r at %d" % (a, b, p))
                pair = (BASE_INDEX[a], BASE_INDEX[b])
                if alleles.setdefault(p, pair) != pair:
                    raise FuzError(FUZ_E_FORMAT, "atable lists different allele pairs for position %d" % p)
            rows.append((p1, p2, int(l[6]), int(l[7]), int(l[8]), int(l[9])))
    pos = np.asarray(sorted(alleles), np.int32)
    index = {int(p): i for i, p in enumerate(pos.tolist())}
    n_sites, n_at = len(pos), len(rows)
    al = np.asarray([alleles[int(p)] for p in pos.tolist()], np.uint8).reshape(-1, 2)
    at_s1 = np.asarray([index[r[0]] for r in rows], np.int32)
    at_s2 = np.asarray([index[r[1]] for r in rows], np.int32)
    at_ct = np.asarray([r[2:] for r in rows], np.int32).reshape(-1, 4)
    eng = engine.get_engine()
    caps = dict(sites=max(n_sites, 16), vmap=16, atable=max(n_at, 16), reads=16)
    do = eng.alloc_outputs(caps)
    do.set("site_ctg", np.zeros(n_sites, np.int32)); do.set("site_pos", pos); do.set("site_al", al)
    do.set("at_s1", at_s1); do.set("at_s2", at_s2); do.set("at_ct", at_ct)
    eng._torch.cuda.synchronize(eng.device)
    eng.blocks_async(1, n_sites, n_at, do)
    st = eng.status()
    a = do.fetch(st, ("sites",))
    res = SimpleNamespace(site_pos=pos, site_al=al, **{k: a[k][:n_sites] for k in (
        "ph_state", "ph_lext", "ph_rext", "ph_lscore", "ph_rscore", "ph_block")})
    with _open_out(p_variant_fn) as f:
        f.write(formats.phased_variants_text(res, 0, n_sites, ref_base))


# --------------------------------------------------------------------------- stage 4
def get_phased_reads(self):
    """reference phasing.py:423-480."""
    q_id_map_fn, vmap_fn = fn(self.q_id_map_file), fn(self.vmap_file)
    p_variant_fn, phased_read_fn = fn(self.phased_variant_file), fn(self.phased_read_file)
    ctg_id = self.parameters["ctg_id"]
    rid_map: Dict[int, str] = {}
    with open(q_id_map_fn) as f:
        for line in f:
            l = line.strip().split()
            if l:
                rid_map[int(l[0])] = l[1]
    pos, ref, al, vm_site, vm_base, vm_qid = parse_variant_map(vmap_fn)
    n_sites, n_vmap = len(pos), len(vm_site)
    index = {(int(p), r): i for i, (p, r) in enumerate(zip(pos.tolist(), ref))}
    ph_block = np.zeros(n_sites, np.int32)
    ph_state = np.full(n_sites, 255, np.uint8)
    with open(p_variant_fn) as f:
        for line in f:
            l = line.strip().split()
            if not l or l[0] != "V":
                continue
            k3, k4 = l[3].split("_"), l[4].split("_")
            if k3[:2] != k4[:2]:
                raise FuzError(FUZ_E_FORMAT, "phased_variants row mixes two sites: %s" % line.strip())
            s = index.get((int(k3[0]), k3[1]))
            if s is None:
                continue                                   # phases a variant no read carries
            h0, h1 = BASE_INDEX.get(k3[2]), BASE_INDEX.get(k4[2])
            if {h0, h1} != set(al[s].tolist()) or ph_block[s] != 0:
                raise FuzError(FUZ_E_FORMAT, "phased_variants row does not match variant_map: %s" % line.strip())
            ph_block[s] = int(l[1])
            ph_state[s] = 0 if h0 == al[s, 0] else 1
    nq = int(vm_qid.max()) + 1 if n_vmap else 0
    eng = engine.get_engine()

    def run(do):
        do.set("site_ctg", np.zeros(n_sites, np.int32)); do.set("site_pos", pos); do.set("site_al", al)
        do.set("vm_site", vm_site); do.set("vm_base", vm_base); do.set("vm_qid", vm_qid)
        do.set("ph_block", ph_block); do.set("ph_state", ph_state)
        eng._torch.cuda.synchronize(eng.device)
        eng.reads_async(1, np.asarray([nq], np.int32), n_sites, n_vmap, do)
    caps = dict(sites=max(n_sites, 16), vmap=max(n_vmap, 16), atable=16, reads=max(16, 2 * nq))
    do, st = eng._retry(caps, 0, run)
    a = do.fetch(st, ("reads",))
    res = SimpleNamespace(vm_qid=vm_qid, **{k: a[k] for k in ("pr_qid", "pr_block", "pr_phase", "pr_n0", "pr_n1")})
    with _open_out(phased_read_fn) as f:
        f.write(formats.phased_reads_text(res, 0, int(st.n_reads), 0, n_vmap, ctg_id, rid_map))


# --------------------------------------------------------------------------- driver / CLI
def phasing(args):
    """reference phasing.py:482-553: the four stages, serially, chained through files."""
    bam_fn, fasta_fn, ctg_id, base_dir, samtools = args.bam, args.fasta, args.ctg_id, args.base_dir, args.samtools
    ref_seq = ""
    for name, seq in bam.read_fasta(fasta_fn):
        rid = name.split()[0]
        if rid != ctg_id:
            continue
        ref_seq = seq.upper()
    vmap_file = makePypeLocalFile(os.path.join(base_dir, ctg_id, "het_call", "variant_map"))
    vpos_file = makePypeLocalFile(os.path.join(base_dir, ctg_id, "het_call", "variant_pos"))
    q_id_map_file = makePypeLocalFile(os.path.join(base_dir, ctg_id, "het_call", "q_id_map"))
    atable_file = makePypeLocalFile(os.path.join(base_dir, ctg_id, "g_atable", "atable"))
    phased_variant_file = makePypeLocalFile(os.path.join(base_dir, ctg_id, "get_phased_blocks", "phased_variants"))
    phased_read_file = makePypeLocalFile(os.path.join(base_dir, ctg_id, "phased_reads"))
    make_het_call(SimpleNamespace(
        bam_file=makePypeLocalFile(bam_fn), vmap_file=vmap_file, vpos_file=vpos_file, q_id_map_file=q_id_map_file,
        parameters=dict(ctg_id=ctg_id, ref_seq=ref_seq, base_dir=base_dir, samtools=samtools)))
    generate_association_table(SimpleNamespace(
        vmap_file=vmap_file, atable_file=atable_file, parameters=dict(ctg_id=ctg_id, base_dir=base_dir)))
    get_phased_blocks(SimpleNamespace(
        vmap_file=vmap_file, atable_file=atable_file, phased_variant_file=phased_variant_file, parameters={}))
    get_phased_reads(SimpleNamespace(
        vmap_file=vmap_file, q_id_map_file=q_id_map_file, phased_variant_file=phased_variant_file,
        phased_read_file=phased_read_file, parameters=dict(ctg_id=ctg_id)))


def write_contig_files(res, sl, c: int, ctg_id: str, ref_seq: str, names: Sequence[str], base_dir: str) -> Dict[str, str]:
    """All six files of contig c of a batch result (layout of phasing.py:501-503,520,534,543)."""
    s0, s1 = int(sl["site"][c]), int(sl["site"][c + 1])
    v0, v1 = int(sl["vmap"][c]), int(sl["vmap"][c + 1])
    a0, a1 = int(sl["atable"][c]), int(sl["atable"][c + 1])
    r0, r1 = int(sl["reads"][c]), int(sl["reads"][c + 1])
    base = os.path.join(base_dir, ctg_id)
    paths = dict(variant_map=os.path.join(base, "het_call", "variant_map"),
                 variant_pos=os.path.join(base, "het_call", "variant_pos"),
                 q_id_map=os.path.join(base, "het_call", "q_id_map"),
                 atable=os.path.join(base, "g_atable", "atable"),
                 phased_variants=os.path.join(base, "get_phased_blocks", "phased_variants"),
                 phased_reads=os.path.join(base, "phased_reads"))
    # every file as bytes straight from the formatters of libfuz (names: list of str, or the QNAME rows of a device batch)
    ref = ref_seq if isinstance(ref_seq, bytes) else ref_seq.encode("latin-1")
    text = dict(variant_pos=formats.variant_pos_bytes(res, s0, s1, ref),
                variant_map=formats.variant_map_bytes(res, s0, v0, v1, ref),
                q_id_map=formats.q_id_map_bytes(names),
                atable=formats.atable_bytes(res, a0, a1),
                phased_variants=formats.phased_variants_bytes(res, s0, s1, ref),
                phased_reads=formats.phased_reads_bytes(res, r0, r1, v0, v1, ctg_id, names))
    for k, p in paths.items():
        with _open_out(p, "wb") as f:          # never leaves a partial output file behind
            f.write(text[k])
    return paths


def phase_contigs(records, ctg_names: Sequence[str], ref_seqs: Sequence[str], base_dir: str,
                  device: int = 0, host_path: bool = True, ctg_rec_off=None, max_batch_bytes: int = 6 << 30):
    """Fused path: every contig of the list in one device call (or, for lists beyond the limits of one call --
    engine.plan_batches -- in consecutive batches), then the per-contig files.
    records: concatenated BAM records grouped by contig, in the order of ctg_names (grouping from
    the refID fields 0..n-1, or given explicitly as record offsets ctg_rec_off).
    Returns (PhaseResult of the batch -- a list of them when the contigs were split --, {contig: {file kind: path}})."""
    if not isinstance(records, np.ndarray):
        records = np.frombuffer(records, dtype=np.uint8)
    lens = [len(s) for s in ref_seqs]
    eng = engine.get_engine(device)
    rec_off = engine.index_records(records)
    if ctg_rec_off is None:
        refid = engine.record_refids(records, rec_off) if len(rec_off) > 1 else np.zeros(0, np.int32)
        if len(refid) and (np.any(np.diff(refid) < 0) or refid.min() < 0 or refid.max() >= len(ctg_names)):
            raise FuzError(5, "records are not grouped by reference id 0..%d" % (len(ctg_names) - 1))
        ctg_rec_off = np.searchsorted(refid, np.arange(len(ctg_names) + 1), side="left")
    ctg_rec_off = np.asarray(ctg_rec_off, np.int64)
    ctg_bytes = rec_off[ctg_rec_off[1:]] - rec_off[ctg_rec_off[:-1]]
    plan = engine.plan_batches(ctg_bytes, np.diff(ctg_rec_off), lens, max_batch_bytes)
    out, results = {}, []
    for lo, hi in plan:
        sub_rec, sub_off, sub_cro = engine.sub_batch(records, rec_off, ctg_rec_off, lo, hi)
        pb = engine.prepare_batch(sub_rec, ctg_names[lo:hi], lens[lo:hi], rec_off=sub_off, ctg_rec_off=sub_cro,
                                  assign_qids=not host_path)   # host entry: q_ids are assigned on the device
        res = eng.phase_host(pb) if host_path else eng.phase_device(pb)
        sl = formats.contig_slices(res, pb.n_ctg)
        for c in range(pb.n_ctg):
            name = ctg_names[lo + c]
            out[name] = write_contig_files(res, sl, c, name, ref_seqs[lo + c], pb.qnames(c), base_dir)
        results.append(res)
    return (results[0] if len(results) == 1 else results), out


def phase_bam(bam_fn, fasta_fn: str, base_dir: str, device: int = 0, verify_crc: bool = True):
    """Every contig of a coordinate-sorted BAM (or of a list of BAMs, e.g. one per contig) in one go, decoded on the device: the file image is
    uploaded as it is, BGZF inflate + record split (the `samtools view` pipe of phasing.py:27), the
    QNAME -> q_id table and the four stages run in HBM; the same six files per contig come out."""
    eng = engine.get_engine(device)
    # The device side (file image into a reused page-locked buffer, upload, inflate, record index, the four stages, rows back)
    # runs on a worker thread -- its calls into libfuz / the CUDA runtime release the interpreter -- while this thread parses
    # the FASTA as bytes.
    import threading
    box = {}

    def _device_side():
        try:
            import torch
            torch.cuda.set_device(device)                 # the worker's current device (libfuz launches on the caller's)
            if isinstance(bam_fn, (list, tuple)):         # one sorted BAM per contig, as unzip.py:90 leaves them
                image = [np.fromfile(fn_, dtype=np.uint8) for fn_ in bam_fn]
            else:
                image = eng.read_file_pinned(bam_fn)
            box["out"] = eng.phase_bam(image, verify_crc=verify_crc)
        except BaseException as e:                        # noqa: BLE001 -- re-raised in the caller's thread
            box["error"] = e
    worker = threading.Thread(target=_device_side)
    worker.start()
    try:
        ref_seqs = {n.split()[0]: s.upper() for n, s in bam.read_fasta_bytes(fasta_fn)}
    finally:
        worker.join()
    if "error" in box:
        raise box["error"]
    res, info = box["out"]
    return res, write_batch_files(res, info, [ref_seqs.get(n, b"") for n in info.ctg_names], base_dir)


def write_batch_files(res, info, ref_seqs: Sequence[str], base_dir: str):
    """The six files of every contig of a device batch (engine.BamBatchInfo: contig names and QNAME rows) -> {contig: paths}."""
    sl = formats.contig_slices(res, info.n_ctg)

    def one(c):
        return write_contig_files(res, sl, c, info.ctg_names[c], ref_seqs[c], info.qname_rows(c), base_dir)
    if info.n_ctg < 4:
        paths = [one(c) for c in range(info.n_ctg)]
    else:
        # formatting (libfuz) and file I/O release the interpreter: a few threads write the contigs side by side
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(8, info.n_ctg, os.cpu_count() or 1)) as pool:
            paths = list(pool.map(one, range(info.n_ctg)))
    return dict(zip(info.ctg_names, paths))


def parse_args(argv):
    parser = argparse.ArgumentParser(description="phasing variants and reads from a bam file")
    parser.add_argument("--bam", type=str, help="path to sorted bam file", required=True)
    parser.add_argument("--fasta", type=str, help="path to the fasta file of contain the contig", required=True)
    parser.add_argument("--ctg_id", type=str, help="contig identifier in the bam file", required=True)
    parser.add_argument("--base_dir", type=str, default="./",
                        help="the output base_dir, default to current working directory")
    parser.add_argument("--samtools", type=str, default="samtools",
                        help="path to samtools (accepted for compatibility; the BAM is decoded natively)")
    return parser.parse_args(argv[1:])


def main(argv=sys.argv):
    logging.basicConfig()
    args = parse_args(argv)
    phasing(args)
