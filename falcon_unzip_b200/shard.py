"""Contig sharding across the GPUs of one box.  Contigs are independent units of the phasing
path (the reference runs one fc_phasing.py process per contig, unzip.py:231-281), so the
steady state needs no collective: every rank phases its own contigs and writes their files;
only the bookkeeping (who did what) is gathered."""
from __future__ import annotations

import os
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np


def assign_contigs(weights: Sequence[float], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first: contigs sorted by weight (record bytes / aligned bases),
    each given to the currently lightest rank.  Deterministic; every rank computes the same."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    load = [0.0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += weights[i]
    return [sorted(x) for x in out]


def bind_to_gpu_cpus(device_index: int) -> Dict[str, object]:
    """One process per GPU: run this process on the CPUs next to its GPU (NVML's CPU affinity of the device, cut to what
    the process may use), BEFORE any page-locked buffer is allocated.  Page-locked host memory is placed by first touch,
    so the records a rank uploads then sit in the memory of the socket its GPU hangs off; without it all ranks of a box
    may read through one socket's memory controllers and the inter-socket link.  Best effort: returns what was done."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[device_index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else device_index
        h = nv.nvmlDeviceGetHandleByIndex(phys)
        allowed = os.sched_getaffinity(0)
        n_words = (max(max(allowed) + 1, os.cpu_count() or 1) + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, n_words)
        near = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus = near & allowed
        if not cpus:
            return {"bound": False, "why": "no allowed CPU next to the device", "allowed": len(allowed), "near": len(near)}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "cpus": len(cpus), "first_cpu": min(cpus), "allowed": len(allowed)}
    except Exception as e:                                  # noqa: BLE001 -- no NVML / no permission: run unbound
        return {"bound": False, "why": "%s: %s" % (type(e).__name__, e)}


def contig_record_ranges(records: np.ndarray, rec_off: np.ndarray, n_ctg: int):
    """Record range and byte weight of every contig of a coordinate-sorted record buffer."""
    from . import engine
    refid = engine.record_refids(records, rec_off) if len(rec_off) > 1 else np.zeros(0, np.int32)
    bounds = np.searchsorted(refid, np.arange(n_ctg + 1), side="left")
    weights = [float(rec_off[bounds[c + 1]] - rec_off[bounds[c]]) for c in range(n_ctg)]
    return bounds, weights


def phase_bam_sharded(bam_fn: str, fasta_fn: str, base_dir: str, rank: int = 0, world_size: int = 1,
                      device: Optional[int] = None, phase_fn: Optional[Callable] = None) -> Dict[str, object]:
    """Every rank phases its LPT share of the contigs of `bam_fn` with one fused device call and
    writes the per-contig files under base_dir.  With torch.distributed initialised the list of
    finished contigs is gathered on every rank (all_gather_object: host bookkeeping only)."""
    from . import bam, phasing
    ref_seqs = {name.split()[0]: seq.upper() for name, seq in bam.read_fasta(fasta_fn)}
    from . import engine
    if phase_fn is None and bam.is_bgzf(bam_fn):
        # product path: no rank inflates the file on the host.  The BGZF image goes to the rank's GPU as it is, is
        # decoded and indexed there (milliseconds per GB), the contigs are dealt by their record bytes (LPT) and the
        # rank's contigs are compacted into a batch of their own on the device (fuz_gather_records) and phased.
        eng = engine.get_engine(device if device is not None else 0)
        db = eng.ingest_bam(np.fromfile(bam_fn, dtype=np.uint8))
        refs = db.refs
        cro = db.ctg_rec_off.cpu().numpy().astype(np.int64)
        off_at = db.rec_off[db.ctg_rec_off.long()].cpu().numpy()
        weights = [float(off_at[c + 1] - off_at[c]) for c in range(len(refs))]
        mine = assign_contigs(weights, world_size)[rank]
        done = []
        if mine:
            sub = eng.select_contigs(db, mine)
            del db
            res, info = eng.phase_ingested(sub)
            phasing.write_batch_files(res, info, [ref_seqs.get(n, "") for n in info.ctg_names], base_dir)
            done = list(info.ctg_names)
    else:
        _text, refs, recs = bam.read_bam(bam_fn)
        records = np.frombuffer(recs, dtype=np.uint8)
        rec_off = engine.index_records(records)
        bounds, weights = contig_record_ranges(records, rec_off, len(refs))
        mine = assign_contigs(weights, world_size)[rank]
        done = []
        if mine:
            sub_names = [refs[c][0] for c in mine]
            parts, sub_off = [], [0]
            for c in mine:
                lo, hi = int(rec_off[bounds[c]]), int(rec_off[bounds[c + 1]])
                parts.append(records[lo:hi])
                sub_off.append(sub_off[-1] + (bounds[c + 1] - bounds[c]))
            sub = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
            seqs = [ref_seqs.get(n, "") for n in sub_names]
            fn = phase_fn or (lambda rec, names, sq, bd: phasing.phase_contigs(
                rec, names, sq, bd, device=device if device is not None else 0, ctg_rec_off=np.asarray(sub_off, np.int32)))
            fn(sub, sub_names, seqs, base_dir)
            done = sub_names
    gathered = [done]
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and world_size > 1:
            gathered = [None] * world_size
            dist.all_gather_object(gathered, done)
    except ImportError:
        pass
    return dict(rank=rank, mine=done, all=[c for part in gathered for c in part], n_contigs=len(refs))


def phase_bam_files_sharded(bam_fns: Sequence[str], fasta_fn: str, base_dir: str, rank: int = 0, world_size: int = 1,
                            device: Optional[int] = None, phase_fn: Optional[Callable] = None) -> Dict[str, object]:
    """The reference's layout -- one sorted BAM per contig (unzip.py:90) -- over the GPUs of a box: the files are
    dealt to the ranks by size (longest-processing-time first), every rank decodes and phases ITS files in one device
    batch (phasing.phase_bam with a list) and writes their per-contig files.  No collective on the data path; the
    list of finished files is gathered when torch.distributed is initialised."""
    from . import phasing
    fns = list(bam_fns)
    mine = [fns[i] for i in assign_contigs([float(os.path.getsize(f)) for f in fns], world_size)[rank]]
    if mine:
        fn = phase_fn or (lambda files, fa, bd: phasing.phase_bam(files, fa, bd, device=device if device is not None else 0))
        fn(mine, fasta_fn, base_dir)
    gathered = [mine]
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and world_size > 1:
            gathered = [None] * world_size
            dist.all_gather_object(gathered, mine)
    except ImportError:
        pass
    return dict(rank=rank, mine=mine, all=[f for part in gathered for f in part], n_files=len(fns))
