"""Drop-in mirror of reference falcon_unzip/rr_hctg_track.py (raw-read -> haplotig tracking):
same function names and arguments, same CLI (rr_hctg_track.py:174-193), byte-identical
``rawread_to_contigs`` -- the overlap filter, the per-read top-``bestn`` heaps and the contig
votes run in the CUDA kernels of libfuz.so (fuz_rr_track).  There is no CPU fallback.

Host work kept here: LA4Falcon text -> int arrays (C++ parser in libfuz), the id tables of
:15-23 and :72-85, and the row formatting of :126-138 with the CPython-2 iteration orders the
reference's output depends on; the LAS file list is sorted (upstream uses
filesystem glob order, which is not deterministic).
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import os
import shlex
import subprocess
import sys
from typing import Callable, Dict, Iterable, List, Sequence

import numpy as np

from . import _lib, engine, la4falcon, py2compat
from ._lib import FuzError, lib


class OrderedStrSet:
    """set of str that remembers insertion order and iterates in CPython-2 order."""

    def __init__(self):
        self._d: Dict[str, None] = {}

    def add(self, x: str) -> None:
        self._d[x] = None

    def __contains__(self, x) -> bool:
        return x in self._d

    def __len__(self) -> int:
        return len(self._d)

    def __iter__(self):
        return iter(py2compat.str_dict_order(self._d))


def get_rid_to_ctg(fn: str) -> Dict[str, OrderedStrSet]:
    """reference rr_hctg_track.py:15-23."""
    rid_to_ctg: Dict[str, OrderedStrSet] = {}
    with open(fn) as f:
        for row in f:
            row = row.strip().split()
            if not row:
                continue
            _pid, rid, _oid, ctg = row
            rid_to_ctg.setdefault(rid, OrderedStrSet()).add(ctg)
    return rid_to_ctg


def read_las_lines(db_fn: str, fn: str):
    """Output of ``LA4Falcon -m <db> <las>`` (reference run_tr_stage1, :25-29) as one bytes object
    (the C++ parser splits it over the host threads; an iterable of text lines works too).  Tests
    replace this function; nothing else of the module touches LA4Falcon."""
    p = subprocess.run(shlex.split("LA4Falcon -m %s %s" % (db_fn, fn)), stdout=subprocess.PIPE)
    if p.returncode != 0:
        raise RuntimeError("LA4Falcon failed on %s" % fn)
    return p.stdout


def _parse_lines(lines):
    if isinstance(lines, (bytes, bytearray, memoryview)):
        blob = bytes(lines)
    else:
        blob = "".join(l if l.endswith("\n") else l + "\n" for l in lines).encode("ascii")
    cap = blob.count(b"\n") + 1
    q, t, ln, tl = (np.empty(cap, np.int32) for _ in range(4))
    n = lib().fuz_host_parse_la4falcon(blob, len(blob), cap, q.ctypes.data, t.ctypes.data, ln.ctypes.data, tl.ctypes.data)
    if n < 0:
        raise ValueError("malformed LA4Falcon -m line (need 12 columns with integer ids / lengths)")
    return q[:n], t[:n], ln[:n], tl[:n]


class _Tables:
    def __init__(self, rid_to_ctg, rid_to_phase, n_reads: int):
        self.n_reads = n_reads
        self.ctg_names: List[str] = []
        ctg_index: Dict[str, int] = {}
        self.in_map = np.zeros(n_reads, np.uint8)
        rc_cnt = np.zeros(n_reads + 1, np.int64)
        rc_lists: Dict[int, List[int]] = {}
        for rid, ctgs in rid_to_ctg.items():
            r = int(rid)
            if "%09d" % r != rid:
                raise FuzError(_lib.FUZ_E_FORMAT, "read id %r is not a %%09d id: string order would differ from numeric" % rid)
            if r >= n_reads:
                continue                      # never probed: ids in overlaps are checked against n_reads
            idx = []
            for c in ctgs:                    # CPython-2 set iteration order
                if c not in ctg_index:
                    ctg_index[c] = len(self.ctg_names)
                    self.ctg_names.append(c)
                idx.append(ctg_index[c])
            rc_lists[r] = idx
            rc_cnt[r] = len(idx)
            self.in_map[r] = 1
        self.rc_off = np.concatenate([[0], np.cumsum(rc_cnt[:n_reads])]).astype(np.int32)
        self.rc_ctg = np.zeros(max(int(self.rc_off[-1]), 1), np.int32)
        for r, idx in rc_lists.items():
            self.rc_ctg[self.rc_off[r]:self.rc_off[r] + len(idx)] = idx
        ph_index: Dict[str, int] = {}
        self.ph_ctg = np.full(n_reads, -1, np.int32)
        self.ph_block = np.zeros(n_reads, np.int32)
        self.ph_phase = np.zeros(n_reads, np.int32)
        for r, ph in enumerate(rid_to_phase[:n_reads]):
            if ph is None:
                continue
            self.ph_ctg[r] = ph_index.setdefault(ph[0], len(ph_index))
            self.ph_block[r], self.ph_phase[r] = ph[1], ph[2]


def _track_device(q, t, ln, tl, file_idx, tab: _Tables, min_len: int, bestn: int, filter_only: bool = False, lines=None):
    """fuz_rr_track on the arrays -> (keep, hp_n, hp_len, hp_q, vt_off, vt_ctg, vt_count, vt_score).
    lines: a la4falcon.DeviceLines whose columns are on the device already (q, t, ln, tl, file_idx ignored)."""
    import torch
    eng = engine.get_engine()
    dev = eng.device
    n_reads = tab.n_reads

    def up(a, dt):
        return torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev) if len(a) else torch.zeros(1, dtype=getattr(
            torch, np.dtype(dt).name), device=dev)
    if lines is not None:
        n_ovl = lines.n
        d = dict(q=lines.d["q"], t=lines.d["t"], len=lines.d["len"], tlen=lines.d["tl"], file=lines.d["file"])
        if n_ovl:
            lo = int(torch.minimum(d["q"][:n_ovl].min(), d["t"][:n_ovl].min()).item())
            hi = int(torch.maximum(d["q"][:n_ovl].max(), d["t"][:n_ovl].max()).item())
            if lo < 0 or hi >= n_reads:
                raise IndexError("list index out of range (read id beyond rawread_ids; rr_hctg_track.py:50,54)")
    else:
        n_ovl = len(q)
        if len(q) and (int(q.min()) < 0 or int(t.min()) < 0 or int(max(q.max(), t.max())) >= n_reads):
            raise IndexError("list index out of range (read id beyond rawread_ids; rr_hctg_track.py:50,54)")
        d = dict(q=up(q, np.int32), t=up(t, np.int32), len=up(ln, np.int32), tlen=up(tl, np.int32), file=up(file_idx, np.int32))
    d.update(in_map=up(tab.in_map, np.uint8), ph_ctg=up(tab.ph_ctg, np.int32), ph_block=up(tab.ph_block, np.int32),
             ph_phase=up(tab.ph_phase, np.int32), rc_off=up(tab.rc_off, np.int32), rc_ctg=up(tab.rc_ctg, np.int32))
    b = max(bestn, 1)
    cap_votes = max(1024, 4 * n_reads)
    for _ in range(6):
        o = dict(keep=torch.zeros(max(n_ovl, 1), dtype=torch.uint8, device=dev),
                 hp_n=torch.zeros(n_reads, dtype=torch.int32, device=dev),
                 hp_len=torch.zeros(n_reads * b, dtype=torch.int32, device=dev),
                 hp_q=torch.zeros(n_reads * b, dtype=torch.int32, device=dev),
                 vt_off=torch.zeros(n_reads + 1, dtype=torch.int32, device=dev),
                 vt_ctg=torch.zeros(cap_votes, dtype=torch.int32, device=dev),
                 vt_count=torch.zeros(cap_votes, dtype=torch.int32, device=dev),
                 vt_score=torch.zeros(cap_votes, dtype=torch.int64, device=dev))
        torch.cuda.synchronize(dev)
        ri = _lib.RRInput()
        ri.n_ovl, ri.n_reads, ri.min_len, ri.bestn, ri.n_ctg = n_ovl, n_reads, min_len, bestn, len(tab.ctg_names)
        for k in ("q", "t", "len", "tlen", "file", "in_map", "ph_ctg", "ph_block", "ph_phase", "rc_off", "rc_ctg"):
            setattr(ri, "d_" + k, d[k].data_ptr())
        ro = _lib.RROutputs()
        ro.cap_votes = cap_votes
        for k in o:
            setattr(ro, "d_" + k, o[k].data_ptr())
        eng.set_option("rr_filter_only", 1 if filter_only else 0)
        try:
            _lib.check(eng.ctx, lib().fuz_rr_track(eng.ctx, C.byref(ri), C.byref(ro)))
        finally:
            eng.set_option("rr_filter_only", 0)
        st = eng.status(raise_on_error=False)
        if st.error == _lib.FUZ_OK:
            n_votes = int(st.reserved[1])
            return (o["keep"][:n_ovl].cpu().numpy(), o["hp_n"].cpu().numpy(), o["hp_len"].cpu().numpy().reshape(n_reads, b),
                    o["hp_q"].cpu().numpy().reshape(n_reads, b), o["vt_off"].cpu().numpy(), o["vt_ctg"][:n_votes].cpu().numpy(),
                    o["vt_count"][:n_votes].cpu().numpy(), o["vt_score"][:n_votes].cpu().numpy())
        if st.error == _lib.FUZ_E_CAPACITY and st.error_index == 8:
            cap_votes = int(st.reserved[2] * 1.1) + 1024
            continue
        eng.status()          # raises with the library's message
    raise FuzError(_lib.FUZ_E_CAPACITY, "vote capacity retry did not converge")


def _first_kept_order(t: np.ndarray, keep: np.ndarray) -> np.ndarray:
    """target ids in order of their first kept overlap line (dict insertion order, :59)."""
    kt = t[keep.astype(bool)]
    if len(kt) == 0:
        return kt
    _, first = np.unique(kt, return_index=True)
    return kt[np.sort(first)]


def tr_stage1(readlines: Callable[[], Iterable[str]], min_len: int, bestn: int, rid_to_ctg, rid_to_phase):
    """reference rr_hctg_track.py:31-65: {t_id: heapq array of (overlap_len, q_id)} for one LAS
    file; targets in first-kept-appearance order."""
    lines = la4falcon.DeviceLines([readlines()], require_id9=False)
    tab = _Tables(rid_to_ctg, rid_to_phase, len(rid_to_phase))
    keep, hp_n, hp_len, hp_q, *_ = _track_device(None, None, None, None, None, tab, min_len, bestn, lines=lines)
    keep = keep[:lines.n]
    t = lines.a["t"]
    rtn = {}
    for tid in _first_kept_order(t, keep).tolist():
        n = int(hp_n[tid])
        rtn["%09d" % tid] = [(int(hp_len[tid, j]), "%09d" % int(hp_q[tid, j])) for j in range(n)]
    return rtn


def run_tr_stage1(db_fn, fn, min_len, bestn, rid_to_ctg, rid_to_phase):
    """reference rr_hctg_track.py:25-29: (LAS file name, its heaps)."""
    return fn, tr_stage1(lambda: read_las_lines(db_fn, fn), min_len, bestn, rid_to_ctg, rid_to_phase)


def _load_tables(phased_read_file_fn, read_to_contig_map_fn, rawread_ids_fn):
    """the id tables of reference rr_hctg_track.py:70-85."""
    rid_to_ctg = get_rid_to_ctg(read_to_contig_map_fn)
    oid_to_phase = {}
    with open(phased_read_file_fn) as f:
        for row in f:
            row = row.strip().split()
            if not row:
                continue
            ctg_id, block, phase = row[1:4]
            oid_to_phase[row[6]] = (ctg_id, int(block), int(phase))
    with open(rawread_ids_fn) as f:
        rid_to_oid = f.read().split("\n")
    rid_to_phase = [oid_to_phase.get(oid) for oid in rid_to_oid]
    return rid_to_ctg, _Tables(rid_to_ctg, rid_to_phase, len(rid_to_phase))


def _parse_files(db_fn, files: Sequence[str], file_ids: Sequence[int]):
    """LA4Falcon lines of `files` -> (q, t, len, tlen, file_idx) in (file, line) order."""
    parts = [_parse_lines(read_las_lines(db_fn, fn)) for fn in files]
    q, t, ln, tl = (np.concatenate([p[k] for p in parts]) if parts else np.zeros(0, np.int32) for k in range(4))
    file_idx = (np.concatenate([np.full(len(p[0]), i, np.int32) for i, p in zip(file_ids, parts)])
                if parts else np.zeros(0, np.int32))
    return q, t, ln, tl, file_idx


def _bread_order(t_kept: np.ndarray, file_kept: np.ndarray) -> List[str]:
    """b-reads in the iteration order of the reference's `bread_to_areads` dict: keys are
    inserted per file (ascending), the file's targets in the CPython-2 order of ITS dict (keys
    inserted at their first kept line), first sight wins (:97-100); iteration = CPython-2 order
    of the merged dict (:113).  Inputs: target and file index of every KEPT line, (file, line) order."""
    inserted: Dict[str, None] = {}
    if len(t_kept):
        cut = np.flatnonzero(np.diff(file_kept)) + 1
        for seg in np.split(t_kept, cut):
            _, first = np.unique(seg, return_index=True)
            keys = ["%09d" % x for x in seg[np.sort(first)].tolist()]
            for k in py2compat.str_dict_order(keys):
                inserted.setdefault(k, None)
    return py2compat.str_dict_order(inserted)


def _format_bread(bread: str, tab: _Tables, rid_to_ctg, vt_off, vt_ctg, vt_count, vt_score) -> str:
    """rows of one b-read (reference rr_hctg_track.py:126-138)."""
    tid = int(bread)
    lo, hi = int(vt_off[tid]), int(vt_off[tid + 1])
    if lo == hi:
        return ""
    names = tab.ctg_names
    ctgs = [names[c] for c in vt_ctg[lo:hi].tolist()]              # dict insertion order
    score = dict(zip(ctgs, zip(vt_score[lo:hi].tolist(), vt_count[lo:hi].tolist())))
    items = [(k, score[k]) for k in py2compat.str_dict_order(ctgs)]   # ctg_score.items() (:126)
    items.sort(key=lambda kv: kv[1][0])                            # stable sort by score (:127)
    own = rid_to_ctg.get(bread)
    out = []
    for rank, (ctg, (sc, cnt)) in enumerate(items):
        in_ctg = 1 if own is not None and ctg in own else 0
        out.append("%s %s %d %d %d %d\n" % (bread, ctg, cnt, rank, sc, in_ctg))
    return "".join(out)


def _bread_order_native(t_kept: np.ndarray, file_kept: np.ndarray) -> np.ndarray:
    """_bread_order through libfuz (fuz_host_rr_bread_order): int ids instead of "%09d" strings."""
    t_kept = np.ascontiguousarray(t_kept, dtype=np.int32)
    file_kept = np.ascontiguousarray(file_kept, dtype=np.int32)
    out = np.empty(max(len(t_kept), 1), np.int32)
    n = lib().fuz_host_rr_bread_order(t_kept.ctypes.data, file_kept.ctypes.data, len(t_kept), out.ctypes.data)
    if n < 0:
        raise FuzError(_lib.FUZ_E_ARG, "fuz_host_rr_bread_order failed")
    return out[:n]


def _format_rows_native(breads: np.ndarray, tab: _Tables, vt_off, vt_ctg, vt_count, vt_score) -> bytes:
    """All rows of rawread_to_contigs through libfuz (fuz_host_rr_format_rows)."""
    names = [c.encode("ascii") for c in tab.ctg_names]
    ctg_off = np.concatenate([[0], np.cumsum([len(c) for c in names])]).astype(np.int64)
    blob = b"".join(names) + b"\0"
    arrs = [np.ascontiguousarray(a, dtype=dt) for a, dt in ((breads, np.int32), (vt_off, np.int32), (vt_ctg, np.int32),
                                                           (vt_count, np.int32), (vt_score, np.int64))]
    if len(arrs[2]) == 0:
        arrs[2], arrs[3], arrs[4] = np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int64)
    args = (arrs[0].ctypes.data, len(arrs[0]), arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data, arrs[4].ctypes.data,
            blob, ctg_off.ctypes.data, tab.in_map.ctypes.data, tab.rc_off.ctypes.data, tab.rc_ctg.ctypes.data)
    size = lib().fuz_host_rr_format_rows(*args, None, 0)
    if size < 0:
        raise FuzError(_lib.FUZ_E_ARG, "fuz_host_rr_format_rows failed")
    buf = C.create_string_buffer(int(size) + 1)
    lib().fuz_host_rr_format_rows(*args, buf, size)
    return buf.raw[:size]


def _write_rows(fn: str, text) -> None:
    os.makedirs(os.path.dirname(fn) or ".", exist_ok=True)
    with open(fn + ".tmp", "wb" if isinstance(text, bytes) else "w") as f:
        f.write(text)
    os.replace(fn + ".tmp", fn)


def run_track_reads(exe_pool, phased_read_file_fn, read_to_contig_map_fn, rawread_ids_fn, file_list, min_len, bestn, db_fn,
                    rawread_to_contigs_fn):
    """reference rr_hctg_track.py:67-138.  `exe_pool` is accepted for signature compatibility:
    every LAS file goes through ONE device call (the per-file heaps and their merge, :97-105,
    are replayed inside the kernel)."""
    rid_to_ctg, tab = _load_tables(phased_read_file_fn, read_to_contig_map_fn, rawread_ids_fn)
    files = list(file_list)          # the caller's order decides heap arrays and row order (reference :88-105)
    # the LA4Falcon text of all files is parsed on the device; the columns stay there for fuz_rr_track
    lines = la4falcon.DeviceLines([read_las_lines(db_fn, fn) for fn in files], require_id9=False)
    keep, _hn, _hl, _hq, vt_off, vt_ctg, vt_count, vt_score = _track_device(None, None, None, None, None, tab, min_len, bestn,
                                                                             lines=lines)
    kept = lines.gather(("t", "file"), np.flatnonzero(keep[:lines.n]))
    # row order (CPython-2 dict orders) and text: C++ twins of _bread_order / _format_bread
    breads = _bread_order_native(kept["t"], kept["file"])
    _write_rows(rawread_to_contigs_fn, _format_rows_native(breads, tab, vt_off, vt_ctg, vt_count, vt_score))


def run_track_reads_sharded(phased_read_file_fn, read_to_contig_map_fn, rawread_ids_fn, file_list, min_len, bestn, db_fn,
                            rawread_to_contigs_fn, rank: int, world_size: int, group=None):
    """run_track_reads over `world_size` processes (one per GPU; torch.distributed initialised
    by the caller).  It mirrors the reference's map + merge (rr_hctg_track.py:88-105):

      map       LAS files are dealt round-robin to the ranks; a rank parses its files and
                runs the overlap filter (:45-57) on its GPU;
      exchange  ONE all-gather of the kept overlap lines (q, t, len, tlen, file: int32 x 5);
      merge     every rank replays the per-file heaps and their merge and takes the contig vote
                for the targets t with t % world_size == rank (a target's result depends only on
                its own kept lines in (file, line) order, so the shard is exact);
      rows      formatted per rank, collected by rank 0, written in the reference's row order.

    The id tables are replicated (every rank reads the three table files)."""
    import torch
    import torch.distributed as dist
    rid_to_ctg, tab = _load_tables(phased_read_file_fn, read_to_contig_map_fn, rawread_ids_fn)
    files = list(file_list)
    mine = list(range(rank, len(files), world_size))
    backend = dist.get_backend(group)
    if backend == "nccl":                                  # text parsed on the device, only the kept lines come back
        lines = la4falcon.DeviceLines([read_las_lines(db_fn, files[i]) for i in mine], require_id9=False)
        keep = _track_device(None, None, None, None, None, tab, min_len, bestn, filter_only=True, lines=lines)[0][:lines.n].astype(bool)
        g = lines.gather(("q", "t", "len", "tl", "file"), np.flatnonzero(keep))
        file_of = np.asarray(mine, np.int32)[g["file"]] if len(mine) else g["file"]
        local = np.stack([g["q"], g["t"], g["len"], g["tl"], file_of], axis=1).astype(np.int32)
        q = np.zeros(lines.n, np.int8)                     # only its length is reported below
    else:
        q, t, ln, tl, file_idx = _parse_files(db_fn, [files[i] for i in mine], mine)
        keep = _track_device(q, t, ln, tl, file_idx, tab, min_len, bestn, filter_only=True)[0].astype(bool)
        local = np.stack([q[keep], t[keep], ln[keep], tl[keep], file_idx[keep]], axis=1).astype(np.int32)
    dev = engine.get_engine().device if backend == "nccl" else torch.device("cpu")
    # exchange: sizes, then the padded line blocks
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world_size)]
    dist.all_gather(sizes, n_local, group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(max(sizes), 1)
    send = torch.zeros((pad, 5), dtype=torch.int32, device=dev)
    if len(local):
        send[:len(local)] = torch.from_numpy(local).to(dev)
    recv = [torch.zeros((pad, 5), dtype=torch.int32, device=dev) for _ in range(world_size)]
    dist.all_gather(recv, send, group=group)
    lines = np.concatenate([r[:n].cpu().numpy() for r, n in zip(recv, sizes)], axis=0)
    lines = lines[np.argsort(lines[:, 4], kind="stable")]          # (file, line) order: files are disjoint across ranks
    # merge + vote for this rank's targets
    own = lines[lines[:, 1] % world_size == rank]
    _k, _hn, _hl, _hq, vt_off, vt_ctg, vt_count, vt_score = _track_device(
        own[:, 0], own[:, 1], own[:, 2], own[:, 3], own[:, 4], tab, min_len, bestn)
    order = _bread_order(lines[:, 1], lines[:, 4])
    blocks = {b: _format_bread(b, tab, rid_to_ctg, vt_off, vt_ctg, vt_count, vt_score)
              for b in order if int(b) % world_size == rank}
    gathered = [None] * world_size
    dist.all_gather_object(gathered, blocks, group=group)
    if rank == 0:
        merged = {}
        for g in gathered:
            merged.update(g)
        _write_rows(rawread_to_contigs_fn, "".join(merged[b] for b in order))
    dist.barrier(group=group)
    return dict(lines_local=int(len(q)), kept_local=int(len(local)), kept_total=int(len(lines)), targets_local=len(blocks))


def try_run_track_reads(n_core, phased_read_file, read_to_contig_map, rawread_ids, min_len, bestn, output):
    """reference rr_hctg_track.py:142-160 (the process pool is gone: one GPU call)."""
    rawread_dir = os.path.abspath("0-rawreads")
    file_list = sorted(glob.glob(os.path.join(rawread_dir, "m*/raw_reads.*.las")))      # upstream: filesystem order (:148)
    db_fn = os.path.join(rawread_dir, "raw_reads.db")
    run_track_reads(None, phased_read_file, read_to_contig_map, rawread_ids, file_list, min_len, bestn, db_fn, output)


def track_reads(n_core, phased_read_file, read_to_contig_map, rawread_ids, min_len, bestn, debug, silent, stream, output):
    try_run_track_reads(n_core, phased_read_file, read_to_contig_map, rawread_ids, min_len, bestn, output)


def parse_args(argv):
    parser = argparse.ArgumentParser(
        description="scan the raw read overlap information to identify the best hit from the reads to the contigs with "
                    "read_to_contig_map generated by `fc_get_read_hctg_map`. Write rawread_ids.",
        formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--n-core", type=int, default=48,
                        help="accepted for compatibility (the work runs on the GPU)")
    parser.add_argument("--phased-read-file", type=str, default="./3-unzip/all_phased_reads", help="phased-read-file")
    parser.add_argument("--read-to-contig-map", type=str, default="./4-quiver/read_maps/read_to_contig_map",
                        help="read_to_contig_map, from fc_get_read_hctg_map")
    parser.add_argument("--rawread-ids", type=str, default="./2-asm-falcon/read_maps/dump_rawread_ids/rawread_ids",
                        help="rawread_ids file")
    parser.add_argument("--output", type=str, default="./2-asm-falcon/read_maps/dump_rawread_ids/rawread_to_contigs",
                        help="Output")
    parser.add_argument("--min-len", type=int, default=2500, help="min length of the reads")
    parser.add_argument("--stream", action="store_true", help="accepted for compatibility")
    parser.add_argument("--debug", "-g", action="store_true", help="accepted for compatibility")
    parser.add_argument("--silent", action="store_true", help="accepted for compatibility")
    parser.add_argument("--bestn", type=int, default=40, help="keep best n hits")
    return parser.parse_args(argv[1:])


def main(argv=sys.argv):
    args = parse_args(argv)
    track_reads(**vars(args))


if __name__ == "__main__":
    main()
