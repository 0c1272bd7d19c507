"""Host side of the device pipeline: batch preparation (record index, QNAME -> q_id),
buffer ownership (PyTorch tensors), calls into libfuz.so, capacity retry.

PyTorch is plumbing here: it owns device / pinned memory and the CUDA stream; every
computation of the phasing path happens in the hand-written kernels of libfuz.so.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import FuzError, lib


def _np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


# --------------------------------------------------------------------------- batches
@dataclasses.dataclass
class PreparedBatch:
    """Host arrays of one batch of contigs (see include/fuz.h, fuz_host_batch)."""
    records: np.ndarray          # uint8 [rec_bytes] verbatim BAM records
    rec_off: np.ndarray          # int64 [n_rec + 1]
    rec_qid: Optional[np.ndarray]  # int32 [n_rec]; None: assigned on the device by phase_host
    ctg_rec_off: np.ndarray      # int32 [n_ctg + 1]
    ctg_len: np.ndarray          # int32 [n_ctg]
    ctg_nq: np.ndarray           # int32 [n_ctg]
    name_first: np.ndarray       # int64 [sum nq] record index of the first record of each q_id
    ctg_names: List[str]
    pinned: Dict[str, object] = dataclasses.field(default_factory=dict)  # torch tensors keeping pinned memory alive

    @property
    def n_ctg(self) -> int:
        return len(self.ctg_len)

    @property
    def n_rec(self) -> int:
        return len(self.rec_off) - 1

    @property
    def ctg_q_off(self) -> np.ndarray:
        return np.concatenate([[0], np.cumsum(self.ctg_nq)]).astype(np.int64)

    def goff(self) -> np.ndarray:
        t = lib().fuz_tile_size()
        padded = np.maximum((self.ctg_len.astype(np.int64) + t - 1) // t * t, t)
        return np.concatenate([[0], np.cumsum(padded)]).astype(np.int64)

    def qname(self, c: int, q: int) -> str:
        r = int(self.name_first[int(self.ctg_q_off[c]) + q])
        o = int(self.rec_off[r])
        l_name = int(self.records[o + 12])
        return self.records[o + 36:o + 36 + l_name - 1].tobytes().decode("ascii")

    def qname_rows(self, c: int):
        """QNAMEs of contig c as they are kept: fixed-width NUL-padded rows (numpy "S"), else the list of str.  The file
        formatters of libfuz take the rows as they are."""
        return self._names[int(self._q_off[c]):int(self._q_off[c + 1])]

    def qnames(self, c: int) -> List[str]:
        return [self.qname(c, q) for q in range(int(self.ctg_nq[c]))]


def index_records(records: np.ndarray) -> np.ndarray:
    """Offsets of the records of a concatenated record buffer (block_size chain)."""
    records = np.ascontiguousarray(records, dtype=np.uint8)
    n = C.c_int64(0)
    rc = lib().fuz_host_index_records(_np_ptr(records), len(records), None, 0, C.byref(n))
    if rc:
        raise FuzError(rc, "corrupt BAM record stream")
    off = np.empty(n.value + 1, dtype=np.int64)
    rc = lib().fuz_host_index_records(_np_ptr(records), len(records), _np_ptr(off), n.value, C.byref(n))
    if rc:
        raise FuzError(rc, "corrupt BAM record stream")
    return off


def record_refids(records: np.ndarray, rec_off: np.ndarray) -> np.ndarray:
    idx = rec_off[:-1, None] + 4 + np.arange(4)[None, :]
    return records[idx].copy().view("<i4").reshape(-1)


def prepare_batch(records, ctg_names: Sequence[str], ctg_lens: Sequence[int],
                  rec_off: Optional[np.ndarray] = None, ctg_rec_off: Optional[np.ndarray] = None,
                  pin: bool = False, assign_qids: bool = True) -> PreparedBatch:
    """records: concatenated BAM records grouped by contig in the order of ctg_names (refID
    ascending, i.e. a coordinate-sorted BAM).  If ctg_rec_off is None the grouping is read
    from the refID fields, which must be 0..n_ctg-1 in order.  assign_qids=False leaves the
    QNAME -> q_id assignment (phasing.py:47-54) to the device (Engine.phase_host fills ctg_nq
    and name_first of the batch)."""
    if not isinstance(records, np.ndarray):
        records = np.frombuffer(records, dtype=np.uint8)
    records = np.ascontiguousarray(records, dtype=np.uint8)
    if rec_off is None:
        rec_off = index_records(records)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
    n_rec, n_ctg = len(rec_off) - 1, len(ctg_names)
    if ctg_rec_off is None:
        refid = record_refids(records, rec_off) if n_rec else np.zeros(0, np.int32)
        if n_rec and (np.any(np.diff(refid) < 0) or refid.min() < 0 or refid.max() >= n_ctg):
            raise FuzError(_lib.FUZ_E_UNSORTED, "records are not grouped by reference id 0..%d" % (n_ctg - 1))
        ctg_rec_off = np.searchsorted(refid, np.arange(n_ctg + 1), side="left")
    ctg_rec_off = np.ascontiguousarray(ctg_rec_off, dtype=np.int32)
    if not assign_qids:
        pb = PreparedBatch(records, rec_off, None, ctg_rec_off, np.asarray(ctg_lens, dtype=np.int32),
                           np.zeros(n_ctg, dtype=np.int32), np.zeros(0, np.int64), list(ctg_names))
        if pin:
            pin_batch(pb)
        return pb
    rec_qid = np.empty(n_rec, dtype=np.int32)
    ctg_nq = np.zeros(n_ctg, dtype=np.int32)
    name_first = np.empty(max(n_rec, 1), dtype=np.int64)
    rc = lib().fuz_host_assign_qids(_np_ptr(records), _np_ptr(rec_off), n_rec, _np_ptr(ctg_rec_off), n_ctg,
                                    _np_ptr(rec_qid), _np_ptr(ctg_nq), _np_ptr(name_first))
    if rc:
        raise FuzError(rc, "fuz_host_assign_qids failed")
    pb = PreparedBatch(records, rec_off, rec_qid, ctg_rec_off, np.asarray(ctg_lens, dtype=np.int32), ctg_nq,
                       name_first[:int(ctg_nq.sum())].copy(), list(ctg_names))
    if pin:
        pin_batch(pb)
    return pb


# --------------------------------------------------------------------------- batch splitting
# One device call takes at most 2^31 - 64 Ki global positions (tile-padded contig lengths; positions are int32 on the
# device) and 2^31 - 1 records; max_bytes bounds the record bytes of a batch (device staging of the host entry, scratch
# that scales with the records).  The reference has no such limit because it runs one process per contig
# (unzip.py:231-281); here an arbitrary contig list is cut into consecutive batches and the batches run back to back.
MAX_BATCH_GLEN = 0x7fff0000
MAX_BATCH_RECORDS = 0x7fffffff - 1


class BatchPacker:
    """Running totals of the batch being filled; fits() says whether one more contig stays inside the limits."""

    def __init__(self, max_bytes: int = 6 << 30):
        self.max_bytes, self.tile = int(max_bytes), lib().fuz_tile_size()
        self.reset()

    def reset(self) -> None:
        self.n = self.bytes = self.nrec = self.glen = 0

    def _pad(self, ctg_len: int) -> int:
        return max((int(ctg_len) + self.tile - 1) // self.tile * self.tile, self.tile)

    def fits(self, nbytes: int, nrec: int, ctg_len: int) -> bool:
        return self.n == 0 or not (self.bytes + nbytes > self.max_bytes or self.nrec + nrec > MAX_BATCH_RECORDS or
                                   self.glen + self._pad(ctg_len) > MAX_BATCH_GLEN)

    def add(self, nbytes: int, nrec: int, ctg_len: int) -> None:
        self.n += 1; self.bytes += int(nbytes); self.nrec += int(nrec); self.glen += self._pad(ctg_len)


def plan_batches(ctg_bytes: Sequence[int], ctg_nrec: Sequence[int], ctg_len: Sequence[int],
                 max_bytes: int = 6 << 30) -> List[Tuple[int, int]]:
    """Greedy cut of a contig list (kept in order) into batches [(first contig, one past the last)].  A contig never
    straddles two batches; a single contig beyond max_bytes gets a batch of its own (the device limits still apply and
    are reported by the library)."""
    pk = BatchPacker(max_bytes)
    out, lo = [], 0
    for c, (cb, cn, cl) in enumerate(zip(ctg_bytes, ctg_nrec, ctg_len)):
        if not pk.fits(cb, cn, cl):
            out.append((lo, c))
            lo = c
            pk.reset()
        pk.add(cb, cn, cl)
    if lo < len(ctg_len):
        out.append((lo, len(ctg_len)))
    return out


def build_batch(parts: Sequence[Tuple[np.ndarray, np.ndarray]], ctg_names: Sequence[str], ctg_lens: Sequence[int],
                pin: bool = False, assign_qids: bool = False) -> PreparedBatch:
    """One batch from per-contig (records, rec_off) pieces: the records are copied ONCE, straight into the (optionally
    page-locked) buffer of the batch.  The contig of a record comes from ctg_rec_off (the refID fields are not read)."""
    total = int(sum(len(r) for r, _o in parts))
    n_rec = int(sum(len(o) - 1 for _r, o in parts))
    if pin:
        import torch
        t_rec = torch.empty(max(total, 1), dtype=torch.uint8).pin_memory()
        records = t_rec.numpy()[:total]
    else:
        t_rec, records = None, np.empty(total, np.uint8)
    rec_off = np.empty(n_rec + 1, np.int64)
    ctg_rec_off = np.zeros(len(parts) + 1, np.int32)
    b = r = 0
    for c, (rec, off) in enumerate(parts):
        records[b:b + len(rec)] = rec
        rec_off[r:r + len(off) - 1] = off[:-1] + b
        b += len(rec); r += len(off) - 1
        ctg_rec_off[c + 1] = r
    rec_off[n_rec] = b
    pb = prepare_batch(records, ctg_names, ctg_lens, rec_off=rec_off, ctg_rec_off=ctg_rec_off, pin=False,
                       assign_qids=assign_qids)
    if pin:
        pb.pinned["records"] = t_rec
        pb.records = records
        for name in ("rec_off", "rec_qid", "ctg_rec_off", "ctg_len", "ctg_nq"):
            a = getattr(pb, name)
            if a is not None:
                import torch
                t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
                pb.pinned[name] = t
                setattr(pb, name, t.numpy())
    return pb


def sub_batch(pb_records: np.ndarray, rec_off: np.ndarray, ctg_rec_off: np.ndarray, lo: int, hi: int):
    """Records, offsets and contig ranges of contigs [lo, hi) of a larger grouped record buffer (views, no copy)."""
    r0, r1 = int(ctg_rec_off[lo]), int(ctg_rec_off[hi])
    b0, b1 = int(rec_off[r0]), int(rec_off[r1])
    return pb_records[b0:b1], rec_off[r0:r1 + 1] - b0, (np.asarray(ctg_rec_off[lo:hi + 1]) - r0).astype(np.int32)


def pin_batch(pb: PreparedBatch) -> None:
    """Move the batch's arrays into pinned host memory (torch owns it)."""
    import torch
    for name in ("records", "rec_off", "rec_qid", "ctg_rec_off", "ctg_len", "ctg_nq"):
        a = getattr(pb, name)
        if a is None:
            continue
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        pb.pinned[name] = t
        setattr(pb, name, t.numpy())


# --------------------------------------------------------------------------- results
@dataclasses.dataclass
class PhaseResult:
    arrays: Dict[str, np.ndarray]
    n_sites: int
    n_vmap: int
    n_atable: int
    n_reads: int
    n_accepted: int
    aligned_bases: int
    h2d_bytes: int = 0
    d2h_bytes: int = 0

    def __getattr__(self, name):
        try:
            return self.__dict__["arrays"][name]
        except KeyError:
            raise AttributeError(name)


_CAP_KEY_N = {"sites": "n_sites", "vmap": "n_vmap", "atable": "n_atable", "reads": "n_reads"}


def default_caps(total_len: int, n_rec: int) -> Dict[str, int]:
    sites = max(4096, total_len // 128)
    return dict(sites=sites, vmap=max(1 << 16, sites * 64), atable=max(1 << 16, sites * 24),
                reads=max(1 << 12, 2 * n_rec))


def _grow(caps: Dict[str, int], st: _lib.Status) -> Dict[str, int]:
    new = dict(caps)
    for key, need in (("sites", st.need_sites), ("vmap", st.need_vmap), ("atable", st.need_atable),
                      ("reads", st.need_reads)):
        if need > caps[key]:
            new[key] = int(need * 1.25) + 1024
    return new


class DeviceOutputs:
    """fuz_outputs backed by torch tensors on the context's device."""

    def __init__(self, caps: Dict[str, int], device, counts_len: int = 0):
        import torch
        self.caps = dict(caps)
        self.t: Dict[str, "torch.Tensor"] = {}
        self.c = _lib.Outputs()
        for key in ("sites", "vmap", "atable", "reads"):
            setattr(self.c, "cap_" + key, caps[key])
        for name, dt, key, width in _lib.OUTPUT_ARRAYS:
            t = torch.zeros(max(caps[key] * width, 4), dtype=torch.int32 if dt == "i4" else torch.uint8,
                            device=device)
            self.t[name] = t
            setattr(self.c, "d_" + name, t.data_ptr())
        self.counts = None
        if counts_len:
            self.counts = torch.zeros(counts_len * 4, dtype=torch.int32, device=device)
            self.c.d_counts = self.counts.data_ptr()

    def set(self, name: str, values: np.ndarray) -> None:
        import torch
        v = torch.from_numpy(np.ascontiguousarray(values).reshape(-1))
        self.t[name][:v.numel()].copy_(v.view(self.t[name].dtype) if v.dtype != self.t[name].dtype else v)

    def fetch(self, st: _lib.Status, keys: Sequence[str] = ("sites", "vmap", "atable", "reads")) -> Dict[str, np.ndarray]:
        """Download the filled prefix of the row arrays of the given groups."""
        out = {}
        for name, dt, key, width in _lib.OUTPUT_ARRAYS:
            if key not in keys:
                continue
            n = min(int(getattr(st, _CAP_KEY_N[key])), self.caps[key])
            a = self.t[name][:n * width].cpu().numpy()
            out[name] = a.reshape(n, width) if width > 1 else a
        return out


class DeviceBatch:
    """fuz_batch backed by torch tensors."""

    def __init__(self, pb: PreparedBatch, device):
        import torch
        self.pb = pb
        goff = pb.goff()

        def up(a, pad=0):
            t = torch.from_numpy(np.ascontiguousarray(a))
            if pad:
                d = torch.zeros(t.numel() + pad, dtype=t.dtype, device=device)
                d[:t.numel()].copy_(t, non_blocking=True)
                return d
            return t.to(device, non_blocking=True)
        self.rec_buf = up(pb.records, pad=64)
        self.rec_off = up(pb.rec_off)
        self.dev_qid = pb.rec_qid is None                # q_ids assigned by fuz_phase_batch
        self.rec_qid = up(pb.rec_qid) if pb.n_rec and not self.dev_qid else torch.zeros(1, dtype=torch.int32, device=device)
        self.qid_nq = torch.zeros(max(pb.n_ctg, 1), dtype=torch.int32, device=device)
        self.qid_first = torch.zeros(max(pb.n_rec, 1), dtype=torch.int64, device=device)
        self.ctg_rec_off = up(pb.ctg_rec_off)
        self.ctg_len = up(pb.ctg_len)
        self.ctg_goff = up(goff)
        self.ctg_nq = up(pb.ctg_nq)
        self.total_glen = int(goff[-1])
        b = _lib.Batch()
        b.n_ctg, b.n_rec, b.rec_bytes = pb.n_ctg, pb.n_rec, len(pb.records)
        b.d_rec_buf, b.d_rec_off = self.rec_buf.data_ptr(), self.rec_off.data_ptr()
        b.d_rec_qid = None if self.dev_qid else self.rec_qid.data_ptr()
        b.d_ctg_rec_off, b.d_ctg_len = self.ctg_rec_off.data_ptr(), self.ctg_len.data_ptr()
        b.d_ctg_goff, b.d_ctg_nq = self.ctg_goff.data_ptr(), self.ctg_nq.data_ptr()
        b.total_glen, b.total_nq = self.total_glen, int(pb.ctg_nq.sum())
        self.c = b


@dataclasses.dataclass
class DeviceBam:
    """A BAM file decoded on the device (Engine.ingest_bam): the inflated stream, the record
    index and the record range of every reference, all in HBM."""
    refs: List                   # [(name, length)] from the BAM header
    raw: object                  # torch uint8: padding + inflated stream + slack
    rec_ptr: int                 # device address of the first alignment record (256-byte aligned)
    rec_bytes: int
    rec_off: object              # torch int64 [>= n_rec + 1]
    ctg_rec_off: object          # torch int32 [n_ref + 1]
    n_rec: int                   # all records; the mapped ones are the first n_mapped
    n_mapped: int
    h2d_bytes: int
    keep: tuple = ()             # tensors that must outlive the asynchronous copies

    def records(self) -> np.ndarray:
        """Inflated alignment records on the host (tests / diagnostics)."""
        a = self.raw.cpu().numpy()
        o = self.rec_ptr - self.raw.data_ptr()
        return a[o:o + self.rec_bytes]


class BamBatchInfo:
    """What the file writers need from a batch that only exists on the device: contig names
    and the QNAME of every q_id (gathered on the device, a few bytes per read; a list of str or a numpy "S" array)."""

    def __init__(self, ctg_names, ctg_lens, ctg_nq, names):
        self.ctg_names, self.ctg_len = list(ctg_names), np.asarray(ctg_lens, np.int32)
        self.ctg_nq = np.asarray(ctg_nq, np.int32)
        self._names = names
        self._q_off = np.concatenate([[0], np.cumsum(self.ctg_nq)]).astype(np.int64)

    @property
    def n_ctg(self) -> int:
        return len(self.ctg_names)

    def qname_rows(self, c: int):
        """QNAMEs of contig c as they are kept: fixed-width NUL-padded rows (numpy "S"), else the list of str.  The file
        formatters of libfuz take the rows as they are."""
        return self._names[int(self._q_off[c]):int(self._q_off[c + 1])]

    def qnames(self, c: int) -> List[str]:
        """QNAME of every q_id of contig c.  The batch keeps the names as fixed-width byte rows (numpy "S"); the str
        objects are made here, for the contig that is being written."""
        part = self._names[int(self._q_off[c]):int(self._q_off[c + 1])]
        if isinstance(part, np.ndarray):
            return [b.decode("latin-1") for b in part.tolist()]
        return part


class Engine:
    """One libfuz context = one GPU = one stream (not thread safe; include/fuz.h)."""

    def __init__(self, device: int = 0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("falcon_unzip_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device_index = device
        self.device = torch.device("cuda", device)
        self.ctx = C.c_void_p()
        rc = lib().fuz_ctx_create(device, C.byref(self.ctx))
        if rc:
            raise FuzError(rc, lib().fuz_last_error(None).decode())
        self._torch = torch

    def close(self) -> None:
        if self.ctx:
            lib().fuz_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- options / status
    def set_option(self, key: str, value: int) -> None:
        _lib.check(self.ctx, lib().fuz_set_option(self.ctx, key.encode(), value))

    def status(self, raise_on_error: bool = True) -> _lib.Status:
        st = _lib.Status()
        rc = lib().fuz_get_status(self.ctx, C.byref(st))
        if rc and raise_on_error:
            _lib.check(self.ctx, rc)
        return st

    def sync(self) -> None:
        _lib.check(self.ctx, lib().fuz_sync(self.ctx))

    def launch_count(self) -> int:
        return int(lib().fuz_launch_count(self.ctx))

    def kernel_timing(self, enable: bool) -> None:
        _lib.check(self.ctx, lib().fuz_kernel_timing(self.ctx, int(enable)))

    def get_kernel_timing(self):
        ms, n = C.c_double(0), C.c_int64(0)
        _lib.check(self.ctx, lib().fuz_get_kernel_timing(self.ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile(self, enable: bool) -> None:
        _lib.check(self.ctx, lib().fuz_profile(self.ctx, int(enable)))

    def profile_report(self):
        """[(kernel name, ms)] for every launch since profile(True)."""
        buf = C.create_string_buffer(1 << 20)
        lib().fuz_profile_report(self.ctx, buf, len(buf))
        out = []
        for line in buf.value.decode().splitlines():
            name, ms = line.split("\t")
            out.append((name, float(ms)))
        return out

    # ---- device-resident path
    def upload(self, pb: PreparedBatch) -> DeviceBatch:
        db = DeviceBatch(pb, self.device)
        self._torch.cuda.synchronize(self.device)
        return db

    def alloc_outputs(self, caps: Dict[str, int], counts_len: int = 0) -> DeviceOutputs:
        do = DeviceOutputs(caps, self.device, counts_len)
        self._torch.cuda.synchronize(self.device)
        return do

    def phase_batch_async(self, db: DeviceBatch, do: DeviceOutputs) -> None:
        if db.dev_qid:                                   # library assigns the q_ids: tables come back here
            do.c.d_ctg_nq, do.c.d_name_first = db.qid_nq.data_ptr(), db.qid_first.data_ptr()
        _lib.check(self.ctx, lib().fuz_phase_batch(self.ctx, C.byref(db.c), C.byref(do.c)))

    def assign_qids(self, db: DeviceBatch):
        """fuz_assign_qids on an uploaded batch -> (rec_qid, ctg_nq, name_first) as numpy arrays."""
        torch = self._torch
        pb = db.pb
        qid = torch.zeros(max(pb.n_rec, 1), dtype=torch.int32, device=self.device)
        _lib.check(self.ctx, lib().fuz_assign_qids(self.ctx, db.rec_buf.data_ptr(), db.rec_off.data_ptr(), pb.n_rec, len(pb.records),
                                                   db.ctg_rec_off.data_ptr(), pb.n_ctg, qid.data_ptr(), db.qid_nq.data_ptr(),
                                                   db.qid_first.data_ptr()))
        self.sync()
        nq = db.qid_nq.cpu().numpy()[:pb.n_ctg]
        return qid.cpu().numpy()[:pb.n_rec], nq, db.qid_first.cpu().numpy()[:int(nq.sum())]

    def het_call_async(self, db: DeviceBatch, do: DeviceOutputs) -> None:
        _lib.check(self.ctx, lib().fuz_het_call(self.ctx, C.byref(db.c), C.byref(do.c)))

    def association_async(self, n_ctg: int, n_sites: int, n_vmap: int, do: DeviceOutputs) -> None:
        _lib.check(self.ctx, lib().fuz_association_table(self.ctx, n_ctg, n_sites, n_vmap, C.byref(do.c)))

    def blocks_async(self, n_ctg: int, n_sites: int, n_atable: int, do: DeviceOutputs) -> None:
        _lib.check(self.ctx, lib().fuz_phased_blocks(self.ctx, n_ctg, n_sites, n_atable, C.byref(do.c)))

    def reads_async(self, n_ctg: int, ctg_nq: np.ndarray, n_sites: int, n_vmap: int, do: DeviceOutputs) -> None:
        t = self._torch.from_numpy(np.ascontiguousarray(ctg_nq, dtype=np.int32)).to(self.device)
        self._torch.cuda.synchronize(self.device)
        _lib.check(self.ctx, lib().fuz_phased_reads(self.ctx, n_ctg, t.data_ptr(), int(np.sum(ctg_nq)), n_sites,
                                                    n_vmap, C.byref(do.c)))
        self.sync()

    def _retry(self, caps, counts_len, run):
        """run(outputs) launches; on FUZ_E_CAPACITY grow the failing capacity and rerun."""
        for _ in range(8):
            do = self.alloc_outputs(caps, counts_len)
            run(do)
            st = self.status(raise_on_error=False)
            if st.error == _lib.FUZ_OK:
                return do, st
            if st.error != _lib.FUZ_E_CAPACITY:
                self.status()      # raises with the library's message
            if st.need_pairs > 0 and st.error_index == 4:
                per_site = st.need_pairs // max(caps["sites"], 1) + 2
                self.set_option("max_pairs_per_site", int(per_site))
            self._grow_scratch(st)
            caps = _grow(caps, st)
        raise FuzError(_lib.FUZ_E_CAPACITY, "capacity retry did not converge")

    def _grow_scratch(self, st) -> None:
        """FUZ_E_CAPACITY index 6 / 9: the batch needs more segment slots / tile entries than the heuristics of
        fuz_het_call reserve (CIGARs far denser than 1.5 bytes of SEQ + QUAL per base): reserve what the status reports."""
        if st.error_index == 6:
            self.set_option("seg_cap", int(st.n_segments * 1.1) + 1024)
        elif st.error_index == 9:
            self.set_option("ent_cap", int(st.reserved[0] * 1.5) + 1024)

    def phase_device(self, pb: PreparedBatch, caps: Optional[Dict[str, int]] = None,
                     want_counts: bool = False, stage: str = "all") -> PhaseResult:
        """Upload, run (all four stages or only "het"), download everything."""
        db = self.upload(pb)
        caps = caps or default_caps(int(pb.ctg_len.sum()), pb.n_rec)
        run = (lambda do: self.phase_batch_async(db, do)) if stage == "all" else (lambda do: self.het_call_async(db, do))
        do, st = self._retry(caps, db.total_glen if want_counts else 0, run)
        arrays = do.fetch(st)
        if db.dev_qid and stage == "all":
            pb.ctg_nq = db.qid_nq.cpu().numpy()[:pb.n_ctg]
            pb.name_first = db.qid_first.cpu().numpy()[:int(pb.ctg_nq.sum())].copy()
        if want_counts:
            arrays["counts"] = do.counts.cpu().numpy().view(np.uint32).reshape(-1, 4)
            arrays["goff"] = pb.goff()
        return PhaseResult(arrays, int(st.n_sites), int(st.n_vmap), int(st.n_atable), int(st.n_reads),
                           int(st.n_accepted), int(st.aligned_bases))

    # ---- BAM ingest on the device (BGZF inflate + record index)
    def read_file_pinned(self, path: str) -> np.ndarray:
        """The bytes of a file in a page-locked buffer that the engine keeps and reuses (grown when a larger file comes): no
        fresh pages per call, and the upload that follows is a DMA from pinned memory.  The view is valid until the next call."""
        torch = self._torch
        n = os.path.getsize(path)
        buf = getattr(self, "_file_pin", None)
        if buf is None or buf.numel() < n:
            self._file_pin = buf = torch.empty(max(n + (n >> 3), 1 << 20), dtype=torch.uint8, pin_memory=True)
        view = buf.numpy()[:n]
        with open(path, "rb", buffering=0) as f:
            got = 0
            mv = memoryview(view)
            while got < n:
                k = f.readinto(mv[got:])
                if not k:
                    raise IOError("%s: short read (%d of %d bytes)" % (path, got, n))
                got += k
        return view

    def ingest_bam(self, image, verify_crc: bool = True, profile: bool = False) -> DeviceBam:
        """image: the bytes of a coordinate-sorted BAM file (numpy uint8, ideally pinned).  The
        compressed image crosses PCIe; inflate, record index and grouping by reference run on
        the device.  Only the header blocks are inflated on the host (names and lengths)."""
        torch = self._torch
        from . import bam
        image = np.ascontiguousarray(image, dtype=np.uint8)
        n_blk = int(lib().fuz_host_bgzf_index(_np_ptr(image), len(image), 0, None, None, None, None))
        if n_blk < 0:
            raise FuzError(_lib.FUZ_E_FORMAT, "not a BGZF file")
        coff, csize = np.empty(n_blk, np.int64), np.empty(n_blk, np.int32)
        uoff, crc = np.empty(n_blk + 1, np.int64), np.empty(n_blk, np.uint32)
        lib().fuz_host_bgzf_index(_np_ptr(image), len(image), n_blk, _np_ptr(coff), _np_ptr(csize), _np_ptr(uoff), _np_ptr(crc))
        _text, refs, hdr_bytes = bam.read_bam_header(image, coff, csize)
        total = int(uoff[-1])
        pad = (-hdr_bytes) % 256
        dev = self.device
        d_comp = torch.empty((len(image) + 3) // 4 * 4 + 16, dtype=torch.uint8, device=dev)
        d_comp[:len(image)].copy_(torch.from_numpy(image), non_blocking=True)
        d_coff, d_csize = torch.from_numpy(coff).to(dev, non_blocking=True), torch.from_numpy(csize).to(dev, non_blocking=True)
        d_uoff = torch.from_numpy(uoff).to(dev, non_blocking=True)
        d_crc = torch.from_numpy(crc.view(np.int32)).to(dev, non_blocking=True) if verify_crc else None
        raw = torch.empty(pad + total + 64, dtype=torch.uint8, device=dev)
        raw[pad + total:].zero_()
        torch.cuda.synchronize(dev)                      # torch's stream -> the context's stream
        if profile:                                      # per-kernel times without the uploads above (bench legs)
            self.profile(True)
        _lib.check(self.ctx, lib().fuz_bgzf_inflate(self.ctx, d_comp.data_ptr(), len(image), d_coff.data_ptr(), d_csize.data_ptr(),
                                                    d_uoff.data_ptr(), d_crc.data_ptr() if verify_crc else None, n_blk,
                                                    raw.data_ptr() + pad, total))
        rec_ptr, rec_bytes = raw.data_ptr() + pad + hdr_bytes, total - hdr_bytes
        ctg_rec_off = torch.zeros(len(refs) + 1, dtype=torch.int32, device=dev)
        cap = rec_bytes // 2048 + 1024
        n_rec, need = C.c_int64(0), C.c_int64(0)
        for _ in range(2):
            rec_off = torch.empty(cap + 1, dtype=torch.int64, device=dev)
            torch.cuda.synchronize(dev)
            rc = lib().fuz_bam_index_records(self.ctx, rec_ptr, rec_bytes, len(refs), cap, rec_off.data_ptr(),
                                             ctg_rec_off.data_ptr(), C.byref(n_rec), C.byref(need))
            if rc != _lib.FUZ_E_CAPACITY:
                break
            cap = int(need.value)
        _lib.check(self.ctx, rc)
        n_mapped = int(ctg_rec_off[-1].item()) if len(refs) else 0
        h2d = len(image) + coff.nbytes + csize.nbytes + uoff.nbytes + (crc.nbytes if verify_crc else 0)
        return DeviceBam(refs, raw, rec_ptr, rec_bytes, rec_off, ctg_rec_off, int(n_rec.value), n_mapped, h2d,
                         keep=(d_comp, d_coff, d_csize, d_uoff, d_crc))

    def ingest_bam_windows(self, path: str, window_bytes: int = 512 << 20, verify_crc: bool = True):
        """A BAM file of any size as a sequence of DeviceBam WINDOWS: the file is memory-mapped, its BGZF blocks are taken
        in runs of about `window_bytes` compressed bytes, every run is inflated on the device behind the bytes of the record
        the previous window cut (fuz_bam_index_window reports where the last whole record ends), so host memory holds one
        window of compressed bytes and HBM one window of records at a time.  The reference streams such files record by
        record through pysam (select_reads_from_bam.py:55-76); raw-read BAMs run to hundreds of GB.  Yields DeviceBam
        objects whose records are whole; refs come from the header."""
        torch = self._torch
        from . import bam
        dev = self.device
        image = np.memmap(path, dtype=np.uint8, mode="r")
        n_blk = int(lib().fuz_host_bgzf_index(_np_ptr(image), len(image), 0, None, None, None, None))
        if n_blk < 0:
            raise FuzError(_lib.FUZ_E_FORMAT, "not a BGZF file")
        coff, csize = np.empty(n_blk, np.int64), np.empty(n_blk, np.int32)
        uoff, crc = np.empty(n_blk + 1, np.int64), np.empty(n_blk, np.uint32)
        lib().fuz_host_bgzf_index(_np_ptr(image), len(image), n_blk, _np_ptr(coff), _np_ptr(csize), _np_ptr(uoff), _np_ptr(crc))
        _text, refs, hdr_bytes = bam.read_bam_header(image, coff, csize)
        carry = torch.zeros(0, dtype=torch.uint8, device=dev)
        b0, skip = 0, hdr_bytes                              # bytes of the first window that belong to the header
        while b0 < n_blk:
            b1 = b0 + 1
            while b1 < n_blk and coff[b1] + csize[b1] - coff[b0] <= window_bytes:
                b1 += 1
            c_lo, c_hi = int(coff[b0]), int(coff[b1 - 1] + csize[b1 - 1])
            u_lo, u_hi = int(uoff[b0]), int(uoff[b1])
            n_c, n_u, n_carry = c_hi - c_lo, u_hi - u_lo, int(carry.numel())
            pad = (-n_carry) % 4                             # the inflate kernel writes from a 4-byte aligned start
            d_comp = torch.empty((n_c + 3) // 4 * 4 + 16, dtype=torch.uint8, device=dev)
            d_comp[:n_c].copy_(torch.from_numpy(np.array(image[c_lo:c_hi])))         # (a writable copy of the mapped bytes)
            d_coff = torch.from_numpy(coff[b0:b1] - c_lo).to(dev)
            d_csize = torch.from_numpy(csize[b0:b1].copy()).to(dev)
            d_uoff = torch.from_numpy(np.concatenate([uoff[b0:b1] - u_lo, [n_u]]).astype(np.int64)).to(dev)
            d_crc = torch.from_numpy(crc[b0:b1].view(np.int32).copy()).to(dev) if verify_crc else None
            raw = torch.empty(256 + pad + n_carry + n_u + 64, dtype=torch.uint8, device=dev)
            base = 256 + pad                                 # records start at raw[base]: carry, then the inflated window
            raw[base:base + n_carry].copy_(carry)
            raw[base + n_carry + n_u:].zero_()
            torch.cuda.synchronize(dev)
            if (raw.data_ptr() + base + n_carry) % 4:
                raise FuzError(_lib.FUZ_E_ARG, "window start is not 4-byte aligned")
            _lib.check(self.ctx, lib().fuz_bgzf_inflate(self.ctx, d_comp.data_ptr(), n_c, d_coff.data_ptr(), d_csize.data_ptr(),
                                                        d_uoff.data_ptr(), d_crc.data_ptr() if verify_crc else None, b1 - b0,
                                                        raw.data_ptr() + base + n_carry, n_u))
            start = base + (skip if n_carry == 0 else 0)
            if n_carry and skip:
                raise FuzError(_lib.FUZ_E_FORMAT, "BAM header longer than a window")
            if skip > n_u and n_carry == 0:                  # header not finished inside this window
                skip -= n_u
                b0 = b1
                continue
            rec_ptr, rec_bytes = raw.data_ptr() + start, base + n_carry + n_u - start
            skip = 0
            ctg_rec_off = torch.zeros(len(refs) + 1, dtype=torch.int32, device=dev)
            cap = rec_bytes // 2048 + 1024
            n_rec, need, tail = C.c_int64(0), C.c_int64(0), C.c_int64(0)
            for _ in range(2):
                rec_off = torch.empty(cap + 1, dtype=torch.int64, device=dev)
                torch.cuda.synchronize(dev)
                rc = lib().fuz_bam_index_window(self.ctx, rec_ptr, rec_bytes, len(refs), cap, rec_off.data_ptr(),
                                                ctg_rec_off.data_ptr(), C.byref(n_rec), C.byref(need), C.byref(tail))
                if rc != _lib.FUZ_E_CAPACITY:
                    break
                cap = int(need.value)
            _lib.check(self.ctx, rc)
            last = b1 == n_blk
            if last and tail.value != rec_bytes:
                raise FuzError(_lib.FUZ_E_BADRECORD, "the BAM file ends inside a record")
            o = start + int(tail.value)
            carry = raw[o:base + n_carry + n_u].clone()
            n_mapped = int(ctg_rec_off[-1].item()) if len(refs) else 0
            yield DeviceBam(refs, raw, rec_ptr, int(tail.value), rec_off, ctg_rec_off, int(n_rec.value), n_mapped, n_c,
                            keep=(d_comp, d_coff, d_csize, d_uoff, d_crc))
            b0 = b1

    def ingest_bams(self, images: Sequence, verify_crc: bool = True) -> DeviceBam:
        """Several BAM files (the reference makes one sorted BAM per contig, unzip.py:90) as ONE device batch.  The
        block tables of all files are joined: one upload buffer, ONE launch of the inflate kernel over the blocks of
        every file (a small file alone would leave the GPU idle behind a few serial block decodes), one record index
        over all files (fuz_bam_index_files), which also lays the mapped records of all files out back to back.  The
        reference lists are concatenated (names must be unique across the files)."""
        torch = self._torch
        from . import bam
        if len(images) == 0:
            raise FuzError(_lib.FUZ_E_ARG, "empty list of BAM files")
        if len(images) == 1:
            return self.ingest_bam(images[0], verify_crc)
        dev = self.device
        images = [np.ascontiguousarray(im, dtype=np.uint8) for im in images]
        tabs, refs_all, seg_start, seg_end, seg_nref = [], [], [], [], []
        cbase, ubase = 0, 0
        for k, im in enumerate(images):
            n_blk = int(lib().fuz_host_bgzf_index(_np_ptr(im), len(im), 0, None, None, None, None))
            if n_blk < 0:
                raise FuzError(_lib.FUZ_E_FORMAT, "file %d of the list is not a BGZF file" % k)
            coff, csize = np.empty(n_blk, np.int64), np.empty(n_blk, np.int32)
            uoff, crc = np.empty(n_blk + 1, np.int64), np.empty(n_blk, np.uint32)
            lib().fuz_host_bgzf_index(_np_ptr(im), len(im), n_blk, _np_ptr(coff), _np_ptr(csize), _np_ptr(uoff), _np_ptr(crc))
            _text, refs, hdr_bytes = bam.read_bam_header(im, coff, csize)
            total = int(uoff[-1])
            tabs.append((cbase, coff + cbase, csize, uoff[:-1] + ubase, crc))
            refs_all.extend(refs)
            seg_start.append(ubase + hdr_bytes)
            seg_end.append(ubase + total)
            seg_nref.append(len(refs))
            cbase += (len(im) + 15) // 16 * 16
            ubase += total
        names = [r[0] for r in refs_all]
        if len(set(names)) != len(names):
            raise FuzError(_lib.FUZ_E_ARG, "the BAM files list the same reference name more than once")
        coff = np.concatenate([t[1] for t in tabs])
        csize = np.concatenate([t[2] for t in tabs])
        uoff = np.concatenate([t[3] for t in tabs] + [np.asarray([ubase], np.int64)])
        crc = np.concatenate([t[4] for t in tabs])
        n_blk = len(coff)
        d_comp = torch.empty(cbase + 16, dtype=torch.uint8, device=dev)
        for t, im in zip(tabs, images):
            d_comp[t[0]:t[0] + len(im)].copy_(torch.from_numpy(im), non_blocking=True)
        d_coff, d_csize = torch.from_numpy(coff).to(dev, non_blocking=True), torch.from_numpy(csize).to(dev, non_blocking=True)
        d_uoff = torch.from_numpy(uoff).to(dev, non_blocking=True)
        d_crc = torch.from_numpy(crc.view(np.int32)).to(dev, non_blocking=True) if verify_crc else None
        raw = torch.empty(ubase + 64, dtype=torch.uint8, device=dev)
        raw[ubase:].zero_()
        seg_start, seg_end = np.asarray(seg_start, np.int64), np.asarray(seg_end, np.int64)
        seg_nref = np.asarray(seg_nref, np.int32)
        cap_bytes = int((seg_end - seg_start).sum())
        rec_out = torch.empty(cap_bytes + 64, dtype=torch.uint8, device=dev)
        ctg_rec_off = torch.zeros(len(refs_all) + 1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)                      # torch's stream -> the context's stream
        _lib.check(self.ctx, lib().fuz_bgzf_inflate(self.ctx, d_comp.data_ptr(), cbase, d_coff.data_ptr(), d_csize.data_ptr(),
                                                    d_uoff.data_ptr(), d_crc.data_ptr() if verify_crc else None, n_blk,
                                                    raw.data_ptr(), ubase))
        cap = cap_bytes // 2048 + 1024
        n_rec, need, nbytes = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        for _ in range(2):
            rec_off = torch.empty(cap + 1, dtype=torch.int64, device=dev)
            torch.cuda.synchronize(dev)
            rc = lib().fuz_bam_index_files(self.ctx, raw.data_ptr(), ubase, len(images), _np_ptr(seg_start), _np_ptr(seg_end),
                                           _np_ptr(seg_nref), cap, rec_out.data_ptr(), cap_bytes, rec_off.data_ptr(),
                                           ctg_rec_off.data_ptr(), C.byref(n_rec), C.byref(need), C.byref(nbytes))
            if rc != _lib.FUZ_E_CAPACITY:
                break
            cap = int(need.value)
        _lib.check(self.ctx, rc)
        total = int(nbytes.value)
        rec_out[total:total + 64].zero_()
        torch.cuda.synchronize(dev)
        h2d = sum(len(im) for im in images) + coff.nbytes + csize.nbytes + uoff.nbytes + (crc.nbytes if verify_crc else 0)
        return DeviceBam(refs_all, rec_out, rec_out.data_ptr(), total, rec_off, ctg_rec_off, int(n_rec.value), int(n_rec.value), h2d)

    def phase_bam(self, image, caps: Optional[Dict[str, int]] = None, verify_crc: bool = True):
        """BAM file image (or a list of images, see ingest_bams) -> (PhaseResult, BamBatchInfo): the BAM is decoded
        on the device, then the four stages run for every reference in one device call (q_ids assigned on the device)."""
        db = self.ingest_bams(image, verify_crc) if isinstance(image, (list, tuple)) else self.ingest_bam(image, verify_crc)
        return self.phase_ingested(db, caps)

    def select_contigs(self, db: DeviceBam, contigs: Sequence[int]) -> DeviceBam:
        """The records of the given references of a device BAM, back to back in a new device buffer (fuz_gather_records), as a
        DeviceBam of its own: what a rank of a multi-GPU run phases when the contigs of one BAM are dealt to the ranks."""
        torch = self._torch
        dev = self.device
        cro = db.ctg_rec_off.cpu().numpy().astype(np.int64)
        contigs = [int(c) for c in contigs]
        counts = np.asarray([cro[c + 1] - cro[c] for c in contigs], np.int64)
        m = int(counts.sum())
        sel = np.concatenate([np.arange(cro[c], cro[c + 1], dtype=np.int64) for c in contigs]) if m else np.zeros(0, np.int64)
        sel_d = torch.from_numpy(sel).to(dev)
        sizes = db.rec_off[sel_d + 1] - db.rec_off[sel_d] if m else torch.zeros(0, dtype=torch.int64, device=dev)
        dst_off = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        if m:
            torch.cumsum(sizes, 0, out=dst_off[1:])
        total = int(dst_off[-1].item())
        dst = torch.zeros(total + 64, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        if m:
            _lib.check(self.ctx, lib().fuz_gather_records(self.ctx, db.rec_ptr, db.rec_off.data_ptr(), db.n_rec, db.rec_bytes,
                                                          sel_d.data_ptr(), dst_off.data_ptr(), m, dst.data_ptr(), total))
            self.sync()
        sub_cro = torch.from_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)).to(dev)
        return DeviceBam([db.refs[c] for c in contigs], dst, dst.data_ptr(), total, dst_off, sub_cro, m, m, 0)

    def phase_ingested(self, db: DeviceBam, caps: Optional[Dict[str, int]] = None):
        """The four stages for every reference of a device BAM (q_ids assigned on the device) -> (PhaseResult, BamBatchInfo)."""
        torch = self._torch
        n_ctg = len(db.refs)
        if n_ctg < 1:
            raise FuzError(_lib.FUZ_E_ARG, "the BAM header lists no reference sequence")
        ctg_len = np.asarray([r[1] for r in db.refs], np.int32)
        t = lib().fuz_tile_size()
        goff = np.concatenate([[0], np.cumsum(np.maximum((ctg_len.astype(np.int64) + t - 1) // t * t, t))]).astype(np.int64)
        dev = self.device
        d_len, d_goff = torch.from_numpy(ctg_len).to(dev), torch.from_numpy(goff).to(dev)
        qid_nq = torch.zeros(n_ctg, dtype=torch.int32, device=dev)
        qid_first = torch.zeros(max(db.n_mapped, 1), dtype=torch.int64, device=dev)
        b = _lib.Batch()
        b.n_ctg, b.n_rec, b.rec_bytes = n_ctg, db.n_mapped, db.rec_bytes
        b.d_rec_buf, b.d_rec_off, b.d_rec_qid = db.rec_ptr, db.rec_off.data_ptr(), None
        b.d_ctg_rec_off, b.d_ctg_len, b.d_ctg_goff = db.ctg_rec_off.data_ptr(), d_len.data_ptr(), d_goff.data_ptr()
        b.d_ctg_nq, b.total_glen, b.total_nq = qid_nq.data_ptr(), int(goff[-1]), 0
        caps = caps or default_caps(int(ctg_len.sum()), db.n_mapped)

        def run(do):
            do.c.d_ctg_nq, do.c.d_name_first = qid_nq.data_ptr(), qid_first.data_ptr()
            _lib.check(self.ctx, lib().fuz_phase_batch(self.ctx, C.byref(b), C.byref(do.c)))
        do, st = self._retry(caps, 0, run)
        arrays = do.fetch(st)
        res = PhaseResult(arrays, int(st.n_sites), int(st.n_vmap), int(st.n_atable), int(st.n_reads),
                          int(st.n_accepted), int(st.aligned_bases), db.h2d_bytes, sum(a.nbytes for a in arrays.values()))
        # QNAME of every q_id: gathered on the device from the first record of each name
        nq = qid_nq.cpu().numpy()
        total_nq = int(nq.sum())
        names = []
        if total_nq:
            names = self.name_rows(db, qid_first[:total_nq])                   # fixed-width byte rows; str on demand (BamBatchInfo.qnames)
            res.d2h_bytes += names.nbytes
        return res, BamBatchInfo([r[0] for r in db.refs], ctg_len, nq, names)

    def name_rows(self, db: DeviceBam, rec_index=None) -> np.ndarray:
        """QNAME of the given records (torch int64 indices on the device; None = every record) of a device BAM as a numpy
        "S" array: gathered on the device into fixed-width rows (NUL and everything behind it zeroed), a few bytes per read
        over PCIe."""
        torch = self._torch
        dev = self.device
        idx = torch.arange(db.n_rec, device=dev) if rec_index is None else rec_index
        if idx.numel() == 0:
            return np.zeros(0, "S1")
        o = db.rec_ptr - db.raw.data_ptr()
        first_off = db.rec_off[idx] + o
        l_name = db.raw[first_off + 12].to(torch.int64)
        width = max(int(l_name.max().item()) - 1, 1)
        col = torch.arange(width, device=dev)[None, :]
        chars_d = db.raw[(first_off[:, None] + 36 + col).clamp_(max=db.raw.numel() - 1)]
        chars_d.masked_fill_(col >= (l_name - 1)[:, None], 0)
        chars = np.ascontiguousarray(chars_d.cpu().numpy())
        return chars.view("S%d" % width).ravel()

    def gather_records(self, db: DeviceBam, sel: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """Records sel[0], sel[1], ... of a device BAM back to back (fuz_gather_records) -> (bytes on the host, offsets
        [len(sel) + 1]).  The bytes sit in the engine's page-locked staging buffer: valid until the next call."""
        torch = self._torch
        dev = self.device
        sel_d = torch.from_numpy(np.ascontiguousarray(sel, dtype=np.int64)).to(dev)
        m = int(sel_d.numel())
        if m == 0:
            return np.zeros(0, np.uint8), np.zeros(1, np.int64)
        if int(sel_d.min().item()) < 0 or int(sel_d.max().item()) >= db.n_rec:
            raise FuzError(_lib.FUZ_E_ARG, "record index out of range")
        sizes = db.rec_off[sel_d + 1] - db.rec_off[sel_d]
        dst_off = torch.zeros(m + 1, dtype=torch.int64, device=dev)
        torch.cumsum(sizes, 0, out=dst_off[1:])
        total = int(dst_off[-1].item())
        dst = torch.empty(total + 16, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)                      # torch's stream -> the context's stream
        _lib.check(self.ctx, lib().fuz_gather_records(self.ctx, db.rec_ptr, db.rec_off.data_ptr(), db.n_rec, db.rec_bytes,
                                                      sel_d.data_ptr(), dst_off.data_ptr(), m, dst.data_ptr(), total))
        # download through a page-locked staging buffer kept by the engine (pageable copies run at a fraction of PCIe)
        pin = getattr(self, "_pin_stage", None)
        if pin is None or pin.numel() < total:
            self._pin_stage = pin = torch.empty(max(total + total // 2, 1 << 20), dtype=torch.uint8).pin_memory()
        self.status()                                    # the context's stream has finished; raises on a device-side error
        pin[:total].copy_(dst[:total], non_blocking=True)
        torch.cuda.synchronize(dev)
        return pin[:total].numpy(), dst_off.cpu().numpy()

    # ---- host-buffer path (what the reference-facing functions and bench e2e use)
    def phase_host(self, pb: PreparedBatch, caps: Optional[Dict[str, int]] = None,
                   host_out: Optional[Dict[str, np.ndarray]] = None) -> PhaseResult:
        caps = caps or default_caps(int(pb.ctg_len.sum()), pb.n_rec)
        hb = _lib.HostBatch()
        hb.n_ctg, hb.n_rec, hb.rec_bytes = pb.n_ctg, pb.n_rec, len(pb.records)
        dev_qid = pb.rec_qid is None
        hb.h_rec_buf, hb.h_rec_off = _np_ptr(pb.records), _np_ptr(pb.rec_off)
        hb.h_rec_qid = None if dev_qid else _np_ptr(pb.rec_qid)
        hb.h_ctg_rec_off, hb.h_ctg_len, hb.h_ctg_nq = _np_ptr(pb.ctg_rec_off), _np_ptr(pb.ctg_len), _np_ptr(pb.ctg_nq)
        q_nq, q_first = np.zeros(pb.n_ctg, np.int32), np.zeros(max(pb.n_rec, 1), np.int64)
        for _ in range(8):
            bufs = host_out if host_out is not None and host_out.get("_caps") == caps else alloc_host_outputs(caps)
            ho = _lib.HostOutputs()
            for key in ("sites", "vmap", "atable", "reads"):
                setattr(ho, "cap_" + key, caps[key])
            for name, _dt, _key, _w in _lib.OUTPUT_ARRAYS:
                setattr(ho, name, _np_ptr(bufs[name]))
            if dev_qid:
                ho.ctg_nq, ho.name_first = _np_ptr(q_nq), _np_ptr(q_first)
            st = _lib.Status()
            up, down = C.c_int64(0), C.c_int64(0)
            rc = lib().fuz_phase_batch_host(self.ctx, C.byref(hb), C.byref(ho), C.byref(st), C.byref(up), C.byref(down))
            if rc == _lib.FUZ_OK:
                if dev_qid:
                    pb.ctg_nq = q_nq
                    pb.name_first = q_first[:int(q_nq.sum())].copy()
                arrays = {}
                for name, _dt, key, width in _lib.OUTPUT_ARRAYS:
                    n = int(getattr(st, _CAP_KEY_N[key]))
                    a = bufs[name][:n * width]
                    arrays[name] = a.reshape(n, width) if width > 1 else a
                return PhaseResult(arrays, int(st.n_sites), int(st.n_vmap), int(st.n_atable), int(st.n_reads),
                                   int(st.n_accepted), int(st.aligned_bases), up.value, down.value)
            if rc != _lib.FUZ_E_CAPACITY:
                _lib.check(self.ctx, rc)
            if st.need_pairs > 0 and st.error_index == 4:
                self.set_option("max_pairs_per_site", int(st.need_pairs // max(caps["sites"], 1) + 2))
            self._grow_scratch(st)
            new_caps = _grow(caps, st)
            if new_caps != caps:
                host_out = None
            caps = new_caps
        raise FuzError(_lib.FUZ_E_CAPACITY, "capacity retry did not converge")


def alloc_host_outputs(caps: Dict[str, int], pin: bool = False) -> Dict[str, np.ndarray]:
    out: Dict[str, object] = {"_caps": dict(caps)}
    keep = []
    for name, dt, key, width in _lib.OUTPUT_ARRAYS:
        n = max(caps[key] * width, 4)
        if pin:
            import torch
            t = torch.empty(n, dtype=torch.int32 if dt == "i4" else torch.uint8).pin_memory()
            keep.append(t)
            out[name] = t.numpy()
        else:
            out[name] = np.empty(n, dtype=np.int32 if dt == "i4" else np.uint8)
    out["_keep"] = keep
    return out


_engines: Dict[int, Engine] = {}


def get_engine(device: Optional[int] = None) -> Engine:
    """Process-wide engine per device (contexts are cheap to keep, costly to create).
    device None = torch's current CUDA device (rank-local GPU in multi-process runs)."""
    if device is None:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    e = _engines.get(device)
    if e is None or not e.ctx:
        e = Engine(device)
        _engines[device] = e
    return e
