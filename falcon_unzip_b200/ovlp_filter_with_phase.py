"""Drop-in mirror of reference falcon_unzip/ovlp_filter_with_phase.py (the three-stage overlap
filter that drops cross-phase overlaps): same function names, the same
``input_`` tuples, the same CLI (ovlp_filter_with_phase.py:293-307) and byte-identical output --
with the phase test, the per-read end counts, the contained set and the per-read best-n selection
computed by the CUDA kernels of libfuz.so (fuz_ovlp_filter).  There is no CPU fallback.

Host work kept here: LA4Falcon -mo text -> int arrays and the output text (C++ in libfuz), the
rid -> (ctg, block, phase) table with its strings interned, and the re-sort of the rare read whose
candidates tie on every numeric sort key (the reference then compares the remaining text columns).
``main`` runs all LAS files of the fofn through ONE device call instead of three pool passes.
"""
from __future__ import annotations

import argparse
import ctypes as C
import shlex
import subprocess
import sys
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib, engine, la4falcon
from ._lib import FuzError, lib

arid2phase: Dict[str, Tuple[str, str, str]] = {}

_COLS = ("q", "t", "len", "qs", "qe", "ql", "ts", "te", "tl")


def read_las_lines(db_fn: str, fn: str) -> bytes:
    """Output of ``LA4Falcon -mo <db> <las>`` (:60,:150,:195).  Tests replace this function."""
    p = subprocess.run(shlex.split("LA4Falcon -mo %s %s" % (db_fn, fn)), stdout=subprocess.PIPE)
    if p.returncode != 0:
        raise RuntimeError("LA4Falcon failed on %s" % fn)
    return p.stdout


class Lines:
    """Parsed LA4Falcon -mo text of one or more files (the text is kept: selected lines are printed again)."""

    def __init__(self, blobs: Sequence[bytes]):
        norm = []
        for b in blobs:
            if not isinstance(b, (bytes, bytearray, memoryview)):        # an iterable of text lines
                b = "".join(x if x.endswith("\n") else x + "\n" for x in b).encode("ascii")
            b = bytes(b)
            norm.append(b if not b or b.endswith(b"\n") else b + b"\n")
        blobs = norm
        self.text = b"".join(blobs)
        cap = self.text.count(b"\n") + 1
        self.a = {k: np.empty(cap, np.int32) for k in _COLS}
        self.a["flags"] = np.empty(cap, np.uint8)
        self.a["off"] = np.empty(cap, np.int64)
        self.a["llen"] = np.empty(cap, np.int32)
        n = lib().fuz_host_parse_la4falcon_mo(self.text, len(self.text), cap, *[self.a[k].ctypes.data for k in _COLS],
                                              self.a["flags"].ctypes.data, self.a["off"].ctypes.data, self.a["llen"].ctypes.data)
        if n == -2:
            raise FuzError(_lib.FUZ_E_FORMAT, "read ids of the overlap lines must be %09d ids (the reference compares them as strings)")
        if n < 0:
            raise ValueError("malformed LA4Falcon -mo line (12+ columns: ids, lengths and coordinates as integers, identity as float)")
        self.n = int(n)
        for k in self.a:
            self.a[k] = self.a[k][:self.n]
        # file index of every line from the blob boundaries
        ends = np.cumsum([len(b) for b in blobs])
        self.file = np.searchsorted(ends, self.a["off"], side="right").astype(np.int32)

    def tokens(self, i: int) -> List[str]:
        o = int(self.a["off"][i])
        return self.text[o:o + int(self.a["llen"][i])].decode("ascii").split()


class PhaseTable:
    """arid2phase (:319-322) as arrays over the read-id space; strings interned so that the kernels
    compare ints where the reference compares strings."""

    def __init__(self, a2p: Dict[str, Tuple[str, str, str]], n_reads: int = 0):
        ids = []
        for k in a2p:
            if len(k) != 9 or not k.isdigit():
                raise FuzError(_lib.FUZ_E_FORMAT, "rid_phase_map key %r is not a %%09d read id" % (k,))
            ids.append(int(k))
        self.n_reads = max(n_reads, (max(ids) + 1) if ids else 0, 1)
        self.in_map = np.zeros(self.n_reads, np.uint8)
        self.ctg, self.blk, self.ph = (np.full(self.n_reads, -1, np.int32) for _ in range(3))
        intern: Dict[str, int] = {}
        texts = [b""] * self.n_reads
        for k, r in zip(a2p, ids):
            c, b, p = a2p[k]
            self.in_map[r] = 1
            self.ctg[r] = intern.setdefault(c, len(intern))
            self.blk[r] = intern.setdefault(b, len(intern))
            self.ph[r] = intern.setdefault(p, len(intern))
            texts[r] = ("%s.%s.%s" % (c, b, p)).encode("ascii")
        self.phase_off = np.concatenate([[0], np.cumsum([len(x) for x in texts])]).astype(np.int64)
        self.phase_text = b"".join(texts) + b"\0"

    def flags_of(self, ids) -> np.ndarray:
        """uint8 [n_reads] from a set of read-id strings (None and foreign strings never match an id)."""
        f = np.zeros(self.n_reads, np.uint8)
        for x in ids:
            if isinstance(x, str) and len(x) == 9 and x.isdigit() and int(x) < self.n_reads:
                f[int(x)] = 1
        return f


def _upload(L: Lines, tab: PhaseTable, ignore_in: Optional[np.ndarray] = None, contained_in: Optional[np.ndarray] = None) -> dict:
    """Columns and tables as device tensors (PyTorch owns the memory)."""
    import torch
    dev = engine.get_engine().device

    def up(a):
        a = np.ascontiguousarray(a)
        return torch.from_numpy(a).to(dev) if a.size else torch.zeros(1, dtype=getattr(torch, a.dtype.name), device=dev)
    if hasattr(L, "d"):                                   # parsed on the device: the columns are there already
        d = {k: L.d[k] for k in _COLS + ("flags", "file")}
    else:
        d = {k: up(L.a[k]) for k in _COLS + ("flags",)}
        d["file"] = up(L.file)
    d.update(in_map=up(tab.in_map), ph_ctg=up(tab.ctg), ph_block=up(tab.blk), ph_phase=up(tab.ph))
    if ignore_in is not None:
        d["ignore_in"] = up(ignore_in)
    if contained_in is not None:
        d["contained_in"] = up(contained_in)
    return d


def _device_filter(L: Lines, tab: PhaseTable, max_diff: int, max_ovlp: int, min_ovlp: int, min_len: int, bestn: int, stage: int,
                   ignore_in: Optional[np.ndarray] = None, contained_in: Optional[np.ndarray] = None, d: Optional[dict] = None) -> dict:
    import torch
    eng = engine.get_engine()
    dev = eng.device
    if d is None:
        d = _upload(L, tab, ignore_in, contained_in)
    ignore_in = d.get("ignore_in")
    contained_in = d.get("contained_in")
    n, nr = L.n, tab.n_reads
    cap_groups, cap_out = max(16, min(n, 2 * nr) + 16), max(1024, n // 2)
    for _ in range(4):
        o = dict(ignore=torch.zeros(nr, dtype=torch.uint8, device=dev), contained=torch.zeros(nr, dtype=torch.uint8, device=dev),
                 grp_q=torch.zeros(cap_groups, dtype=torch.int32, device=dev), grp_line=torch.zeros(cap_groups, dtype=torch.int32, device=dev),
                 grp_ignore=torch.zeros(cap_groups, dtype=torch.uint8, device=dev), grp_tie=torch.zeros(cap_groups, dtype=torch.uint8, device=dev),
                 grp_off=torch.zeros(cap_groups + 1, dtype=torch.int32, device=dev), out_line=torch.zeros(cap_out, dtype=torch.int32, device=dev),
                 cand=torch.zeros(max(n, 1), dtype=torch.uint8, device=dev))
        torch.cuda.synchronize(dev)
        fi = _lib.OvlpInput()
        fi.n_ovl, fi.n_reads = n, nr
        for k in _COLS + ("flags", "file", "in_map", "ph_ctg", "ph_block", "ph_phase"):
            setattr(fi, "d_" + k, d[k].data_ptr())
        fi.max_diff, fi.max_ovlp, fi.min_ovlp, fi.min_len, fi.bestn, fi.stage = max_diff, max_ovlp, min_ovlp, min_len, bestn, stage
        fi.d_ignore_in = d["ignore_in"].data_ptr() if ignore_in is not None else None
        fi.d_contained_in = d["contained_in"].data_ptr() if contained_in is not None else None
        fo = _lib.OvlpOutputs()
        fo.cap_groups, fo.cap_out = cap_groups, cap_out
        for k in o:
            setattr(fo, "d_" + k, o[k].data_ptr())
        _lib.check(eng.ctx, lib().fuz_ovlp_filter(eng.ctx, C.byref(fi), C.byref(fo)))
        st = eng.status(raise_on_error=False)
        if st.error == _lib.FUZ_OK:
            ng, n_out = int(st.reserved[0]), int(st.reserved[1])
            r = dict(n_groups=ng, ignore=o["ignore"].cpu().numpy(), contained=o["contained"].cpu().numpy(),
                     grp_q=o["grp_q"][:ng].cpu().numpy(), grp_line=o["grp_line"][:ng].cpu().numpy(),
                     grp_ignore=o["grp_ignore"][:ng].cpu().numpy())
            if stage == 3:
                r.update(grp_tie=o["grp_tie"][:ng].cpu().numpy(), grp_off=o["grp_off"][:ng + 1].cpu().numpy(),
                         out_line=o["out_line"][:n_out].cpu().numpy())
                r["cand"] = o["cand"][:n].cpu().numpy() if r["grp_tie"].any() else None
            return r
        if st.error == _lib.FUZ_E_CAPACITY and st.error_index == 10:
            cap_groups = int(st.reserved[0]) + 16
            continue
        if st.error == _lib.FUZ_E_CAPACITY and st.error_index == 11:
            cap_out = int(st.reserved[2]) + 16
            continue
        if st.error == _lib.FUZ_E_CAPACITY and st.error_index == 9:
            raise FuzError(st.error, "a read has more than 512 candidate overlaps on one end (max_cov beyond the kernel's capacity)")
        eng.status()
    raise FuzError(_lib.FUZ_E_CAPACITY, "capacity retry did not converge")


def _verdict(left: int, right: int, max_diff: int, max_ovlp: int, min_ovlp: int) -> bool:
    return abs(left - right) > max_diff or left > max_ovlp or right > max_ovlp or left < min_ovlp or right < min_ovlp


def _resolve_ties(L: Lines, tab: PhaseTable, r: dict, bestn: int) -> np.ndarray:
    """Selected lines, with the groups the kernel flagged re-sorted the way the reference sorts them: tuples
    (-inphase, -len, range, token list) -- the token list decides among candidates that tie on every number
    (:234,:275).  The candidates of both ends come from the device (d_cand); only their order is settled here.
    Groups without such ties come straight from the device."""
    tie = np.flatnonzero(r["grp_tie"])
    if len(tie) == 0:
        return r["out_line"].astype(np.int64)
    a, out, cand, parts, at = L.a, r["out_line"], r["cand"], [], 0
    for g in tie.tolist():
        lo = int(r["grp_line"][g])
        hi = int(r["grp_line"][g + 1]) if g + 1 < r["n_groups"] else L.n      # passing lines in [lo, hi) all belong to g
        q = int(r["grp_q"][g])
        sel = []
        for side in (1, 2):
            rows = []
            for i in (lo + np.flatnonzero(cand[lo:hi] == side)).tolist():
                t = int(a["t"][i])
                inphase = 1 if (tab.ctg[t], tab.blk[t], tab.ph[t]) == (tab.ctg[q], tab.blk[q], tab.ph[q]) else 0
                rows.append((-inphase, -int(a["len"][i]), int(a["tl"][i]) - (int(a["te"][i]) - int(a["ts"][i])), L.tokens(i), i))
            rows.sort(key=lambda c: c[:4])
            for k, c in enumerate(rows):
                sel.append(c[4])
                if k >= bestn and c[2] > 1000:
                    break
        parts.append(out[at:int(r["grp_off"][g])])
        parts.append(np.asarray(sel, np.int64))
        at = int(r["grp_off"][g + 1])
    parts.append(out[at:])
    return np.concatenate([p.astype(np.int64) for p in parts])


def _format_piece(text: bytes, g: dict, tab: PhaseTable) -> bytes:
    n = len(g["off"])
    idx = np.arange(n, dtype=np.int64)
    off, llen, q, t = (np.ascontiguousarray(g[k]) for k in ("off", "llen", "q", "t"))
    args = (text, off.ctypes.data, llen.ctypes.data, q.ctypes.data, t.ctypes.data, idx.ctypes.data, n, tab.phase_text,
            tab.phase_off.ctypes.data)
    size = lib().fuz_host_format_ovlp(*args, None, 0)
    if size < 0:
        raise FuzError(_lib.FUZ_E_ARG, "fuz_host_format_ovlp failed")
    buf = C.create_string_buffer(int(size) + 1)
    lib().fuz_host_format_ovlp(*args, buf, size)
    return buf.raw[:size]


def _format(L, tab: PhaseTable, sel: np.ndarray) -> bytes:
    """Output text of the selected lines (fuz_host_format_ovlp)."""
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    if not hasattr(L, "gather"):
        return _format_piece(L.text, {k: L.a[k][sel] for k in ("off", "llen", "q", "t")}, tab)
    # columns live on the device: fetch the selected lines only; the text of every file is formatted from its own
    # buffer (no concatenation of the files on the host) -- selected lines come file after file
    g = L.gather(("off", "llen", "q", "t"), sel)
    starts = L.starts
    k_of = np.searchsorted(starts, g["off"], side="right") - 1
    out = []
    for k in np.unique(k_of).tolist():
        m = k_of == k
        piece = {key: g[key][m] for key in ("llen", "q", "t")}
        piece["off"] = g["off"][m] - starts[k]
        out.append(_format_piece(L.blobs[k], piece, tab))
    return b"".join(out)


# --------------------------------------------------------------------------- the reference's functions
def filter_stage1(input_):
    """reference :49-143 -> (fn, ids to ignore in the order the reference appends them)."""
    db_fn, fn, max_diff, max_ovlp, min_ovlp, min_len = input_
    L, tab = la4falcon.DeviceLines([read_las_lines(db_fn, fn)], require_id9=True), PhaseTable(arid2phase)
    r = _device_filter(L, tab, max_diff, max_ovlp, min_ovlp, min_len, 0, 1)
    rtn: List[Optional[str]] = []
    if r["n_groups"] and _verdict(0, 0, max_diff, max_ovlp, min_ovlp):
        rtn.append(None)                                  # the run of `None` judged on counts (0, 0), :75-87
    rtn.extend("%09d" % q for q in r["grp_q"][r["grp_ignore"].astype(bool)].tolist())
    return fn, rtn


def filter_stage2(input_):
    """reference :145-186 -> (fn, set of contained read ids)."""
    db_fn, fn, max_diff, max_ovlp, min_ovlp, min_len, ignore_set = input_
    L, tab = la4falcon.DeviceLines([read_las_lines(db_fn, fn)], require_id9=True), PhaseTable(arid2phase)
    r = _device_filter(L, tab, max_diff, max_ovlp, min_ovlp, min_len, 0, 2, ignore_in=tab.flags_of(ignore_set))
    return fn, set("%09d" % x for x in np.flatnonzero(r["contained"]).tolist())


def filter_stage3(input_):
    """reference :188-290 -> (fn, selected overlaps as token lists, phase strings appended)."""
    db_fn, fn, max_diff, max_ovlp, min_ovlp, min_len, ignore_set, contained_set, bestn = input_
    L, tab = la4falcon.DeviceLines([read_las_lines(db_fn, fn)], require_id9=True), PhaseTable(arid2phase)
    ig, ct = tab.flags_of(ignore_set), tab.flags_of(contained_set)
    r = _device_filter(L, tab, max_diff, max_ovlp, min_ovlp, min_len, bestn, 3, ignore_in=ig, contained_in=ct)
    sel = _resolve_ties(L, tab, r, bestn)
    return fn, [x.split(" ") for x in _format(L, tab, sel).decode("ascii").splitlines()]


def run_ovlp_filter(file_list: Sequence[str], db_fn: str, max_diff: int, max_cov: int, min_cov: int, min_len: int, bestn: int,
                    device_parse: bool = True) -> bytes:
    """main() (:324-352) for all LAS files in one device call -> the text the reference prints.  The LA4Falcon
    text is parsed on the device (fuz_parse_la4falcon); device_parse=False takes the host parser."""
    blobs = [read_las_lines(db_fn, fn) for fn in file_list]
    L = la4falcon.DeviceLines(blobs, require_id9=True) if device_parse else Lines(blobs)
    tab = PhaseTable(arid2phase)
    r = _device_filter(L, tab, max_diff, max_cov, min_cov, min_len, bestn, 3)
    sel = _resolve_ties(L, tab, r, bestn)
    return _format(L, tab, sel)


def run_ovlp_filter_sharded(file_list: Sequence[str], db_fn: str, max_diff: int, max_cov: int, min_cov: int, min_len: int, bestn: int,
                            rank: int, world_size: int, group=None) -> Optional[bytes]:
    """run_ovlp_filter over `world_size` processes (one per GPU; torch.distributed initialised by the caller).
    It mirrors the reference's three pool passes (:324-352), whose only cross-file products are two SETS:

      stage 1   LAS files are dealt round-robin to the ranks; every rank judges the reads of its files
      exchange  ignore set = union over files (:327-330): ONE all-reduce (max) of n_reads flag bytes
      stage 2   contained reads of the rank's files, given the global ignore set
      exchange  contained set = union over files (:337-339): ONE all-reduce (max) of n_reads flag bytes
      stage 3   selection for the rank's files with both global sets; the text of every file is collected by
                rank 0 and returned in fofn order (other ranks return None).

    The rid -> phase table is replicated (every rank reads rid_phase_map)."""
    import torch
    import torch.distributed as dist
    mine = list(range(rank, len(file_list), world_size))
    backend = dist.get_backend(group)
    blobs = [read_las_lines(db_fn, file_list[i]) for i in mine]
    L = la4falcon.DeviceLines(blobs, require_id9=True) if backend == "nccl" else Lines(blobs)
    tab = PhaseTable(arid2phase)
    dev = engine.get_engine().device if backend == "nccl" else torch.device("cpu")

    def union(flags: np.ndarray) -> np.ndarray:
        t = torch.from_numpy(np.ascontiguousarray(flags, dtype=np.uint8)).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return t.cpu().numpy()
    r1 = _device_filter(L, tab, max_diff, max_cov, min_cov, min_len, 0, 1)
    ignore = union(r1["ignore"])
    r2 = _device_filter(L, tab, max_diff, max_cov, min_cov, min_len, 0, 2, ignore_in=ignore)
    contained = union(r2["contained"])
    r3 = _device_filter(L, tab, max_diff, max_cov, min_cov, min_len, bestn, 3, ignore_in=ignore, contained_in=contained)
    sel = _resolve_ties(L, tab, r3, bestn)
    # text per file (selected lines are in (file, line-group) order already)
    per_file = {}
    f_of = (L.gather(("file",), sel)["file"] if hasattr(L, "gather") else L.file[sel]) if len(sel) else np.zeros(0, np.int32)
    for k, i in enumerate(mine):
        per_file[i] = _format(L, tab, sel[f_of == k])
    gathered = [None] * world_size
    dist.all_gather_object(gathered, per_file, group=group)
    if rank != 0:
        return None
    merged = {}
    for g in gathered:
        merged.update(g)
    return b"".join(merged[i] for i in range(len(file_list)))


def parse_args(argv):
    parser = argparse.ArgumentParser(description="a simple multi-processes LAS ovelap data filter")
    parser.add_argument("--n_core", type=int, default=4, help="accepted for compatibility (the work runs on the GPU)")
    parser.add_argument("--fofn", type=str, help="file contains the path of all LAS file to be processed in parallel")
    parser.add_argument("--db", type=str, help="read db file path")
    parser.add_argument("--max_diff", type=int, help="max difference of 5' and 3' coverage")
    parser.add_argument("--max_cov", type=int, help="max coverage of 5' or 3' coverage")
    parser.add_argument("--min_cov", type=int, help="min coverage of 5' or 3' coverage")
    parser.add_argument("--min_len", type=int, default=2500, help="min length of the reads")
    parser.add_argument("--bestn", type=int, default=10, help="output at least best n overlaps on 5' or 3' ends if possible")
    parser.add_argument("--rid_phase_map", type=str, help="the file that encode the relationship of the read id to phase blocks",
                        required=True)
    return parser.parse_args(argv[1:])


def main(argv=sys.argv):
    args = parse_args(argv)
    arid2phase.clear()
    with open(args.rid_phase_map) as f:
        for row in f:
            row = row.strip().split()
            arid2phase[row[0]] = (row[1], row[2], row[3])      # ctg_id, phase_blk_id, phase_id (:322)
    with open(args.fofn) as f:
        file_list = [fn for fn in f.read().split("\n") if len(fn) != 0]
    out = run_ovlp_filter(file_list, args.db, args.max_diff, args.max_cov, args.min_cov, args.min_len, args.bestn)
    sys.stdout.write(out.decode("ascii"))


if __name__ == "__main__":
    main()
