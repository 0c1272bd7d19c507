"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes front-end of oracle/phasing_oracle.c plus
the file-level stage functions (same files, same bytes as reference
falcon_unzip/phasing.py:14-480).  Used by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never by falcon_unzip_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import py2emu

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "phasing_oracle.c")
_SO = os.path.join(_HERE, "_build", "libphasing_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 the C restatement into oracle/_build/ (idempotent)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        tmp = _SO + ".tmp%d" % os.getpid()
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", tmp, _SRC])
        os.replace(tmp, _SO)
    return _SO


class _HetRes(C.Structure):
    _fields_ = [("n_sites", C.c_int64), ("site_pos", C.POINTER(C.c_int32)),
                ("site_total", C.POINTER(C.c_int32)), ("site_base", C.POINTER(C.c_uint8)),
                ("site_count", C.POINTER(C.c_int32)), ("n_vmap", C.c_int64),
                ("vm_pos", C.POINTER(C.c_int32)), ("vm_allele", C.POINTER(C.c_uint8)),
                ("vm_qid", C.POINTER(C.c_int32)), ("n_accepted", C.c_int64),
                ("aligned_bases", C.c_int64), ("pos_last", C.c_int32)]


class _ATRes(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("pos1", C.POINTER(C.c_int32)),
                ("pos2", C.POINTER(C.c_int32)), ("b", C.POINTER(C.c_uint8)),
                ("ct", C.POINTER(C.c_int32))]


class _BlkRes(C.Structure):
    _fields_ = [("n_v", C.c_int64), ("pid", C.POINTER(C.c_int32)), ("pos", C.POINTER(C.c_int32)),
                ("h", C.POINTER(C.c_uint8)), ("lext", C.POINTER(C.c_int32)),
                ("rext", C.POINTER(C.c_int32)), ("lscore", C.POINTER(C.c_int32)),
                ("rscore", C.POINTER(C.c_int32)), ("n_blocks", C.c_int32)]


class _RdRes(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("qid", C.POINTER(C.c_int32)), ("pid", C.POINTER(C.c_int32)),
                ("phase", C.POINTER(C.c_int32)), ("n0", C.POINTER(C.c_int32)),
                ("n1", C.POINTER(C.c_int32))]


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _check(rc, what):
    if rc:
        raise RuntimeError("oracle %s failed: status %d (%s)" % (
            what, rc, {1: "out of memory", 2: "bad record: the reference raises here",
                       3: "input outside the reference-produced format"}.get(rc, "?")))


# --------------------------------------------------------------------------- array level
def assign_qids(names: Sequence[str]) -> Tuple[np.ndarray, List[str]]:
    """first-seen QNAME -> q_id (phasing.py:47-54); returns per-record q_id and the name list."""
    table: Dict[str, int] = {}
    qid = np.empty(len(names), dtype=np.int32)
    for i, nm in enumerate(names):
        q = table.get(nm)
        if q is None:
            q = len(table)
            table[nm] = q
        qid[i] = q
    return qid, list(table)


def index_records(records) -> np.ndarray:
    """Offsets of the records in a concatenated BAM record buffer (block_size chain)."""
    mv = memoryview(records)
    offs, o, n = [0], 0, len(mv)
    while o < n:
        o += 4 + int.from_bytes(mv[o:o + 4], "little", signed=True)
        offs.append(o)
    if o != n:
        raise ValueError("truncated BAM record stream")
    return np.asarray(offs, dtype=np.int64)


def record_names(records, rec_off: np.ndarray) -> List[str]:
    mv = memoryview(records)
    out = []
    for i in range(len(rec_off) - 1):
        o = int(rec_off[i])
        l_name = mv[o + 12]
        out.append(bytes(mv[o + 36:o + 36 + l_name - 1]).decode("ascii"))
    return out


def het_call(records, rec_off: np.ndarray, qid: np.ndarray) -> dict:
    recs = np.frombuffer(records, dtype=np.uint8) if not isinstance(records, np.ndarray) else records
    recs = np.ascontiguousarray(recs)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
    qid = np.ascontiguousarray(qid, dtype=np.int32)
    res = _HetRes()
    rc = lib().fo_make_het_call(_p(recs, C.c_uint8), _p(rec_off, C.c_int64),
                                C.c_int64(len(rec_off) - 1), _p(qid, C.c_int32), C.byref(res))
    _check(rc, "make_het_call")
    out = dict(site_pos=_arr(res.site_pos, res.n_sites, np.int32),
               site_total=_arr(res.site_total, res.n_sites, np.int32),
               site_base=_arr(res.site_base, 4 * res.n_sites, np.uint8).reshape(-1, 4),
               site_count=_arr(res.site_count, 4 * res.n_sites, np.int32).reshape(-1, 4),
               vm_pos=_arr(res.vm_pos, res.n_vmap, np.int32),
               vm_allele=_arr(res.vm_allele, res.n_vmap, np.uint8),
               vm_qid=_arr(res.vm_qid, res.n_vmap, np.int32),
               n_accepted=int(res.n_accepted), aligned_bases=int(res.aligned_bases),
               pos_last=int(res.pos_last))
    lib().fo_free_hetcall(C.byref(res))
    return out


def pileup_counts(records, rec_off: np.ndarray, ctg_len: int) -> np.ndarray:
    """[ctg_len, 4] depth of A,C,G,T over accepted records (phasing.py:63-96)."""
    recs = np.frombuffer(records, dtype=np.uint8) if not isinstance(records, np.ndarray) else records
    recs = np.ascontiguousarray(recs)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.int64)
    counts = np.zeros((ctg_len, 4), dtype=np.uint32)
    rc = lib().fo_pileup_counts(_p(recs, C.c_uint8), _p(rec_off, C.c_int64), C.c_int64(len(rec_off) - 1),
                                C.c_int64(ctg_len), _p(counts, C.c_uint32))
    _check(rc, "pileup_counts")
    return counts


def association_table(vm_pos, vm_allele, vm_qid) -> dict:
    vm_pos = np.ascontiguousarray(vm_pos, dtype=np.int32)
    vm_allele = np.ascontiguousarray(vm_allele, dtype=np.uint8)
    vm_qid = np.ascontiguousarray(vm_qid, dtype=np.int32)
    res = _ATRes()
    rc = lib().fo_association_table(_p(vm_pos, C.c_int32), _p(vm_allele, C.c_uint8),
                                    _p(vm_qid, C.c_int32), C.c_int64(len(vm_pos)), C.byref(res))
    _check(rc, "generate_association_table")
    out = dict(pos1=_arr(res.pos1, res.n_rows, np.int32), pos2=_arr(res.pos2, res.n_rows, np.int32),
               b=_arr(res.b, 4 * res.n_rows, np.uint8).reshape(-1, 4),
               ct=_arr(res.ct, 4 * res.n_rows, np.int32).reshape(-1, 4))
    lib().fo_free_atable(C.byref(res))
    return out


def phased_blocks(pos1, pos2, b, ct) -> dict:
    pos1 = np.ascontiguousarray(pos1, dtype=np.int32)
    pos2 = np.ascontiguousarray(pos2, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    ct = np.ascontiguousarray(ct, dtype=np.int32).reshape(-1)
    res = _BlkRes()
    rc = lib().fo_phased_blocks(_p(pos1, C.c_int32), _p(pos2, C.c_int32), _p(b, C.c_uint8),
                                _p(ct, C.c_int32), C.c_int64(len(pos1)), C.byref(res))
    _check(rc, "get_phased_blocks")
    n = res.n_v
    out = dict(pid=_arr(res.pid, n, np.int32), pos=_arr(res.pos, n, np.int32),
               h=_arr(res.h, 2 * n, np.uint8).reshape(-1, 2), lext=_arr(res.lext, n, np.int32),
               rext=_arr(res.rext, n, np.int32), lscore=_arr(res.lscore, n, np.int32),
               rscore=_arr(res.rscore, n, np.int32), n_blocks=int(res.n_blocks))
    lib().fo_free_blocks(C.byref(res))
    return out


def phased_reads(vm_pos, vm_allele, vm_qid, v_pid, v_pos, v_h) -> dict:
    vm_pos = np.ascontiguousarray(vm_pos, dtype=np.int32)
    vm_allele = np.ascontiguousarray(vm_allele, dtype=np.uint8)
    vm_qid = np.ascontiguousarray(vm_qid, dtype=np.int32)
    v_pid = np.ascontiguousarray(v_pid, dtype=np.int32)
    v_pos = np.ascontiguousarray(v_pos, dtype=np.int32)
    v_h = np.ascontiguousarray(v_h, dtype=np.uint8).reshape(-1)
    res = _RdRes()
    rc = lib().fo_phased_reads(_p(vm_pos, C.c_int32), _p(vm_allele, C.c_uint8), _p(vm_qid, C.c_int32),
                               C.c_int64(len(vm_pos)), _p(v_pid, C.c_int32), _p(v_pos, C.c_int32),
                               _p(v_h, C.c_uint8), C.c_int64(len(v_pid)), C.byref(res))
    _check(rc, "get_phased_reads")
    n = res.n_rows
    out = dict(qid=_arr(res.qid, n, np.int32), pid=_arr(res.pid, n, np.int32),
               phase=_arr(res.phase, n, np.int32), n0=_arr(res.n0, n, np.int32),
               n1=_arr(res.n1, n, np.int32))
    lib().fo_free_reads(C.byref(res))
    return out


# --------------------------------------------------------------------------- text level
def format_variant_pos(h: dict, ref_seq: str) -> str:
    rows = []
    for i in range(len(h["site_pos"])):
        p = int(h["site_pos"][i])
        rows.append("%d %s %d %s\n" % (p + 1, ref_seq[p], h["site_total"][i], " ".join(
            "%s %d" % (chr(h["site_base"][i, k]), h["site_count"][i, k]) for k in range(4))))
    return "".join(rows)


def format_variant_map(h: dict, ref_seq: str) -> str:
    return "".join("%d %s %s %d\n" % (p + 1, ref_seq[p], chr(a), q)
                   for p, a, q in zip(h["vm_pos"].tolist(), h["vm_allele"].tolist(),
                                      h["vm_qid"].tolist()))


def parse_variant_map(text: str):
    pos, ref, al, qid = [], [], [], []
    for line in text.splitlines():
        f = line.split()
        if not f:
            continue
        pos.append(int(f[0])); ref.append(f[1]); al.append(ord(f[2])); qid.append(int(f[3]))
    return (np.asarray(pos, np.int32), ref, np.asarray(al, np.uint8), np.asarray(qid, np.int32))


def make_het_call_files(records, ref_seq: str, vmap_fn: str, vpos_fn: str, q_id_map_fn: str) -> dict:
    """phasing.py:14-134 on BAM records of ONE contig (the `samtools view bam ctg` set)."""
    rec_off = index_records(records)
    qid, names = assign_qids(record_names(records, rec_off))
    h = het_call(records, rec_off, qid)
    for p in (vmap_fn, vpos_fn, q_id_map_fn):
        os.makedirs(os.path.dirname(p) or ".", exist_ok=True)
    with open(vpos_fn, "w") as f:
        f.write(format_variant_pos(h, ref_seq))
    with open(vmap_fn, "w") as f:
        f.write(format_variant_map(h, ref_seq))
    with open(q_id_map_fn, "w") as f:      # dense int keys iterate ascending (B.3)
        f.write("".join("%d %s\n" % (i, nm) for i, nm in enumerate(names)))
    return h


def generate_association_table_files(vmap_fn: str, atable_fn: str) -> dict:
    with open(vmap_fn) as f:
        pos, _ref, al, qid = parse_variant_map(f.read())
    t = association_table(pos, al, qid)
    os.makedirs(os.path.dirname(atable_fn) or ".", exist_ok=True)
    with open(atable_fn, "w") as f:
        f.write("".join("%d %s %s %d %s %s %d %d %d %d\n" % (
            t["pos1"][i], chr(t["b"][i, 0]), chr(t["b"][i, 1]), t["pos2"][i], chr(t["b"][i, 2]),
            chr(t["b"][i, 3]), t["ct"][i, 0], t["ct"][i, 1], t["ct"][i, 2], t["ct"][i, 3])
            for i in range(len(t["pos1"]))))
    return t


def parse_atable(text: str):
    p1, p2, b, ct = [], [], [], []
    for line in text.splitlines():
        f = line.split()
        if not f:
            continue
        p1.append(int(f[0])); p2.append(int(f[3]))
        b.append([ord(f[1]), ord(f[2]), ord(f[4]), ord(f[5])])
        ct.append([int(x) for x in f[6:10]])
    return (np.asarray(p1, np.int32), np.asarray(p2, np.int32),
            np.asarray(b, np.uint8).reshape(-1, 4), np.asarray(ct, np.int32).reshape(-1, 4))


def get_phased_blocks_files(vmap_fn: str, atable_fn: str, p_variant_fn: str) -> dict:
    with open(vmap_fn) as f:
        pos, ref, _al, _qid = parse_variant_map(f.read())
    ref_base = dict(zip(pos.tolist(), ref))                         # :230-238
    with open(atable_fn) as f:
        p1, p2, b, ct = parse_atable(f.read())
    r = phased_blocks(p1, p2, b, ct)
    out = []
    for pid in range(1, r["n_blocks"] + 1):
        idx = np.flatnonzero(r["pid"] == pid)
        ps = r["pos"][idx]
        mn, mx = int(ps.min()), int(ps.max())
        out.append("P %d %d %d %d %d %s\n" % (pid, mn, mx, mx - mn, len(idx),
                                                py2emu.py27_float_str(1.0 * (mx - mn) / len(idx))))
        for i in idx:
            p = int(r["pos"][i]); rb = ref_base[p]
            out.append("V %d %d %d_%s_%s %d_%s_%s %d %d %d %d\n" % (
                pid, p, p, rb, chr(r["h"][i, 0]), p, rb, chr(r["h"][i, 1]), r["lext"][i],
                r["rext"][i], r["lscore"][i], r["rscore"][i]))
    os.makedirs(os.path.dirname(p_variant_fn) or ".", exist_ok=True)
    with open(p_variant_fn, "w") as f:
        f.write("".join(out))
    return r


def get_phased_reads_files(vmap_fn: str, q_id_map_fn: str, p_variant_fn: str, ctg_id: str,
                           phased_read_fn: str) -> dict:
    rid_map = {}
    with open(q_id_map_fn) as f:
        for line in f:
            l = line.split()
            rid_map[int(l[0])] = l[1]
    with open(vmap_fn) as f:
        pos, _ref, al, qid = parse_variant_map(f.read())
    v_pid, v_pos, v_h = [], [], []
    with open(p_variant_fn) as f:
        for line in f:
            l = line.split()
            if not l or l[0] != "V":
                continue
            v_pid.append(int(l[1])); v_pos.append(int(l[2]))
            v_h.append([ord(l[3].split("_")[2]), ord(l[4].split("_")[2])])
    r = phased_reads(pos, al, qid, np.asarray(v_pid, np.int32), np.asarray(v_pos, np.int32),
                     np.asarray(v_h, np.uint8).reshape(-1, 2))
    # row order = py2 dict order of read_to_variants (B.3), keys in first-appearance order
    _, first = np.unique(qid, return_index=True)
    insertion = qid[np.sort(first)].tolist()
    rows_by_q: Dict[int, List[str]] = {}
    for i in range(len(r["qid"])):
        q = int(r["qid"][i])
        rows_by_q.setdefault(q, []).append("%d %s %d %d %d %d %s\n" % (
            q, ctg_id, r["pid"][i], r["phase"][i], r["n0"][i], r["n1"][i], rid_map[q]))
    os.makedirs(os.path.dirname(phased_read_fn) or ".", exist_ok=True)
    with open(phased_read_fn, "w") as f:
        for q in py2emu.py27_int_dict_order(insertion):
            f.write("".join(rows_by_q.get(q, [])))
    return r


def run_phasing_stages(records, ctg_id: str, ref_seq: str, out_dir: str) -> Dict[str, str]:
    """All four stages with the file layout of phasing.py:501-503,520,534,543."""
    base = os.path.join(out_dir, ctg_id)
    paths = dict(variant_map=os.path.join(base, "het_call", "variant_map"),
                 variant_pos=os.path.join(base, "het_call", "variant_pos"),
                 q_id_map=os.path.join(base, "het_call", "q_id_map"),
                 atable=os.path.join(base, "g_atable", "atable"),
                 phased_variants=os.path.join(base, "get_phased_blocks", "phased_variants"),
                 phased_reads=os.path.join(base, "phased_reads"))
    make_het_call_files(records, ref_seq, paths["variant_map"], paths["variant_pos"], paths["q_id_map"])
    generate_association_table_files(paths["variant_map"], paths["atable"])
    get_phased_blocks_files(paths["variant_map"], paths["atable"], paths["phased_variants"])
    get_phased_reads_files(paths["variant_map"], paths["q_id_map"], paths["phased_variants"], ctg_id,
                           paths["phased_reads"])
    return paths
