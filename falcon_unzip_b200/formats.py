"""Text renderings of the device results: byte-identical to the files the reference
writes (reference falcon_unzip/phasing.py:124-134,199,418-421,
478-480) including the Python-2 semantics of Appendix B (row order of phased_reads B.3,
``str(float)`` B.5, single-space separators B.6)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np

from ._lib import lib

BASES = "ACGT"


def py27_float_str(v: float) -> str:
    """Python 2 ``str(float)`` = '%.12g', plus '.0' when no '.', 'e', inf or nan appears."""
    s = "%.12g" % v
    if "." not in s and "e" not in s and "n" not in s:
        s += ".0"
    return s


def py27_int_dict_order(keys) -> np.ndarray:
    """Iteration order of a CPython-2.7 dict whose int keys were inserted in this order."""
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    uniq_n = len(np.unique(keys))
    out = np.empty(max(uniq_n, 1), dtype=np.int64)
    rc = lib().fuz_host_py27_int_dict_order(keys.ctypes.data, len(keys), out.ctypes.data)
    if rc:
        raise RuntimeError("fuz_host_py27_int_dict_order failed: %d" % rc)
    return out[:uniq_n]


def contig_slices(res, n_ctg: int) -> Dict[str, np.ndarray]:
    """Row ranges of every contig inside the batch-wide row arrays."""
    site_off = np.searchsorted(res.site_ctg, np.arange(n_ctg + 1), side="left")
    vm_off = np.searchsorted(res.vm_site, site_off, side="left")
    at_off = np.searchsorted(res.at_s1, site_off, side="left")
    pr_off = np.searchsorted(res.pr_ctg, np.arange(n_ctg + 1), side="left")
    return dict(site=site_off, vmap=vm_off, atable=at_off, reads=pr_off)


def _c32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _ref_bytes(ref_seq) -> bytes:
    return ref_seq if isinstance(ref_seq, bytes) else ref_seq.encode("latin-1")


def variant_pos_bytes(res, s0: int, s1: int, ref_seq) -> bytes:
    """het_call/variant_pos: ``pos ref total b0 c0 b1 c1 b2 c2 b3 c3`` (phasing.py:116-124), bases by descending (count, base);
    rows formatted in libfuz."""
    site_pos, site_cnt = _c32(res.site_pos), _c32(res.site_cnt)
    ref = _ref_bytes(ref_seq)
    cap = 80 * (s1 - s0) + 16
    buf = C.create_string_buffer(cap)
    n = lib().fuz_host_format_variant_pos(site_pos.ctypes.data, site_cnt.ctypes.data, s0, s1, ref, len(ref), buf, cap)
    if n == -2:
        raise IndexError("string index out of range (ref_seq shorter than a het position; phasing.py:123)")
    if n < 0:
        raise RuntimeError("fuz_host_format_variant_pos failed")
    return buf.raw[:n]


def variant_pos_text(res, s0: int, s1: int, ref_seq: str) -> str:
    return variant_pos_bytes(res, s0, s1, ref_seq).decode("latin-1")


def variant_map_bytes(res, s0: int, v0: int, v1: int, ref_seq) -> bytes:
    """het_call/variant_map: ``pos ref allele q_id`` (phasing.py:126,128); rows formatted in libfuz."""
    site_pos, vm_site, vm_qid = _c32(res.site_pos), _c32(res.vm_site), _c32(res.vm_qid)
    vm_base = np.ascontiguousarray(res.vm_base, dtype=np.uint8)
    ref = _ref_bytes(ref_seq)
    cap = 32 * (v1 - v0) + 16
    buf = C.create_string_buffer(cap)
    n = lib().fuz_host_format_variant_map(site_pos.ctypes.data, vm_site.ctypes.data, vm_base.ctypes.data, vm_qid.ctypes.data,
                                          v0, v1, ref, len(ref), buf, cap)
    if n == -2:
        raise IndexError("string index out of range (ref_seq shorter than a het position; phasing.py:123)")
    if n < 0:
        raise RuntimeError("fuz_host_format_variant_map failed")
    return buf.raw[:n]


def variant_map_text(res, s0: int, v0: int, v1: int, ref_seq: str) -> str:
    return variant_map_bytes(res, s0, v0, v1, ref_seq).decode("latin-1")


def _name_rows(names):
    """names as fixed-width NUL-padded rows (numpy "S"): (array, width), or None for a list / dict of str."""
    if isinstance(names, np.ndarray) and names.dtype.kind == "S":
        a = np.ascontiguousarray(names)
        return a, a.dtype.itemsize
    return None


def q_id_map_bytes(names) -> bytes:
    """het_call/q_id_map (phasing.py:132-134); dense int keys iterate ascending.  names: list of str, or the fixed-width
    QNAME rows of a device batch (formatted in libfuz, no str objects)."""
    rows = _name_rows(names)
    if rows is None:
        return "".join("%d %s\n" % (i, n) for i, n in enumerate(names)).encode("latin-1")
    a, width = rows
    cap = len(a) * (width + 13) + 16
    buf = C.create_string_buffer(cap)
    n = lib().fuz_host_format_q_id_map_rows(a.ctypes.data if len(a) else None, width, len(a), buf, cap)
    if n < 0:
        raise RuntimeError("fuz_host_format_q_id_map_rows failed")
    return buf.raw[:n]


def q_id_map_text(names) -> str:
    return q_id_map_bytes(names).decode("latin-1")


def atable_bytes(res, a0: int, a1: int) -> bytes:
    """g_atable/atable: ``pos1 b11 b12 pos2 b21 b22 c11 c12 c21 c22`` (phasing.py:199); rows formatted in libfuz."""
    site_pos, at_s1, at_s2, at_ct = _c32(res.site_pos), _c32(res.at_s1), _c32(res.at_s2), _c32(res.at_ct)
    site_al = np.ascontiguousarray(res.site_al, dtype=np.uint8)
    cap = 96 * (a1 - a0) + 16
    buf = C.create_string_buffer(cap)
    n = lib().fuz_host_format_atable(site_pos.ctypes.data, site_al.ctypes.data, at_s1.ctypes.data, at_s2.ctypes.data,
                                     at_ct.ctypes.data, a0, a1, buf, cap)
    if n < 0:
        raise RuntimeError("fuz_host_format_atable failed")
    return buf.raw[:n]


def atable_text(res, a0: int, a1: int) -> str:
    return atable_bytes(res, a0, a1).decode("ascii")


def phased_variants_bytes(res, s0: int, s1: int, ref_seq) -> bytes:
    """get_phased_blocks/phased_variants: P and V rows (phasing.py:411-421), formatted in libfuz."""
    ref = _ref_bytes(ref_seq)
    cap = 224 * (s1 - s0) + 80
    buf = C.create_string_buffer(cap)
    u8 = lambda a: np.ascontiguousarray(a, dtype=np.uint8)                     # noqa: E731
    cols = [_c32(res.site_pos), u8(res.site_al), _c32(res.ph_block), u8(res.ph_state), _c32(res.ph_lext), _c32(res.ph_rext),
            _c32(res.ph_lscore), _c32(res.ph_rscore)]
    n = lib().fuz_host_format_phased_variants(*[c.ctypes.data for c in cols], s0, s1, ref, len(ref), buf, cap)
    if n == -2:
        raise IndexError("string index out of range (ref_seq shorter than a phased position; phasing.py:418)")
    if n < 0:
        raise RuntimeError("fuz_host_format_phased_variants failed")
    return buf.raw[:n]


def phased_variants_text(res, s0: int, s1: int, ref_base) -> str:
    """get_phased_blocks/phased_variants: P and V rows (phasing.py:411-421).
    ref_base: str indexed by pos-1, or a dict pos -> base (file-level stage)."""
    blk = res.ph_block[s0:s1]
    out: List[str] = []
    n_blocks = int(blk.max()) if len(blk) else 0
    order = np.argsort(blk, kind="stable")
    bounds = np.searchsorted(blk[order], np.arange(1, n_blocks + 2), side="left")
    pos = res.site_pos[s0:s1]
    for pid in range(1, n_blocks + 1):
        idx = order[bounds[pid - 1]:bounds[pid]]
        if len(idx) == 0:
            continue
        ps = pos[idx]
        mn, mx = int(ps.min()), int(ps.max())
        out.append("P %d %d %d %d %d %s\n" % (pid, mn, mx, mx - mn, len(idx),
                                                py27_float_str(1.0 * (mx - mn) / len(idx))))
        st = res.ph_state[s0:s1][idx].tolist()
        al = res.site_al[s0:s1][idx].tolist()
        for p, s, a, le, re_, ls, rs in zip(ps.tolist(), st, al, res.ph_lext[s0:s1][idx].tolist(),
                                            res.ph_rext[s0:s1][idx].tolist(),
                                            res.ph_lscore[s0:s1][idx].tolist(),
                                            res.ph_rscore[s0:s1][idx].tolist()):
            rb = ref_base[p] if isinstance(ref_base, dict) else ref_base[p - 1]
            out.append("V %d %d %d_%s_%s %d_%s_%s %d %d %d %d\n" % (
                pid, p, p, rb, BASES[a[s]], p, rb, BASES[a[1 - s]], le, re_, ls, rs))
    return "".join(out)


def phased_reads_bytes(res, r0: int, r1: int, v0: int, v1: int, ctg_id: str, names) -> bytes:
    """phased_reads: ``q_id ctg block phase n0 n1 qname`` (phasing.py:478,480), reads in the
    iteration order of the reference's ``read_to_variants`` dict (Python-2 int dict, keys
    inserted in order of first appearance in variant_map).  Rows formatted in libfuz.
    names: list of QNAMEs indexed by q_id, or a dict q_id -> QNAME (file-level stage)."""
    vq = _c32(res.vm_qid[v0:v1])
    if len(vq) == 0:
        return b""
    cols = [_c32(getattr(res, k)[r0:r1]) for k in ("pr_qid", "pr_block", "pr_phase", "pr_n0", "pr_n1")]
    rows = _name_rows(names)
    if rows is not None:                                   # QNAME rows of a device batch: no str objects
        a, width = rows
        args = (vq.ctypes.data, len(vq), *[c.ctypes.data for c in cols], r1 - r0, ctg_id.encode("latin-1"),
                a.ctypes.data if len(a) else None, width, len(a))
        cap = 72 * (r1 - r0) + (len(ctg_id) + width) * (r1 - r0) + 16
        buf = C.create_string_buffer(cap)
        n = lib().fuz_host_format_phased_reads_rows(*args, buf, cap)
        if n < 0:
            raise IndexError("phased_reads row with a q_id that has no QNAME")
        return buf.raw[:n]
    if isinstance(names, dict):
        n_names = max(names) + 1 if names else 0
        get = names.get
        name_list = [get(i, "") for i in range(n_names)]
        missing = [int(q) for q in np.unique(res.pr_qid[r0:r1]).tolist() if q not in names]
        if missing:
            raise KeyError(missing[0])                     # rid_map[q_id] of the reference (phasing.py:478)
    else:
        name_list = names
    enc = [n.encode("latin-1") for n in name_list]
    name_off = np.concatenate([[0], np.cumsum([len(e) for e in enc])]).astype(np.int64)
    blob = b"".join(enc) + b"\0"
    args = (vq.ctypes.data, len(vq), *[c.ctypes.data for c in cols], r1 - r0, ctg_id.encode("latin-1"), blob, name_off.ctypes.data,
            len(enc))
    cap = lib().fuz_host_format_phased_reads(*args, None, 0)
    if cap < 0:
        raise IndexError("phased_reads row with a q_id that has no QNAME")
    buf = C.create_string_buffer(int(cap) + 16)
    n = lib().fuz_host_format_phased_reads(*args, buf, int(cap) + 16)
    if n < 0:
        raise RuntimeError("fuz_host_format_phased_reads failed")
    return buf.raw[:n]


def phased_reads_text(res, r0: int, r1: int, v0: int, v1: int, ctg_id: str, names) -> str:
    return phased_reads_bytes(res, r0, r1, v0, v1, ctg_id, names).decode("latin-1")
