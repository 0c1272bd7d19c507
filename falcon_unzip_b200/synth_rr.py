"""Synthetic inputs of the raw-read -> haplotig tracking (BASELINE.json config 4): LA4Falcon -m
text lines per LAS file, read_to_contig_map, rawread_ids and all_phased_reads, in the formats
of SURVEY.md Appendix D (reference falcon_unzip/rr_hctg_track.py:15-23,38-44,72-85).

Reads are placed on primary contigs; part of every primary contig is covered by a haplotig,
so that reads anchored there map to both the primary contig and the haplotig (the score ties
of SURVEY.md B.4) and carry a phase.
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List

import numpy as np


@dataclasses.dataclass
class RRSet:
    las_lines: Dict[str, List[str]]      # LAS file name -> LA4Falcon -m lines
    read_to_contig_map: List[str]
    rawread_ids: str                     # file content (one original read name per line)
    phased_reads: List[str]
    n_reads: int


def generate_rr(n_reads: int = 3000, n_ctg: int = 4, ctg_len: int = 150_000, mean_len: int = 9000, n_files: int = 3,
                seed: int = 4, min_ovl: int = 1000, map_frac: float = 0.5, short_frac: float = 0.08) -> RRSet:
    rng = np.random.default_rng(seed)
    lens = np.clip(rng.normal(mean_len, 0.25 * mean_len, n_reads), 800, ctg_len).astype(np.int64)
    short = rng.random(n_reads) < short_frac
    lens[short] = rng.integers(800, 2499, int(short.sum()))            # below --min-len 2500
    ctg = rng.integers(0, n_ctg, n_reads)
    start = (rng.random(n_reads) * (ctg_len - lens)).astype(np.int64)
    end = start + lens
    names = ["m%08d/%d/0_%d" % (i, i, lens[i]) for i in range(n_reads)]
    # haplotig regions: the middle third of every primary contig, two haplotigs per contig
    h_lo, h_mid, h_hi = ctg_len // 3, ctg_len // 2, 2 * ctg_len // 3
    pname = ["%06dF" % c for c in range(n_ctg)]
    in_h1 = (start < h_mid) & (end > h_lo)
    in_h2 = (start < h_hi) & (end > h_mid)
    # read_to_contig_map: a subset of reads are contig edges; reads in a haplotig region are
    # listed for the primary contig and / or the haplotig (row order varies)
    r2c = []
    mapped = rng.random(n_reads) < map_frac
    pid = 0
    for r in np.flatnonzero(mapped):
        rows = []
        hap = rng.integers(0, 2)
        if in_h1[r] or in_h2[r]:
            hname = "%s_%03d" % (pname[ctg[r]], 1 if in_h1[r] else 2)
            kind = rng.integers(0, 3)
            if kind == 0:
                rows = [pname[ctg[r]], hname]
            elif kind == 1:
                rows = [hname, pname[ctg[r]]]
            else:
                rows = [hname] if hap else [pname[ctg[r]]]
        else:
            rows = [pname[ctg[r]]]
        for cn in rows:
            r2c.append("%09d %09d %s %s" % (pid, r, names[r], cn))
            pid += 1
    # phases: reads in haplotig regions, block = 1 / 2 (sometimes -1), phase 0 / 1
    phased = []
    for r in range(n_reads):
        if (in_h1[r] or in_h2[r]) and rng.random() < 0.8:
            block = (1 if in_h1[r] else 2) if rng.random() < 0.95 else -1
            phase = int(rng.integers(0, 2))
            phased.append("%d %s %d %d %d %d %s" % (r, pname[ctg[r]], block, phase, 5 * (1 - phase), 5 * phase, names[r]))
    # overlaps: LA4Falcon prints, per A-read in ascending id, every B-read overlapping it
    order = np.lexsort((start, ctg))
    partners: List[List[int]] = [[] for _ in range(n_reads)]
    for ii, a in enumerate(order):
        for b in order[ii + 1:]:
            if ctg[b] != ctg[a] or start[b] >= end[a] - min_ovl:
                break
            if min(end[a], end[b]) - start[b] >= min_ovl:
                partners[a].append(int(b))
                partners[b].append(int(a))
    per_file = (n_reads + n_files - 1) // n_files
    las: Dict[str, List[str]] = {}
    for f in range(n_files):
        lines = []
        for a in range(f * per_file, min(n_reads, (f + 1) * per_file)):
            for b in sorted(partners[a]):
                lo, hi = max(start[a], start[b]), min(end[a], end[b])
                ovl = int(hi - lo)
                if rng.random() < 0.05:
                    ovl = int(rng.choice([5000, 7000]))            # equal lengths: heap ties
                lines.append("%09d %09d %d %.2f 0 %d %d %d 0 %d %d %d overlap" % (
                    a, b, -ovl, 99.0 - rng.random(), lo - start[a], hi - start[a], lens[a], lo - start[b], hi - start[b],
                    lens[b]))
                if rng.random() < 0.02:                            # the same pair twice
                    lines.append(lines[-1])
        las["0-rawreads/m_%05d/raw_reads.%d.las" % (f + 1, f + 1)] = lines
    return RRSet(las, r2c, "\n".join(names) + "\n", phased, n_reads)


@dataclasses.dataclass
class OvlpSet:
    las_lines: Dict[str, List[str]]      # LAS file name -> LA4Falcon -mo lines (preads)
    rid_phase_rows: List[str]            # rid_to_phase.all rows: "%09d ctg block phase"
    n_reads: int


def generate_ovlp(n_reads: int = 2500, n_ctg: int = 3, ctg_len: int = 120_000, mean_len: int = 8000, n_files: int = 3,
                  seed: int = 5, min_ovl: int = 800, unmapped_frac: float = 0.06, short_frac: float = 0.05,
                  low_idt_frac: float = 0.04, dup_frac: float = 0.03) -> OvlpSet:
    """Inputs of the overlap filter with phase (ovlp_filter_with_phase.py): preads laid out on contigs with
    a phased region per contig (two blocks, two phases), LA4Falcon -mo lines with overlap / contains /
    contained tags and 5' / 3' end structure (q_s == 0, q_e == q_l), reads missing from the phase map,
    short reads, low-identity lines, repeated pairs with equal length (sort-key ties)."""
    rng = np.random.default_rng(seed)
    lens = np.clip(rng.normal(mean_len, 0.3 * mean_len, n_reads), 600, ctg_len).astype(np.int64)
    short = rng.random(n_reads) < short_frac
    lens[short] = rng.integers(600, 2499, int(short.sum()))
    ctg = rng.integers(0, n_ctg, n_reads)
    start = (rng.random(n_reads) * (ctg_len - lens)).astype(np.int64)
    end = start + lens
    hap = rng.integers(0, 2, n_reads)
    lo, mid, hi = ctg_len // 4, ctg_len // 2, 3 * ctg_len // 4
    rows = []
    for r in range(n_reads):
        if rng.random() < unmapped_frac:
            continue                                     # not in rid_to_phase.all
        centre = (start[r] + end[r]) // 2
        if lo <= centre < hi and rng.random() < 0.85:
            block, phase = (0 if centre < mid else 1), int(hap[r])
        else:
            block, phase = -1, 0
        rows.append("%09d %06dF %d %d" % (r, ctg[r], block, phase))
    order = np.lexsort((start, ctg))
    partners: List[List[int]] = [[] for _ in range(n_reads)]
    for ii, a in enumerate(order):
        for b in order[ii + 1:]:
            if ctg[b] != ctg[a] or start[b] >= end[a] - min_ovl:
                break
            partners[a].append(int(b))
            partners[b].append(int(a))
    per_file = (n_reads + n_files - 1) // n_files
    las: Dict[str, List[str]] = {}
    for f in range(n_files):
        lines = []
        for a in range(f * per_file, min(n_reads, (f + 1) * per_file)):
            for b in sorted(partners[a]):
                o_lo, o_hi = max(start[a], start[b]), min(end[a], end[b])
                ovl = int(o_hi - o_lo)
                qs, qe, ts, te = o_lo - start[a], o_hi - start[a], o_lo - start[b], o_hi - start[b]
                if qs == 0 and qe == lens[a]:
                    tag = "contained"
                elif ts == 0 and te == lens[b]:
                    tag = "contains"
                else:
                    tag = "overlap"
                if rng.random() < 0.01:
                    tag = "none"
                if rng.random() < 0.05:
                    ovl = int(rng.choice([4000, 6000]))    # equal lengths: the sort falls to the range and the t id
                idt = 99.9 - 3.0 * rng.random() if rng.random() > low_idt_frac else 80.0 + 9.9 * rng.random()
                strand = int(rng.integers(0, 2))
                lines.append("%09d %09d %d %.2f 0 %d %d %d %d %d %d %d %s" % (a, b, -ovl, idt, qs, qe, lens[a], strand, ts, te, lens[b], tag))
                if rng.random() < dup_frac:               # the same pair again: identical line, or a different identity
                    lines.append(lines[-1] if rng.random() < 0.5 else
                                 "%09d %09d %d %.2f 0 %d %d %d %d %d %d %d %s" % (a, b, -ovl, idt - 0.5, qs, qe, lens[a], strand, ts, te, lens[b], tag))
        las["1-preads_ovl/m_%05d/preads.%d.las" % (f + 1, f + 1)] = lines
    return OvlpSet(las, rows, n_reads)


# --------------------------------------------------------------------------- config 4 at size (arrays, vectorised)
@dataclasses.dataclass
class RRArrays:
    """The tracking inputs as the integer columns fuz_rr_track takes (what parsing the LA4Falcon -m text gives) plus the
    id tables of rr_hctg_track.py:70-85 in array form."""
    q: np.ndarray
    t: np.ndarray
    len: np.ndarray
    tlen: np.ndarray
    file: np.ndarray
    n_reads: int
    n_files: int
    in_map: np.ndarray
    rc_off: np.ndarray
    rc_ctg: np.ndarray
    ph_ctg: np.ndarray
    ph_block: np.ndarray
    ph_phase: np.ndarray
    n_ctg_names: int
    read_len: np.ndarray
    q_s: np.ndarray
    q_e: np.ndarray
    t_s: np.ndarray
    t_e: np.ndarray


def generate_rr_arrays(total_len: int = 100_000_000, n_ctg: int = 50, coverage: float = 60.0, mean_len: int = 10_000,
                       n_files: int = 48, seed: int = 20240605, min_ovl: int = 1000, map_frac: float = 0.5,
                       short_frac: float = 0.08) -> RRArrays:
    """BASELINE.json configs[3]: raw reads at `coverage` over `total_len` of primary contigs + haplotigs (the middle third of
    every primary contig carries two haplotigs), every pair of reads of a contig that overlaps by >= min_ovl gives the two
    LA4Falcon lines (A -> B and B -> A); lines are ordered like LA4Falcon prints them (per LAS file, A ascending, B
    ascending); LAS file = A // reads-per-file.  ~108 overlaps per read at 60x / 10 kb."""
    rng = np.random.default_rng(seed)
    ctg_len = total_len // n_ctg
    n_reads = int(total_len * coverage / mean_len)
    lens = np.clip(rng.normal(mean_len, 0.25 * mean_len, n_reads), 800, ctg_len).astype(np.int64)
    short = rng.random(n_reads) < short_frac
    lens[short] = rng.integers(800, 2499, int(short.sum()))
    ctg = rng.integers(0, n_ctg, n_reads)
    start = (rng.random(n_reads) * (ctg_len - lens)).astype(np.int64)
    end = start + lens
    # pairs: reads sorted by (contig, start); partners of a = following reads of the contig starting before end[a] - min_ovl
    key = ctg * (4 * ctg_len) + start
    order = np.argsort(key, kind="stable")
    ks, ke = key[order], (ctg * (4 * ctg_len) + end - min_ovl)[order]
    hi = np.searchsorted(ks, ke, side="left")
    cnt = np.maximum(hi - np.arange(n_reads) - 1, 0)
    ia = np.repeat(np.arange(n_reads), cnt)
    ib = ia + 1 + (np.arange(int(cnt.sum())) - np.repeat(np.cumsum(cnt) - cnt, cnt))
    a, b = order[ia], order[ib]
    ok = np.minimum(end[a], end[b]) - start[b] >= min_ovl
    a, b = a[ok], b[ok]
    qa = np.concatenate([a, b]).astype(np.int32)
    tb = np.concatenate([b, a]).astype(np.int32)
    o = np.lexsort((tb, qa))
    qa, tb = qa[o], tb[o]
    lo, hi2 = np.maximum(start[qa], start[tb]), np.minimum(end[qa], end[tb])
    ovl = (hi2 - lo).astype(np.int32)
    tie = rng.random(len(ovl)) < 0.05
    ovl[tie] = rng.choice(np.asarray([5000, 7000], np.int32), int(tie.sum()))      # equal lengths: heap ties
    per_file = (n_reads + n_files - 1) // n_files
    file = (qa // per_file).astype(np.int32)
    # id tables: half of the reads are contig edges; reads in a haplotig region map to primary and / or haplotig
    h_lo, h_mid, h_hi = ctg_len // 3, ctg_len // 2, 2 * ctg_len // 3
    in_h1 = (start < h_mid) & (end > h_lo)
    in_h2 = (start < h_hi) & (end > h_mid)
    in_h = in_h1 | in_h2
    mapped = rng.random(n_reads) < map_frac
    kind = rng.integers(0, 3, n_reads)
    hap = rng.integers(0, 2, n_reads)
    prim = ctg.astype(np.int32)                                   # contig name ids: primaries 0 .. n_ctg-1, haplotigs behind
    hname = (n_ctg + 2 * ctg + np.where(in_h1, 0, 1)).astype(np.int32)
    first = np.where(in_h & ((kind == 1) | ((kind == 2) & (hap == 1))), hname, prim)
    second = np.where(in_h & (kind == 0), hname, np.where(in_h & (kind == 1), prim, -1))
    rc_cnt = np.where(mapped, 1 + (second >= 0), 0)
    rc_off = np.concatenate([[0], np.cumsum(rc_cnt)]).astype(np.int32)
    rc_ctg = np.zeros(max(int(rc_off[-1]), 1), np.int32)
    m = np.flatnonzero(mapped)
    rc_ctg[rc_off[m]] = first[m]
    m2 = m[second[m] >= 0]
    rc_ctg[rc_off[m2] + 1] = second[m2]
    phased = in_h & (rng.random(n_reads) < 0.8)
    ph_ctg = np.where(phased, prim, -1).astype(np.int32)
    ph_block = np.where(rng.random(n_reads) < 0.95, np.where(in_h1, 1, 2), -1).astype(np.int32)
    ph_phase = rng.integers(0, 2, n_reads).astype(np.int32)
    return RRArrays(qa, tb, ovl, lens[tb].astype(np.int32), file, n_reads, n_files, mapped.astype(np.uint8), rc_off, rc_ctg,
                    ph_ctg, np.where(phased, ph_block, 0).astype(np.int32), np.where(phased, ph_phase, 0).astype(np.int32),
                    3 * n_ctg, lens.astype(np.int32), (lo - start[qa]).astype(np.int32), (hi2 - start[qa]).astype(np.int32),
                    (lo - start[tb]).astype(np.int32), (hi2 - start[tb]).astype(np.int32))


def rr_text_lines(rr: RRArrays, i0: int, i1: int):
    """LA4Falcon -m lines [i0, i1) of an RRArrays (for the CPU oracle leg and the text leg of the bench)."""
    return ["%09d %09d %d 99.00 0 %d %d %d 0 %d %d %d overlap" % (
        rr.q[i], rr.t[i], -rr.len[i], rr.q_s[i], rr.q_e[i], rr.read_len[rr.q[i]], rr.t_s[i], rr.t_e[i], rr.tlen[i])
        for i in range(i0, i1)]
