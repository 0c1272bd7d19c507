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
