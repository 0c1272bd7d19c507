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
