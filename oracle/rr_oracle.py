"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the raw-read -> haplotig tracking of
reference falcon_unzip/rr_hctg_track.py with the CPython-2 container semantics of SURVEY.md
Appendix B.4 (dict / set iteration order, heapq array order).  Restates

    get_rid_to_ctg     rr_hctg_track.py:15-23
    tr_stage1          rr_hctg_track.py:31-65
    run_track_reads    rr_hctg_track.py:67-138  (tables :70-85, merge :97-105, vote :112-138)

Pinned against the reference's own source (oracle/ref_exec.load_rr_hctg_track, patched for
Python 3 + py2 order emulators) by tests/test_rr_oracle_vs_reference.py.
"""
from __future__ import annotations

from heapq import heappush, heappushpop
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

from .py2emu import Py27StrSet, py27_str_dict_order


def get_rid_to_ctg(lines: Iterable[str]) -> Dict[str, Py27StrSet]:
    """rid (str) -> set of contigs; the set iterates in CPython-2 order (:15-23)."""
    rid_to_ctg: Dict[str, Py27StrSet] = {}
    for row in lines:
        row = row.strip().split()
        if not row:
            continue
        _pid, rid, _oid, ctg = row
        rid_to_ctg.setdefault(rid, Py27StrSet()).add(ctg)
    return rid_to_ctg


def phase_table(phased_read_lines: Iterable[str], rawread_ids_text: str) -> List[Optional[Tuple[str, int, int]]]:
    """rid -> (ctg, block, phase) or None (:72-85); later rows of the same read overwrite."""
    oid_to_phase = {}
    for row in phased_read_lines:
        row = row.strip().split()
        if not row:
            continue
        ctg_id, block, phase = row[1:4]
        oid_to_phase[row[6]] = (ctg_id, int(block), int(phase))
    rid_to_oid = rawread_ids_text.split("\n")
    return [oid_to_phase.get(oid) for oid in rid_to_oid]


def tr_stage1(lines: Iterable[str], min_len: int, bestn: int, rid_to_ctg, rid_to_phase) -> Dict[str, list]:
    """Per target read the bestn largest (overlap_len, q_id) of the kept overlaps, as a heapq
    array (:31-65).  The dict keeps first-kept-appearance order of the targets."""
    rtn: Dict[str, list] = {}
    for l in lines:
        l = l.strip().split()
        q_id, t_id = l[:2]
        overlap_len = -int(l[2])
        t_l = int(l[11])
        if t_l < min_len:
            continue
        if q_id not in rid_to_ctg:
            continue
        t_phase = rid_to_phase[int(t_id)]
        if t_phase is not None:
            ctg_id, block, phase = t_phase
            if block != -1:
                q_phase = rid_to_phase[int(q_id)]
                if q_phase is not None and q_phase[0] == ctg_id and q_phase[1] == block and q_phase[2] != phase:
                    continue
        h = rtn.setdefault(t_id, [])
        if len(h) < bestn:
            heappush(h, (overlap_len, q_id))
        else:
            heappushpop(h, (overlap_len, q_id))
    return rtn


def run_track_reads(las_lines: Dict[str, Sequence[str]], phased_read_lines: Iterable[str],
                    read_to_contig_map_lines: Iterable[str], rawread_ids_text: str, min_len: int = 2500,
                    bestn: int = 40, file_order: Sequence[str] = None) -> str:
    """-> text of rawread_to_contigs (:67-138).  las_lines: LAS file name -> LA4Falcon -m
    lines; files are processed in the order of file_order (the reference walks file_list as given: imap keeps the
    order of its inputs, :88-98), default: sorted names (upstream: glob order, B.4)."""
    rid_to_ctg = get_rid_to_ctg(read_to_contig_map_lines)
    rid_to_phase = phase_table(phased_read_lines, rawread_ids_text)
    bread_to_areads: Dict[str, list] = {}
    for fn in (file_order if file_order is not None else sorted(las_lines)):
        res = tr_stage1(las_lines[fn], min_len, bestn, rid_to_ctg, rid_to_phase)
        for k in py27_str_dict_order(res):                          # :99
            h = bread_to_areads.setdefault(k, [])
            for item in res[k]:                                     # :101-105, heap array order
                if len(h) < bestn:
                    heappush(h, item)
                else:
                    heappushpop(h, item)
    out = []
    for bread in py27_str_dict_order(bread_to_areads):              # :113
        ctg_score: Dict[str, list] = {}
        for s, rid in bread_to_areads[bread]:
            if rid not in rid_to_ctg:
                continue
            for ctg in rid_to_ctg[rid]:                             # py2 set order
                sc = ctg_score.setdefault(ctg, [0, 0])
                sc[0] += -s
                sc[1] += 1
        items = [(k, ctg_score[k]) for k in py27_str_dict_order(ctg_score)]   # :126
        items.sort(key=lambda k: k[1][0])                           # stable (:127)
        for rank, (ctg, (score, count)) in enumerate(items):
            in_ctg = 1 if bread in rid_to_ctg and ctg in rid_to_ctg[bread] else 0
            out.append("%s %s %d %d %d %d\n" % (bread, ctg, count, rank, score, in_ctg))
    return "".join(out)
