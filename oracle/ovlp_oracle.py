"""TEST INFRASTRUCTURE ONLY (oracle/): restatement of reference falcon_unzip/ovlp_filter_with_phase.py
(filter_stage1 :49-143, filter_stage2 :145-186, filter_stage3 :188-290, main :309-352) on lists of
LA4Falcon -mo lines instead of the LA4Falcon pipe.  Pinned against the reference's own source run
through oracle/ref_exec.py (tests/test_ovlp_oracle_vs_reference.py) and by the golden fixture
tests/golden/ovlp_*.  Only tests/ and scripts/bench_ovlp.py's cpu_baseline leg use it.

Quirks kept (all visible in the output or in the returned lists):
  * every stage drops a line unless BOTH reads are keys of arid2phase, lie on the same contig and are
    not in the same block with different phases (:64-73); the strings are compared as strings;
  * stage 1 evaluates a run of a q when the NEXT passing q shows up, starting with a run of `None`
    that is judged on counts (0, 0) (:75-87): `None` is reported first whenever (0, 0) fails;
  * idt / length tests come AFTER the run bookkeeping (:100-104), so a q whose lines all fail them
    still forms a run with counts (0, 0);
  * stage 3: `if q_s == 0 ... elif q_e == q_l` (:265,:270) while stage 1 counts both ends (:109,:113);
    the candidates of an end are sorted as tuples (-inphase, -len, t_l - (t_e - t_s), token list)
    and printed until the first index >= bestn whose range is > 1000, inclusive (:234-240).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

Phase = Tuple[str, str, str]


def _phase_ok(a2p: Dict[str, Phase], q: str, t: str) -> bool:
    if q not in a2p or t not in a2p:
        return False
    pq, pt = a2p[q], a2p[t]
    if pt[0] != pq[0]:
        return False
    return not (pt[1] == pq[1] and pt[2] != pq[2])


def _verdict(left: int, right: int, max_diff: int, max_ovlp: int, min_ovlp: int) -> bool:
    if abs(left - right) > max_diff:
        return True
    if left > max_ovlp or right > max_ovlp:
        return True
    return left < min_ovlp or right < min_ovlp


def stage1(lines: Iterable[str], a2p: Dict[str, Phase], max_diff: int, max_ovlp: int, min_ovlp: int, min_len: int) -> List[Optional[str]]:
    ignore: List[Optional[str]] = []
    cur, left, right, seen = None, 0, 0, False
    for line in lines:
        l = line.strip().split()
        q, t = l[:2]
        if not _phase_ok(a2p, q, t):
            continue
        seen = True
        if q != cur:
            if _verdict(left, right, max_diff, max_ovlp, min_ovlp):
                ignore.append(cur)
            cur, left, right = q, 0, 0
        if float(l[3]) < 90:
            continue
        q_s, q_e, q_l, t_l = int(l[5]), int(l[6]), int(l[7]), int(l[11])
        int(l[2]); int(l[9]); int(l[10])
        if q_l < min_len or t_l < min_len:
            continue
        left += q_s == 0
        right += q_e == q_l
    if seen and _verdict(left, right, max_diff, max_ovlp, min_ovlp):
        ignore.append(cur)
    return ignore


def stage2(lines: Iterable[str], a2p: Dict[str, Phase], min_len: int, ignore: Set) -> Set[str]:
    contained: Set[str] = set()
    for line in lines:
        l = line.strip().split()
        q, t = l[:2]
        if not _phase_ok(a2p, q, t):
            continue
        if float(l[3]) < 90 or int(l[7]) < min_len or int(l[11]) < min_len:
            continue
        if q in ignore or t in ignore:
            continue
        if l[-1] == "contained":
            contained.add(q)
        if l[-1] == "contains":
            contained.add(t)
    return contained


def _flush(ends: Sequence[list], bestn: int, out: List[List[str]]) -> None:
    for cands in ends:
        for i, (_ip, _sc, m_range, l) in enumerate(sorted(cands)):
            out.append(l)
            if i >= bestn and m_range > 1000:
                break


def stage3(lines: Iterable[str], a2p: Dict[str, Phase], min_len: int, ignore: Set, contained: Set, bestn: int) -> List[List[str]]:
    out: List[List[str]] = []
    cur, five, three = None, [], []
    for line in lines:
        l = line.strip().split()
        q, t = l[:2]
        if not _phase_ok(a2p, q, t):
            continue
        if cur is None:
            cur = q
        elif q != cur:
            _flush((five, three), bestn, out)
            cur, five, three = q, [], []
        if q in contained or t in contained or q in ignore or t in ignore:
            continue
        ovl = -int(l[2])
        if float(l[3]) < 90:
            continue
        q_s, q_e, q_l = int(l[5]), int(l[6]), int(l[7])
        t_s, t_e, t_l = int(l[9]), int(l[10]), int(l[11])
        if q_l < min_len or t_l < min_len:
            continue
        if q_s == 0 or q_e == q_l:
            pq, pt = a2p.get(cur, "NA"), a2p.get(t, "NA")
            l.extend([".".join(pq), ".".join(pt)])
            (five if q_s == 0 else three).append((-(1 if pq == pt else 0), -ovl, t_l - (t_e - t_s), l))
    _flush((five, three), bestn, out)
    return out


def run_filter(files: Sequence[Tuple[str, Sequence[str]]], a2p: Dict[str, Phase], max_diff: int, max_cov: int, min_cov: int,
               min_len: int, bestn: int) -> str:
    """main() (:309-352): what the reference prints, as one string."""
    ignore_all: List = []
    for _fn, lines in files:
        ignore_all.extend(stage1(lines, a2p, max_diff, max_cov, min_cov, min_len))
    ignore = set(ignore_all)
    contained: Set[str] = set()
    for _fn, lines in files:
        contained.update(stage2(lines, a2p, min_len, ignore))
    out = []
    for _fn, lines in files:
        out.extend(" ".join(l) for l in stage3(lines, a2p, min_len, ignore, contained, bestn))
    return "".join(x + "\n" for x in out)


def load_rid_phase_map(path: str) -> Dict[str, Phase]:
    a2p: Dict[str, Phase] = {}
    with open(path) as f:
        for row in f:
            row = row.strip().split()
            a2p[row[0]] = (row[1], row[2], row[3])
    return a2p
