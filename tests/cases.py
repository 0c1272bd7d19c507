"""Hand-made inputs for the quirks of SURVEY.md Appendix E (the reference has no tests of its
own).  Each case is a list of SAM-like records for ONE contig; `build` renders them to BAM
records.  Used against the patched reference (CPU, build container only), the C oracle and
the CUDA path."""
import numpy as np

from falcon_unzip_b200 import bam

CTG = "000000F"


def ref_seq(n, seed=3):
    rng = np.random.default_rng(seed)
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


def rec(name, pos, cigar, seq, flag=0):
    return (name, pos, cigar, seq, flag)


def build(records, ctg_len):
    """-> (records bytes, refs)"""
    refs = [(CTG, ctg_len)]
    recs = sorted(records, key=lambda r: r[1])        # coordinate order, stable
    return b"".join(bam.encode_record(0, pos, name, flag, 254, bam.parse_cigar_string(cig), seq)
                    for name, pos, cig, seq, flag in recs), refs


def mutate(seq, edits):
    s = list(seq)
    for p, b in edits.items():
        s[p] = b
    return "".join(s)


def other(base, k=1):
    return "ACGT"[("ACGT".index(base) + k) % 4]


def pile_case(site_bases, n_extra_last=1, length=3000, site=1500, seed=3, name_prefix="r", clip=None):
    """`site_bases`: list of bases the reads show at `site`; every read spans [0, length).
    One more read starting after the site makes the site evaluated (A.1 step 4)."""
    ref = ref_seq(length + 3000, seed)
    recs = []
    for i, b in enumerate(site_bases):
        seq = mutate(ref[:length], {site: b})
        recs.append(rec("%s%d" % (name_prefix, i), 0, "%d=" % length, seq))
    for j in range(n_extra_last):
        recs.append(rec("last%d" % j, site + 10 + j, "%d=" % length, ref[site + 10 + j:site + 10 + j + length]))
    return recs, ref


def all_cases():
    """name -> (records, ref_seq)"""
    out = {}
    L = 3000
    ref = ref_seq(12000, 5)
    # E1 / E2 / E8 / E11: clip boundary (<=), total length, filtered last record, q_ids of filtered reads
    base = [rec("a%d" % i, 0, "%d=" % L, mutate(ref[:L], {1500: "T" if i % 2 else ref[1500]})) for i in range(12)]
    base = [(n, p, c, mutate(s, {1500: other(ref[1500]) if i % 2 else ref[1500]}), f)
            for i, (n, p, c, s, f) in enumerate(base)]
    e1 = list(base)
    e1.append(rec("clip90", 100, "9000S1000=", "A" * 9000 + ref[100:1100]))        # 1 - .9 < .1 -> dropped
    e1.append(rec("clip8999", 100, "8999S1001=", "A" * 8999 + ref[100:1101]))      # kept
    e1.append(rec("short1999", 200, "1999=", ref[200:2199]))                       # dropped
    e1.append(rec("ok2000", 200, "2000=", ref[200:2200]))                          # kept
    e1.append(rec("lastok", 2000, "%d=" % L, ref[2000:2000 + L]))                  # accepted: POS_last = 2000
    e1.append(rec("lastshort", 2500, "1500=", ref[2500:4000]))                     # filtered, last in file (E8)
    out["e1_filters"] = (e1, ref)
    # E3: tie 5/5 at depth 10 -> sorted by letter descending; E5 depth 9 vs 10
    r3, ref3 = pile_case(["A"] * 5 + ["T"] * 5)
    out["e3_tie_AT"] = (r3, ref3)
    r5, ref5 = pile_case(["C"] * 5 + ["G"] * 4)
    out["e5_depth9"] = (r5, ref5)
    # E4: 9/3 of 12 is not het (0.75 / 0.25 not strict); 14/6 of 20 is
    out["e4_9_3"] = pile_case(["A"] * 9 + ["C"] * 3)
    out["e4_14_6"] = pile_case(["G"] * 14 + ["T"] * 6)
    out["e4_three_alleles"] = pile_case(["G"] * 8 + ["T"] * 7 + ["A"] * 6 + ["C"] * 2)
    # E6: N and ambiguity codes do not count towards the depth
    out["e6_A6_C5_N20"] = pile_case(["A"] * 6 + ["C"] * 5 + ["N"] * 20)
    out["e6_A12_N5"] = pile_case(["A"] * 12 + ["N"] * 5)
    out["e6_ambiguity"] = pile_case(["A"] * 6 + ["C"] * 5 + ["M", "R", "W", "S", "Y", "K", "V", "H", "D", "B", "="])
    # E7: het site at / after the start of the last accepted read is never evaluated
    r7, ref7 = pile_case(["A"] * 6 + ["C"] * 6, n_extra_last=0)
    out["e7_no_final_flush"] = (r7, ref7)
    r7b, ref7b = pile_case(["A"] * 6 + ["C"] * 6, n_extra_last=1)
    r7b[-1] = rec("last0", 1500, "3000=", ref7b[1500:4500])                       # starts AT the site
    out["e7_last_starts_at_site"] = (r7b, ref7b)
    # E9: N / H / P operations advance nothing; long deletion; adjacent I and D; soft clips
    refq = ref_seq(30000, 9)
    rq = []
    for i in range(14):
        alt = {2600: other(refq[2600]) if i % 2 else refq[2600], 9100: other(refq[9100], 2) if i % 3 else refq[9100]}
        s = mutate(refq, alt)
        if i % 4 == 0:      # 1200= 50N 1300= : bases after N pile up right after the first run
            rq.append(rec("n%d" % i, 1000, "1200=50N1300=", s[1000:2200] + s[2200:3500]))
        elif i % 4 == 1:    # hard clip + padding + insertion next to deletion
            rq.append(rec("p%d" % i, 1000, "5H700=3I2D500=4P1297=10S", s[1000:1700] + "ACG" + s[1702:2202] + s[2202:3499] + "T" * 10))
        elif i % 4 == 2:    # long deletion: 1500= 6000D 1500=
            rq.append(rec("d%d" % i, 1000, "1500=6000D1500=", s[1000:2500] + s[8500:10000]))
        else:               # M-style with mismatches, leading soft clip
            rq.append(rec("m%d" % i, 1000, "7S2500M", "G" * 7 + s[1000:3500]))
    for i in range(12):
        alt = {9100: other(refq[9100], 2) if i % 2 else refq[9100]}
        rq.append(rec("w%d" % i, 8000, "2500=", mutate(refq, alt)[8000:10500]))
    rq.append(rec("tail", 12000, "2500=", refq[12000:14500]))
    out["e9_ops"] = (rq, refq)
    # E10: the same QNAME on two overlapping records
    r10, ref10 = pile_case(["A"] * 7 + ["G"] * 7)
    r10 = [(("dup" if n in ("r0", "r1", "r8") else n), p, c, s, f) for n, p, c, s, f in r10]
    out["e10_dup_qname"] = (r10, ref10)
    return out


def window_case(gap):
    """E12: two het sites `gap` bp apart, covered by reads long enough to link them."""
    n = gap + 6000
    ref = ref_seq(n + 4000, 21)
    s1, s2 = 2000, 2000 + gap
    recs = []
    for i in range(12):
        a = other(ref[s1]) if i % 2 else ref[s1]
        b = other(ref[s2]) if i % 2 else ref[s2]
        recs.append(rec("L%d" % i, 0, "%d=" % n, mutate(ref[:n], {s1: a, s2: b})))
    recs.append(rec("last", s2 + 50, "3000=", ref[s2 + 50:s2 + 3050]))
    return recs, ref


def random_vmap(rng, n_sites, depth, n_reads, dup_rate=0.1, span=12):
    """A variant_map in the reference-produced format (for stage-level fuzzing):
    -> (pos, ref letters, rows[(pos, ref, allele, qid)])."""
    pos = np.sort(rng.choice(np.arange(1, 40 * n_sites + 50), size=n_sites, replace=False))
    rows, refs = [], []
    hap = rng.integers(0, 2, n_reads)
    for i, p in enumerate(pos):
        a, b = rng.choice(4, size=2, replace=False)
        refb = "ACGT"[a]
        refs.append(refb)
        lo = max(0, int(i * n_reads / n_sites) - span)
        cand = np.arange(lo, min(n_reads, lo + 2 * span + depth))
        qs = rng.choice(cand, size=min(depth, len(cand)), replace=False)
        qs.sort()
        major, minor = [], []
        for q in qs:
            noisy = rng.random() < 0.08
            (major if (hap[q] == 0) != noisy else minor).append(int(q))
            if rng.random() < dup_rate:
                (major if (hap[q] == 0) != noisy else minor).append(int(q))
        if len(major) < 3 or len(minor) < 3:
            major, minor = [int(q) for q in qs[::2]] + [int(qs[0])], [int(q) for q in qs[1::2]] + [int(qs[1])]
        if len(minor) > len(major):
            major, minor, a, b = minor, major, b, a
        for q in major:
            rows.append((int(p), refb, "ACGT"[a], q))
        for q in minor:
            rows.append((int(p), refb, "ACGT"[b], q))
    return pos, refs, rows
