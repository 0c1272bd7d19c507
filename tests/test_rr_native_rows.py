"""The C++ twins of the CPython-2 order emulation used for rawread_to_contigs (fuz_host_py27_str_dict_order,
fuz_host_rr_bread_order, fuz_host_rr_format_rows) against the Python versions of py2compat / rr_hctg_track on
random inputs (no GPU).  The Python versions are pinned by the hand-derived vectors of tests/test_host_formats.py."""
import numpy as np


def test_str_dict_order_matches_python():
    from falcon_unzip_b200 import _lib, py2compat
    rng = np.random.default_rng(1)
    L = _lib.lib()
    for n in (0, 1, 5, 6, 21, 22, 500, 70000):
        keys = ["%09d" % int(x) for x in rng.integers(0, max(n, 1) * 2, n)] if n < 60000 else ["k%d_%s" % (i % 45000, "ab"[i % 2]) for i in range(n)]
        if n in (5, 500):
            keys += ["", "a", "000001F_001", "000001F"]
        blob = "".join(keys).encode("ascii") + b"\0"
        off = np.concatenate([[0], np.cumsum([len(k) for k in keys])]).astype(np.int64)
        out = np.empty(max(len(keys), 1), np.int64)
        m = L.fuz_host_py27_str_dict_order(blob, off.ctypes.data, len(keys), out.ctypes.data)
        assert [keys[i] for i in out[:m].tolist()] == py2compat.str_dict_order(dict.fromkeys(keys))


def test_bread_order_and_rows_match_python():
    from falcon_unzip_b200 import rr_hctg_track as rrm
    rng = np.random.default_rng(2)
    for trial in range(4):
        n_reads, n_files, n_ctg = 5000, 5, 7
        n = 40000
        file_kept = np.sort(rng.integers(0, n_files, n)).astype(np.int32)
        t_kept = rng.integers(0, n_reads, n).astype(np.int32)
        want = rrm._bread_order(t_kept, file_kept)
        got = rrm._bread_order_native(t_kept, file_kept)
        assert ["%09d" % x for x in got.tolist()] == want
        # vote rows per read, contig names with shared prefixes (haplotigs), random membership table
        names = ["%06dF" % c for c in range(n_ctg)] + ["%06dF_%03d" % (c, 1) for c in range(n_ctg)]
        rid_to_ctg, rid_to_phase = {}, [None] * n_reads
        for r in rng.choice(n_reads, 2500, replace=False).tolist():
            s = rrm.OrderedStrSet()
            for c in rng.choice(len(names), int(rng.integers(1, 4)), replace=False).tolist():
                s.add(names[c])
            rid_to_ctg["%09d" % r] = s
        tab = rrm._Tables(rid_to_ctg, rid_to_phase, n_reads)
        for nm in names:                                       # contigs that never occur in rid_to_ctg
            if nm not in tab.ctg_names:
                tab.ctg_names.append(nm)
        cnt = rng.integers(0, 4, n_reads)
        vt_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        vt_ctg = np.concatenate([rng.choice(len(tab.ctg_names), int(k), replace=False) for k in cnt] + [np.zeros(0, int)]).astype(np.int32)
        vt_count = rng.integers(1, 40, len(vt_ctg)).astype(np.int32)
        vt_score = (-rng.integers(1, 6, len(vt_ctg)) * 5000).astype(np.int64)          # few distinct scores: stable-sort ties
        text_py = "".join(rrm._format_bread(b, tab, rid_to_ctg, vt_off, vt_ctg, vt_count, vt_score) for b in want)
        assert rrm._format_rows_native(got, tab, vt_off, vt_ctg, vt_count, vt_score).decode("ascii") == text_py
        assert len(text_py) > 10000
