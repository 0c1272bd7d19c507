"""The id joins of SURVEY.md 8f-2 (falcon_unzip_b200/readmaps.py) against the reference's own source run
under Python 3 (oracle/ref_exec.py; CPython-2 container orders through oracle/py2emu.py), byte for byte."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from oracle import py2emu, ref_exec

needs_ref = pytest.mark.skipif(not ref_exec.available(), reason="reference tree not present")


def _world(tmp, seed=3, n_raw=800, n_pread=600):
    """read_map_dir with rawread_ids / pread_ids / pread_to_contigs, a phased_reads file and contig edge files."""
    rng = np.random.default_rng(seed)
    rmd = os.path.join(tmp, "read_maps")
    os.makedirs(os.path.join(rmd, "dump_rawread_ids")); os.makedirs(os.path.join(rmd, "dump_pread_ids"))
    oids = ["m%06d/%d/0_%d" % (i, i, 5000 + i) for i in range(n_raw)]
    open(os.path.join(rmd, "dump_rawread_ids", "rawread_ids"), "w").write("\n".join(oids) + "\n")
    p2r = rng.choice(n_raw, n_pread, replace=False)
    fids = ["prolog/%d/0_%d" % (int(r) * 10 + int(rng.integers(0, 10)), 4000) for r in p2r]
    open(os.path.join(rmd, "dump_pread_ids", "pread_ids"), "w").write("\n".join(fids) + "\n")
    ctgs = ["000000F", "000001F", "000001F_001", "000002F"]
    with open(os.path.join(rmd, "pread_to_contigs"), "w") as f:
        for p in range(n_pread):
            for rank in range(int(rng.integers(1, 3))):
                f.write("%09d %s %d %d %d %d\n" % (p, ctgs[int(rng.integers(0, 4))], 5, rank, -100, 1))
    pr = os.path.join(tmp, "phased_reads")
    with open(pr, "w") as f:
        for i in rng.choice(n_raw, n_raw // 2, replace=False):
            f.write("%d 000001F %d %d 3 1 %s\n" % (i, int(rng.integers(-1, 4)), int(rng.integers(0, 2)), oids[int(i)]))
    for name, tigs in (("all_p_ctg_edges", ctgs[:2] + ctgs[3:]), ("all_h_ctg_edges", ["000001F_001", "000001F_002", "000000F_001"])):
        with open(os.path.join(tmp, name), "w") as f:
            for _ in range(700):
                a, b = rng.integers(0, n_pread, 2)
                f.write("%s %09d:%s %09d:%s x 1 2 3\n" % (tigs[int(rng.integers(0, len(tigs)))], a, "BE"[int(rng.integers(0, 2))], b, "E"))
    open(os.path.join(tmp, "all_h_ctg_ids"), "w").write("000001F_001\n000000F_001\n")
    return rmd, pr


@needs_ref
@pytest.mark.parametrize("ctg", ["000001F", "000000F"])
def test_phasing_readmap(tmp_path, ctg):
    from falcon_unzip_b200 import readmaps
    rmd, pr = _world(str(tmp_path))
    a = SimpleNamespace(phased_reads=pr, read_map_dir=rmd, ctg_id=ctg, base_dir=str(tmp_path / "ref"))
    os.makedirs(a.base_dir)
    ref_exec.load_phasing_readmap().get_phasing_readmap(a)
    b = SimpleNamespace(phased_reads=pr, read_map_dir=rmd, ctg_id=ctg, base_dir=str(tmp_path / "got"))
    readmaps.get_phasing_readmap(b)
    want = open(os.path.join(a.base_dir, "rid_to_phase.%s" % ctg)).read()
    assert open(os.path.join(b.base_dir, "rid_to_phase.%s" % ctg)).read() == want and len(want) > 500


@needs_ref
def test_rid_to_phase_all_and_read_to_hctg_map(tmp_path):
    from falcon_unzip_b200 import readmaps
    rmd, pr = _world(str(tmp_path), seed=4)
    files = {}
    for ctg in ("000001F", "000000F", "000002F"):
        readmaps.get_phasing_readmap(SimpleNamespace(phased_reads=pr, read_map_dir=rmd, ctg_id=ctg, base_dir=str(tmp_path)))
        files[ctg] = str(tmp_path / ("rid_to_phase.%s" % ctg))
    ref_exec.get_rid_to_phase_all_source().get_rid_to_phase_all(SimpleNamespace(rid_to_phase_all=str(tmp_path / "all.ref"), inputs=files))
    readmaps.get_rid_to_phase_all(SimpleNamespace(rid_to_phase_all=str(tmp_path / "all.got"), inputs=files))
    assert open(tmp_path / "all.got").read() == open(tmp_path / "all.ref").read()
    task = dict(rawread_id_file=os.path.join(rmd, "dump_rawread_ids", "rawread_ids"),
                pread_id_file=os.path.join(rmd, "dump_pread_ids", "pread_ids"), h_ctg_edges=str(tmp_path / "all_h_ctg_edges"),
                p_ctg_edges=str(tmp_path / "all_p_ctg_edges"), h_ctg_ids=str(tmp_path / "all_h_ctg_ids"))
    ref_exec.load_get_read_hctg_map().generate_read_to_hctg_map(SimpleNamespace(read_to_contig_map=str(tmp_path / "r2c.ref"), **task))
    readmaps.generate_read_to_hctg_map(SimpleNamespace(read_to_contig_map=str(tmp_path / "r2c.got"), **task))
    want = open(tmp_path / "r2c.ref").read()
    assert open(tmp_path / "r2c.got").read() == want and len(want) > 5000


def test_tuple_hash_known_answers():
    """CPython 2.7 (64-bit) values computed by hand from Objects/tupleobject.c: hash(()) and hash((1,)) are
    documented constants of that algorithm; the two emulations must agree with each other and with them."""
    from falcon_unzip_b200 import py2compat
    assert py2compat.tuple_hash([]) == 3527539                      # 0x345678 + 97531
    assert py2compat.tuple_hash([1]) == 3430019387558               # ((0x345678 ^ 1) * 1000003) + 97531
    for t in [(1, 2, "a"), (0, 0, ""), (123456, 12345, "m000001/1/0_5001")]:
        hs = [py2compat.str_hash(x) if isinstance(x, str) else py2compat.int_hash(x) for x in t]
        assert py2compat.tuple_hash(hs) == py2emu.py27_tuple_hash(t)
