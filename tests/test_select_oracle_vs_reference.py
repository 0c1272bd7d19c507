"""oracle/select_oracle.py against the reference's own select_reads_from_bam.py (executed through oracle/ref_exec.py
with a pysam stand-in; build container only), and the host-side tables of the product module against both."""
import os

import pytest

import select_cases
from oracle import ref_exec, select_oracle

needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(ref_exec.REF_ROOT, "falcon_unzip", "select_reads_from_bam.py")),
                               reason="reference tree not present")


@needs_ref
@pytest.mark.parametrize("seed", [7, 8])
def test_oracle_matches_reference(tmp_path, seed):
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=seed)
    sam_dir = str(tmp_path / "ref_out")
    ref = ref_exec.run_select_reads(fofn, r2c, ids, sam_dir)
    header, out = select_oracle.select(fofn, r2c, ids)
    assert sorted(ref) == sorted(os.path.join(sam_dir, "%s.bam" % c) for c in out)
    assert set(out) == {"000000F", "000001F", "000004F"}
    for ctg, recs in out.items():
        ref_header, ref_recs = ref[os.path.join(sam_dir, "%s.bam" % ctg)]
        assert ref_recs == recs, ctg
        assert ref_header == header, ctg
    assert "PG" not in header and [d["ID"] for d in header["RG"]] == ["rg0", "rg1", "rg1b", "rg2"]


def test_product_tables_match_oracle(tmp_path):
    """read -> contig table and merged header text of the product module (host code, no device) against the oracle."""
    from falcon_unzip_b200 import bam, select_reads_from_bam as srb
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=9)
    header, out = select_oracle.select(fofn, r2c, ids)
    part, r2ctgs = srb.read_tables(r2c, ids)
    target = srb.read_to_selected_ctg(part, r2ctgs)
    assert set(target.values()) == set(out)
    names_written = {rec[36:36 + rec[12] - 1].decode() for recs in out.values() for rec in recs}
    assert names_written <= set(target) and all(target[n] in out for n in names_written)
    base = os.path.dirname(fofn)
    texts = [bam.read_bam(os.path.join(base, "movie%d.subreads.bam" % k))[0] for k in range(3)]
    assert select_oracle.parse_header(srb.merged_header_text(texts)) == header
    with pytest.raises(KeyError):
        srb.merged_header_text([texts[0], "@HD\tVN:1.5\n"])
    with pytest.raises(KeyError):
        srb.merged_header_text(["@HD\tVN:1.5\n", texts[1]])
    assert srb.merged_header_text(["@HD\tVN:1.5\n@PG\tID:x\n"]) == "@HD\tVN:1.5\n"


def test_bam_writer_round_trip(tmp_path):
    from falcon_unzip_b200 import bam
    import numpy as np
    rng = np.random.default_rng(3)
    recs = [bam.encode_record(-1, -1, "r%d" % i, 4, 255, [], "ACGT" * int(rng.integers(1, 9000))) for i in range(40)]
    fn = str(tmp_path / "w.bam")
    w = bam.BamWriter(fn, "@HD\tVN:1.5\n", [], level=1)
    for i in range(0, 40, 7):
        w.write(b"".join(recs[i:i + 7]))
    w.close()
    w.close()
    text, refs, got = bam.read_bam(fn)
    assert text == "@HD\tVN:1.5\n" and refs == [] and bytes(got) == b"".join(recs)
    # one large write (blocks compressed on the thread pool) gives the bytes of the serial writer
    big = b"".join(recs) * 3
    fn2, fn3 = str(tmp_path / "big.bam"), str(tmp_path / "serial.bam")
    w = bam.BamWriter(fn2, "@HD\tVN:1.5\n", [("c", 9)], level=6)
    w.write(big)
    w.close()
    bam.write_bam(fn3, [("c", 9)], big, header_text="@HD\tVN:1.5\n", level=6)
    assert open(fn2, "rb").read() == open(fn3, "rb").read()


@needs_ref
def test_oracle_matches_reference_on_random_tables(tmp_path):
    """Random rawread_to_contigs tables over fixed BAMs: contig sizes around the > 20 threshold, score ties between contigs,
    several rank-0 rows per read, NA rows."""
    import numpy as np
    from falcon_unzip_b200 import bam
    rng = np.random.default_rng(123)
    root = str(tmp_path)
    names = ["mv/%d/0_%d" % (i, 100 + i) for i in range(90)]
    with open(os.path.join(root, "ids"), "w") as f:
        f.write("\n".join(names))                               # no trailing newline here
    fns = []
    for k in range(2):
        recs = [bam.encode_record(-1, -1, names[int(j)], 4, 255, [], "ACGT" * int(rng.integers(1, 50)))
                for j in rng.integers(0, 90, 120)]
        fn = os.path.join(root, "f%d.bam" % k)
        bam.write_bam(fn, [], b"".join(recs), header_text="@HD\tVN:1.5\n@RG\tID:r%d\n@PG\tID:p\n" % k)
        fns.append(fn)
    fofn = os.path.join(root, "fofn")
    with open(fofn, "w") as f:
        f.write("\n".join(os.path.basename(x) for x in fns) + "\n")
    ctgs = ["000000F", "000000F_001", "000001F", "NA"]
    n_nonempty, sizes = 0, set()
    for trial in range(25):
        rows = []
        for i in range(90):
            for _ in range(int(rng.integers(0, 4))):
                rows.append("%09d %s 3 %d %d 1" % (i, ctgs[int(rng.choice(4, p=[0.35, 0.3, 0.25, 0.1]))],
                                                 int(rng.choice(3, p=[0.6, 0.2, 0.2])), -100 * int(rng.integers(1, 4))))
        r2c = os.path.join(root, "r2c_%d" % trial)
        with open(r2c, "w") as f:
            f.write("".join(r + "\n" for r in rows))
        sam_dir = os.path.join(root, "out_%d" % trial)
        ref = ref_exec.run_select_reads(fofn, r2c, os.path.join(root, "ids"), sam_dir)
        header, out = select_oracle.select(fofn, r2c, os.path.join(root, "ids"))
        assert sorted(ref) == sorted(os.path.join(sam_dir, "%s.bam" % c) for c in out), trial
        for ctg, recs in out.items():
            assert ref[os.path.join(sam_dir, "%s.bam" % ctg)] == (header, recs), (trial, ctg)
        n_nonempty += bool(out)
        sizes.add(len(out))
    n_sizes = len(sizes)
    assert n_nonempty >= 10 and n_sizes > 1                 # the > 20 rule cuts both ways in the sample
