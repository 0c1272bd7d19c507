"""Per-contig read selection on the device (falcon_unzip_b200/select_reads_from_bam.py: BAM ingest, QNAME rows, record
gather through the C ABI) against oracle/select_oracle.py."""
import os

import numpy as np
import pytest

import select_cases
from oracle import select_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [7, 8])
def test_select_reads_from_bam_matches_oracle(eng, tmp_path, seed, capsys):
    from falcon_unzip_b200 import bam, select_reads_from_bam as srb
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=seed)
    header, want = select_oracle.select(fofn, r2c, ids)
    sam_dir = str(tmp_path / "reads")
    made = srb.select_reads_from_bam(fofn, r2c, ids, sam_dir, level=1)
    assert made == sorted(want) and sorted(os.listdir(sam_dir)) == ["%s.bam" % c for c in sorted(want)]
    for ctg, recs in want.items():
        text, refs, got = bam.read_bam(os.path.join(sam_dir, "%s.bam" % ctg))
        assert bytes(got) == b"".join(recs), ctg
        assert select_oracle.parse_header(text) == header and refs == []
    out = capsys.readouterr().out
    assert "num read_partitions: 5" in out and "ctg, len: 000003F 20" in out


def test_cli_and_empty_selection(eng, tmp_path):
    from falcon_unzip_b200 import select_reads_from_bam as srb
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=5)
    sam_dir = str(tmp_path / "cli")
    srb.main(["fc_select_reads_from_bam.py", "--rawread-to-contigs", r2c, "--rawread-ids", ids, "--sam-dir", sam_dir, fofn])
    assert sorted(os.listdir(sam_dir)) == ["000000F.bam", "000001F.bam", "000004F.bam"]
    # nothing selected: no file at all
    with open(r2c) as f:
        rows = [ln for ln in f if ln.split()[1] == "000002F_001"]
    small = str(tmp_path / "r2c_small")
    with open(small, "w") as f:
        f.write("".join(rows))
    assert srb.select_reads_from_bam(fofn, small, ids, str(tmp_path / "none")) == []
    assert os.listdir(str(tmp_path / "none")) == []


def test_gather_records_and_name_rows(eng, tmp_path):
    """fuz_gather_records / Engine.name_rows on their own: any order, repeats, empty selection, bad indices."""
    from falcon_unzip_b200 import _lib, bam
    rng = np.random.default_rng(2)
    names = ["n%d/%s" % (i, "x" * int(rng.integers(0, 40))) for i in range(300)]
    recs = [bam.encode_record(-1, -1, n, 4, 255, [], "ACGT" * int(rng.integers(0, 600)), aux=bytes(int(rng.integers(0, 9)))) for n in names]
    recs[17] = bam.encode_record(-1, -1, names[17], 4, 255, [], "ACGT" * 40000)          # > 64 KiB
    fn = str(tmp_path / "u.bam")
    bam.write_bam(fn, [], b"".join(recs), header_text="@HD\tVN:1.5\n")
    db = eng.ingest_bam(np.fromfile(fn, dtype=np.uint8))
    assert db.n_rec == 300
    got = eng.name_rows(db)
    assert [b.decode() for b in got.tolist()] == names
    sel = np.concatenate([rng.permutation(300), [17, 17, 0, 299]])
    data, off = eng.gather_records(db, sel)
    assert len(off) == len(sel) + 1 and off[-1] == len(data)
    assert data.tobytes() == b"".join(recs[i] for i in sel.tolist())
    data, off = eng.gather_records(db, np.zeros(0, np.int64))
    assert len(data) == 0 and off.tolist() == [0]
    with pytest.raises(_lib.FuzError):
        eng.gather_records(db, np.asarray([0, 300]))


@pytest.mark.parametrize("window", [70_000, 200_000])
def test_windowed_ingest_equals_oracle(eng, tmp_path, window, capsys):
    """Inputs decoded in windows of BGZF blocks (Engine.ingest_bam_windows + fuz_bam_index_window): windows of one or three
    blocks cut records (one of them 100 kb long) at every window border; the carried bytes must make them whole again."""
    from falcon_unzip_b200 import bam, select_reads_from_bam as srb
    fofn, r2c, ids = select_cases.make_case(str(tmp_path), seed=7)
    header, want = select_oracle.select(fofn, r2c, ids)
    sam_dir = str(tmp_path / "reads_w")
    made = srb.select_reads_from_bam(fofn, r2c, ids, sam_dir, level=1, window_bytes=window)
    assert made == sorted(want)
    for ctg, recs in want.items():
        _text, _refs, got = bam.read_bam(os.path.join(sam_dir, "%s.bam" % ctg))
        assert bytes(got) == b"".join(recs), ctg


def test_window_index_reports_the_cut_record(eng):
    """fuz_bam_index_window on a buffer that ends inside a record: the whole records are indexed, the tail offset is the
    start of the cut one; the plain index rejects the same buffer (broken chain)."""
    import ctypes as C
    import torch
    from conftest import synth_set
    from falcon_unzip_b200 import _lib, engine
    sset = synth_set("tiny")
    off = engine.index_records(sset.records)
    cut = int(off[40]) + 100                                   # 100 bytes into record 40
    d = torch.zeros(cut + 64, dtype=torch.uint8, device=eng.device)
    d[:cut].copy_(torch.from_numpy(sset.records[:cut].copy()))
    rec_off = torch.zeros(200, dtype=torch.int64, device=eng.device)
    cro = torch.zeros(len(sset.refs) + 1, dtype=torch.int32, device=eng.device)
    n_rec, need, tail = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    torch.cuda.synchronize()
    rc = _lib.lib().fuz_bam_index_window(eng.ctx, d.data_ptr(), cut, len(sset.refs), 199, rec_off.data_ptr(), cro.data_ptr(),
                                         C.byref(n_rec), C.byref(need), C.byref(tail))
    assert rc == 0 and n_rec.value == 40 and tail.value == int(off[40])
    assert rec_off[:41].cpu().numpy().tolist() == off[:41].tolist()
    rc = _lib.lib().fuz_bam_index_records(eng.ctx, d.data_ptr(), cut, len(sset.refs), 199, rec_off.data_ptr(), cro.data_ptr(),
                                          C.byref(n_rec), C.byref(need))
    assert rc == _lib.FUZ_E_BADRECORD
