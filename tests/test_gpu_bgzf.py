"""BAM ingest on the device, through the C ABI: k_bgzf_inflate against zlib on every
block type / table shape / stream alignment, CRC32 check, corrupt streams; the parallel record index
against the sequential host walk (records longer than a region, tiny records, header-like bytes inside
aux data at region starts, unmapped tail, broken chains); whole files: inflated stream, index and the
six output files identical to the host-decoded path and to the oracle."""
import ctypes as C
import os
import struct
import zlib

import numpy as np
import pytest

import deflate_cases
from conftest import synth_set

pytestmark = pytest.mark.gpu


def _inflate_blocks(eng, blocks, crcs=None, sizes=None):
    """blocks: list of (payload or None, raw deflate stream).  Returns (status, list of outputs)."""
    import torch
    from falcon_unzip_b200 import _lib
    L = _lib.lib()
    comp, coff = bytearray(), []
    for i, (_p, c) in enumerate(blocks):
        comp += b"\x99" * (i % 4)                       # every stream alignment
        coff.append(len(comp))
        comp += c
    csize = [len(c) for _p, c in blocks]
    sizes = sizes or [len(p) for p, _c in blocks]
    uoff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    dev = eng.device
    d_comp = torch.zeros(len(comp) + 16, dtype=torch.uint8, device=dev)
    d_comp[:len(comp)] = torch.from_numpy(np.frombuffer(bytes(comp) or b"\0", dtype=np.uint8).copy())[:len(comp)].to(dev)
    d_coff = torch.tensor(coff, dtype=torch.int64, device=dev)
    d_cs = torch.tensor(csize, dtype=torch.int32, device=dev)
    d_uoff = torch.from_numpy(uoff).to(dev)
    d_crc = None
    if crcs is not None:
        d_crc = torch.from_numpy(np.asarray(crcs, np.uint32).view(np.int32)).to(dev)
    out = torch.full((int(uoff[-1]) + 64,), 0xEE, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    rc = L.fuz_bgzf_inflate(eng.ctx, d_comp.data_ptr(), len(comp), d_coff.data_ptr(), d_cs.data_ptr(), d_uoff.data_ptr(),
                            d_crc.data_ptr() if d_crc is not None else None, len(blocks), out.data_ptr(), int(uoff[-1]))
    assert rc == 0
    st = eng.status(raise_on_error=False)
    host = out.cpu().numpy()
    assert np.all(host[int(uoff[-1]):] == 0xEE), "wrote past the inflated size"
    return st, [host[uoff[i]:uoff[i + 1]].tobytes() for i in range(len(blocks))]


def test_inflate_matches_zlib(eng):
    streams = deflate_cases.streams()
    blocks = [(d, c) for _n, d, c in streams]
    crcs = [zlib.crc32(d) & 0xFFFFFFFF for d, _c in blocks]
    st, outs = _inflate_blocks(eng, blocks, crcs)
    assert st.error == 0, (st.error, st.error_index, st.reserved[3], streams[st.error_index][0])
    for (name, d, _c), got in zip(streams, outs):
        assert got == d, name
    st, outs = _inflate_blocks(eng, blocks)                 # without the CRC check
    assert st.error == 0 and all(g == d for g, (d, _c) in zip(outs, blocks))


def test_inflate_rejects_corrupt_blocks(eng):
    from falcon_unzip_b200 import _lib
    for name, comp, size in deflate_cases.corrupt_streams():
        try:
            want = zlib.decompress(comp, -15)
        except zlib.error:
            want = None
        st, outs = _inflate_blocks(eng, [(None, comp)], sizes=[size])
        if want is not None and len(want) == size:
            assert st.error == 0 and outs[0] == want, name
        else:
            assert st.error == _lib.FUZ_E_FORMAT and st.error_index == 0, name
    data = b"payload whose CRC does not match" * 100
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = c.compress(data) + c.flush()
    st, _ = _inflate_blocks(eng, [(data, comp), (data, comp)], [zlib.crc32(data), zlib.crc32(data) ^ 1])
    assert st.error == _lib.FUZ_E_FORMAT and st.error_index == 1 and st.reserved[3] == 8
    st, _ = _inflate_blocks(eng, [(data, comp)] * 3, [zlib.crc32(data)] * 3)
    assert st.error == 0


def _index_device(eng, records: bytes, n_ref: int, cap=None):
    import torch
    from falcon_unzip_b200 import _lib
    L = _lib.lib()
    dev = eng.device
    n = len(records)
    d = torch.zeros(n + 64, dtype=torch.uint8, device=dev)
    if n:
        d[:n] = torch.from_numpy(np.frombuffer(records, dtype=np.uint8).copy()).to(dev)
    cap = n // 36 + 2 if cap is None else cap
    off = torch.full((cap + 1,), -7, dtype=torch.int64, device=dev)
    cro = torch.full((n_ref + 1,), -7, dtype=torch.int32, device=dev)
    n_rec, need = C.c_int64(-1), C.c_int64(-1)
    torch.cuda.synchronize(dev)
    rc = L.fuz_bam_index_records(eng.ctx, d.data_ptr(), n, n_ref, cap, off.data_ptr(), cro.data_ptr(), C.byref(n_rec), C.byref(need))
    return rc, n_rec.value, need.value, off.cpu().numpy(), cro.cpu().numpy()


def _host_index(records: bytes, n_ref: int):
    from falcon_unzip_b200 import bam, engine
    off = bam.index_records(records)
    a = np.frombuffer(records, np.uint8)
    ref = engine.record_refids(a, off) if len(off) > 1 else np.zeros(0, np.int32)
    key = np.where(ref < 0, n_ref, ref)
    return off, np.searchsorted(key, np.arange(n_ref + 1), side="left").astype(np.int32)


def _rec(refid, pos, name, l_seq, aux=b"", rng=None):
    from falcon_unzip_b200 import bam
    seq = "".join("ACGT"[i] for i in (rng.integers(0, 4, l_seq) if rng is not None else np.zeros(l_seq, int)))
    return bam.encode_record(refid, pos, name, 0, 254, [(l_seq, "=")] if l_seq else [], seq or "*", aux=aux)


def _fake_header(block_size: int) -> bytes:
    body = struct.pack("<iiBBHHHiiii", 0, 5, 2, 0, 0, 0, 0, 0, -1, -1, 0) + b"x\0"
    return struct.pack("<i", block_size) + body


def test_record_index_matches_host_walk(eng):
    rng = np.random.default_rng(5)
    cases = {}
    cases["empty"] = (b"", 3)
    cases["one"] = (_rec(0, 0, "r0", 10, rng=rng), 1)
    cases["tiny_records"] = (b"".join(_rec(i // 2000, i, "t%d" % i, int(rng.integers(0, 3)), rng=rng) for i in range(6000)), 3)
    cases["long_records"] = (b"".join(_rec(0, 10 * i, "m/%d/0_1" % i, int(rng.integers(90000, 140000)), rng=rng) for i in range(12)), 1)
    mixed = [_rec(int(i >= 150), i * 7 % 1000 + (i >= 150) * 0, "m/%d/0_%d" % (i, i), int(rng.integers(1, 30000)), rng=rng) for i in range(300)]
    mixed += [_rec(-1, -1, "unmapped%d" % i, 100, rng=rng) for i in range(5)]
    cases["mixed_unmapped_tail"] = (b"".join(mixed), 4)
    # header-like bytes inside aux data, exactly at the start of regions 1 and 2, whose own chain
    # (two fake records) lands on a true record start: the guess of those regions must be discarded
    head = b"".join(_rec(0, i, "h%d" % i, 5000, rng=rng) for i in range(5))
    for target in (65536, 131072):
        aux_start = len(head) + 36 + 4          # core + name "big\0" (no CIGAR, no SEQ)
        aux_len = 200000
        aux = bytearray(rng.integers(0, 256, aux_len, dtype=np.uint8).tobytes())
        p1 = target - aux_start
        bs1 = 1000
        p2 = p1 + 4 + bs1
        bs2 = aux_len - p2 - 4
        aux[p1:p1 + len(_fake_header(bs1))] = _fake_header(bs1)
        aux[p2:p2 + len(_fake_header(bs2))] = _fake_header(bs2)
        big = _rec(0, 9, "big", 0, aux=b"XYB" + bytes([255]) + bytes(aux[4:]))   # keep offsets: overwrite first 4 aux bytes
        tail = b"".join(_rec(0, 10 + i, "t%d" % i, 3000, rng=rng) for i in range(40))
        cases["fake_headers_%d" % target] = (head + big + tail, 2)
    for name, (records, n_ref) in cases.items():
        want_off, want_cro = _host_index(records, n_ref)
        rc, n_rec, need, off, cro = _index_device(eng, records, n_ref)
        assert rc == 0, (name, rc)
        assert n_rec == len(want_off) - 1 == need, name
        assert np.array_equal(off[:n_rec + 1], want_off), name
        assert np.array_equal(cro, want_cro), name
    # capacity protocol
    from falcon_unzip_b200 import _lib
    records, n_ref = cases["tiny_records"]
    rc, n_rec, need, _off, _cro = _index_device(eng, records, n_ref, cap=100)
    assert rc == _lib.FUZ_E_CAPACITY and need == 6000
    # broken chain / unsorted reference ids
    records = cases["mixed_unmapped_tail"][0]
    rc, *_ = _index_device(eng, records[:-3], 4)
    assert rc == _lib.FUZ_E_BADRECORD
    bad = bytearray(records); bad[0:4] = struct.pack("<i", 20)
    rc, *_ = _index_device(eng, bytes(bad), 4)
    assert rc == _lib.FUZ_E_BADRECORD
    swapped = _rec(1, 0, "a", 50, rng=rng) + _rec(0, 0, "b", 50, rng=rng)
    rc, *_ = _index_device(eng, swapped, 2)
    assert rc == _lib.FUZ_E_UNSORTED


@pytest.mark.parametrize("cfg,level", [("quirks", 1), ("quirks", 6), ("long", 1)])
def test_ingest_bam_file(eng, tmp_path, cfg, level):
    from falcon_unzip_b200 import bam
    sset = synth_set(cfg)
    fn = str(tmp_path / "in.bam")
    bam.write_bam(fn, sset.refs, sset.records.tobytes(), level=level)
    db = eng.ingest_bam(np.fromfile(fn, dtype=np.uint8))
    assert [tuple(r) for r in db.refs] == [tuple(r) for r in sset.refs]
    assert np.array_equal(db.records(), np.asarray(sset.records))
    want_off, want_cro = _host_index(sset.records.tobytes(), len(sset.refs))
    assert db.n_rec == len(want_off) - 1 == db.n_mapped
    assert np.array_equal(db.rec_off.cpu().numpy()[:db.n_rec + 1], want_off)
    assert np.array_equal(db.ctg_rec_off.cpu().numpy(), want_cro)
    # a flipped payload bit is caught by the CRC check
    from falcon_unzip_b200 import _lib
    img = np.fromfile(fn, dtype=np.uint8)
    n_blk = _lib.lib().fuz_host_bgzf_index(img.ctypes.data, len(img), 0, None, None, None, None)
    coff, csize, uoff = np.empty(n_blk, np.int64), np.empty(n_blk, np.int32), np.empty(n_blk + 1, np.int64)
    _lib.lib().fuz_host_bgzf_index(img.ctypes.data, len(img), n_blk, coff.ctypes.data, csize.ctypes.data, uoff.ctypes.data, None)
    b = n_blk // 2
    img[coff[b] + csize[b] // 2] ^= 0x10                 # inside the deflate stream of a middle block
    with pytest.raises(_lib.FuzError):
        eng.ingest_bam(img)


@pytest.mark.parametrize("cfg", ["quirks", "tiny"])
def test_phase_bam_equals_host_decoded_path_and_oracle(eng, tmp_path, cfg):
    from falcon_unzip_b200 import bam, phasing, synth
    from oracle import c_oracle
    sset = synth_set(cfg)
    fn, fa = str(tmp_path / "in.bam"), str(tmp_path / "ref.fa")
    bam.write_bam(fn, sset.refs, sset.records.tobytes())
    synth.write_fasta(fa, sset)
    res_b, files_b = phasing.phase_bam(fn, fa, str(tmp_path / "dev"))
    names = [r[0] for r in sset.refs]
    res_h, files_h = phasing.phase_contigs(sset.records, names, sset.ref_seqs, str(tmp_path / "host"))
    assert (res_b.n_sites, res_b.n_vmap, res_b.n_atable, res_b.n_reads) == (res_h.n_sites, res_h.n_vmap, res_h.n_atable, res_h.n_reads)
    assert res_b.aligned_bases == res_h.aligned_bases and res_b.n_sites > 0
    for c, name in enumerate(names):
        want = c_oracle.run_phasing_stages(sset.contig_records(c), name, sset.ref_seqs[c], str(tmp_path / "oracle"))
        for k in want:
            a = open(want[k]).read()
            assert open(files_b[name][k]).read() == a, (name, k)
            assert open(files_h[name][k]).read() == a, (name, k)


def test_phase_bam_with_empty_contig_and_unmapped_tail(eng, tmp_path):
    """A reference without reads in the middle, unmapped records (refID -1) behind the last contig, and a BAM
    that holds no record at all: the device-decoded path must agree with the host-decoded one."""
    import struct
    from falcon_unzip_b200 import bam, engine, phasing, synth
    sset = synth_set("tiny")
    names = [sset.refs[0][0], "empty_ctg", sset.refs[1][0]]
    refs = [sset.refs[0], ("empty_ctg", 5000), sset.refs[1]]
    recs = np.asarray(sset.records).copy()
    off = np.asarray(sset.rec_off)
    for r in np.flatnonzero(np.asarray(sset.rec_ctg) == 1):          # second contig becomes refID 2
        recs[off[r] + 4:off[r] + 8] = np.frombuffer(struct.pack("<i", 2), np.uint8)
    tail = b"".join(bam.encode_record(-1, -1, "unmapped/%d" % i, 4, 0, [], "ACGT" * 50) for i in range(7))
    fn, fa = str(tmp_path / "in.bam"), str(tmp_path / "ref.fa")
    bam.write_bam(fn, refs, recs.tobytes() + tail)
    with open(fa, "w") as f:
        for (n, _l), seq in zip(refs, [sset.ref_seqs[0], "A" * 5000, sset.ref_seqs[1]]):
            f.write(">%s\n%s\n" % (n, seq))
    db = eng.ingest_bam(np.fromfile(fn, dtype=np.uint8))
    assert db.n_rec == len(off) - 1 + 7 and db.n_mapped == len(off) - 1
    cro = db.ctg_rec_off.cpu().numpy()
    assert cro[1] == cro[2] and cro[3] == db.n_mapped
    res_b, files_b = phasing.phase_bam(fn, fa, str(tmp_path / "dev"))
    res_h, files_h = phasing.phase_contigs(recs, names, [sset.ref_seqs[0], "A" * 5000, sset.ref_seqs[1]], str(tmp_path / "host"))
    assert res_b.n_sites == res_h.n_sites > 0 and res_b.n_reads == res_h.n_reads
    for n in names:
        for k in files_h[n]:
            assert open(files_b[n][k]).read() == open(files_h[n][k]).read(), (n, k)
    # no records at all
    bam.write_bam(fn, refs, b"")
    res_e, _ = phasing.phase_bam(fn, fa, str(tmp_path / "none"))
    assert (res_e.n_sites, res_e.n_vmap, res_e.n_reads) == (0, 0, 0)


def test_inflate_fuzz(eng):
    """Random payload mixtures (runs, repeats at random distances, random bytes, small alphabets) x random zlib
    parameters, 300 blocks in one launch, against zlib; sizes from 0 to 65536."""
    rng = np.random.default_rng(2024)
    blocks = []
    for k in range(300):
        n = int(rng.choice([0, 1, 2, 31, 32, 33, 255, 4096, 65535, 65536, int(rng.integers(0, 65537))]))
        parts, size = [], 0
        while size < n:
            kind = int(rng.integers(0, 5))
            m = int(min(n - size, rng.integers(1, 3000)))
            if kind == 0:
                piece = bytes([int(rng.integers(0, 256))]) * m
            elif kind == 1:
                piece = bytes(rng.integers(0, 256, m, dtype=np.uint8))
            elif kind == 2:
                piece = bytes(rng.integers(0, int(rng.integers(2, 20)), m, dtype=np.uint8))
            elif kind == 3 and size > 0:
                joined = b"".join(parts)
                d = int(rng.integers(1, min(size, 32768) + 1))
                src = joined[size - d:size - d + m]
                piece = (src * (m // max(len(src), 1) + 1))[:m]
            else:
                piece = (b"ACGT" * (m // 4 + 1))[:m]
            parts.append(piece)
            size += len(piece)
        data = b"".join(parts)[:n]
        level = int(rng.choice([0, 1, 1, 6, 9]))
        strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        c = zlib.compressobj(level, zlib.DEFLATED, -15, int(rng.integers(1, 10)), strat)
        comp = c.compress(data)
        if rng.random() < 0.3 and len(data) > 10:          # several deflate blocks inside one stream
            comp += c.flush(zlib.Z_FULL_FLUSH)
            extra = bytes(rng.integers(0, 256, int(rng.integers(0, 50)), dtype=np.uint8))
            if len(data) + len(extra) <= 65536:
                comp += c.compress(extra)
                data += extra
        comp += c.flush()
        assert zlib.decompress(comp, -15) == data
        blocks.append((data, comp))
    st, outs = _inflate_blocks(eng, blocks, [zlib.crc32(d) & 0xFFFFFFFF for d, _c in blocks])
    assert st.error == 0, (st.error, st.error_index, st.reserved[3])
    for k, ((d, _c), got) in enumerate(zip(blocks, outs)):
        assert got == d, k


def test_phase_bam_from_one_file_per_contig(eng, tmp_path):
    """The reference leaves one sorted BAM per contig (unzip.py:90), each with its own one-entry header and refID 0:
    a list of such files goes through one device batch and gives the files of the single-BAM run."""
    import struct
    from falcon_unzip_b200 import bam, phasing, synth
    sset = synth_set("quirks")
    fa = str(tmp_path / "ref.fa")
    synth.write_fasta(fa, sset)
    fns = []
    for c, (name, ln) in enumerate(sset.refs):
        rec = np.frombuffer(sset.contig_records(c), np.uint8).copy()
        off = bam.index_records(rec.tobytes())
        for o in off[:-1].tolist():
            rec[o + 4:o + 8] = np.frombuffer(struct.pack("<i", 0), np.uint8)          # refID 0 inside its own file
        tail = bam.encode_record(-1, -1, "unmapped/%d" % c, 4, 0, [], "ACGT" * 10) if c % 2 == 0 else b""
        fn = str(tmp_path / ("%s_sorted.bam" % name))
        bam.write_bam(fn, [(name, ln)], rec.tobytes() + tail)
        fns.append(fn)
    one = str(tmp_path / "all.bam")
    bam.write_bam(one, sset.refs, sset.records.tobytes())
    res_m, files_m = phasing.phase_bam(fns, fa, str(tmp_path / "many"))
    res_1, files_1 = phasing.phase_bam(one, fa, str(tmp_path / "one"))
    assert (res_m.n_sites, res_m.n_vmap, res_m.n_atable, res_m.n_reads) == (res_1.n_sites, res_1.n_vmap, res_1.n_atable, res_1.n_reads)
    assert res_m.n_sites > 0
    for n in files_1:
        for k in files_1[n]:
            assert open(files_m[n][k]).read() == open(files_1[n][k]).read(), (n, k)


def _bam_image(tmp_path, tag, refs, records: bytes, level=1):
    from falcon_unzip_b200 import bam
    fn = str(tmp_path / ("%s.bam" % tag))
    bam.write_bam(fn, refs, records, level=level)
    return np.fromfile(fn, dtype=np.uint8)


def test_ingest_bams_many_files_one_batch(eng, tmp_path):
    """fuz_bam_index_files: a list of BAM files (one inflate launch, one record index over all of them) gives the mapped
    records of the files back to back, with the offsets and per-reference ranges of the sequential host walk."""
    from falcon_unzip_b200 import _lib
    rng = np.random.default_rng(11)
    files = []                                           # (refs, mapped records, unmapped tail)
    # tiny records: hundreds per region, several regions
    files.append(([("a0", 9000), ("a1", 9000), ("a2", 9000)],
                  b"".join(_rec(i // 2000, i % 2000, "t%d" % i, int(rng.integers(0, 3)), rng=rng) for i in range(6000)), b""))
    files.append(([("none", 500)], b"", b""))                                            # header only
    files.append(([("only_unmapped", 500)], b"", b"".join(_rec(-1, -1, "u%d" % i, 40, rng=rng) for i in range(7))))
    # records longer than a region, and an unmapped tail
    files.append(([("long", 200000)], b"".join(_rec(0, 10 * i, "m/%d/0_1" % i, int(rng.integers(90000, 140000)), rng=rng) for i in range(9)),
                  _rec(-1, -1, "ux", 100, rng=rng)))
    # two references, the first one empty; PacBio-like sizes
    files.append(([("e0", 1000), ("e1", 50000)], b"".join(_rec(1, 3 * i, "p/%d/0_%d" % (i, i), int(rng.integers(1, 30000)), rng=rng) for i in range(150)),
                  b"".join(_rec(-1, -1, "uy%d" % i, 3000, rng=rng) for i in range(3))))
    files.append(([], b"", b""))                                                          # no reference at all
    # header-like bytes inside aux data at the start of a region of THIS file's record stream
    head = b"".join(_rec(0, i, "h%d" % i, 5000, rng=rng) for i in range(5))
    aux_start, aux_len = len(head) + 36 + 4, 200000
    aux = bytearray(rng.integers(0, 256, aux_len, dtype=np.uint8).tobytes())
    p1 = 65536 - aux_start
    p2 = p1 + 4 + 1000
    aux[p1:p1 + len(_fake_header(1000))] = _fake_header(1000)
    aux[p2:p2 + len(_fake_header(aux_len - p2 - 4))] = _fake_header(aux_len - p2 - 4)
    big = _rec(0, 9, "big", 0, aux=b"XYB" + bytes([255]) + bytes(aux[4:]))
    files.append(([("fake", 9999), ("fake2", 10)], head + big + b"".join(_rec(0, 10 + i, "t%d" % i, 3000, rng=rng) for i in range(40)), b""))
    images = [_bam_image(tmp_path, "f%d" % k, refs, rec + tail, level=(1 if k % 2 else 6)) for k, (refs, rec, tail) in enumerate(files)]
    db = eng.ingest_bams(images)
    want = b"".join(rec for _r, rec, _t in files)
    assert [tuple(r) for r in db.refs] == [tuple(r) for refs, _r, _t in files for r in refs]
    assert db.rec_bytes == len(want)
    assert db.records().tobytes() == want
    want_off, want_cro, base = [], [], 0
    at = 0
    for refs, rec, _t in files:
        off, cro = _host_index(rec, len(refs))
        want_off.append(off[:-1] + at)
        want_cro.append(cro[:-1] + base)
        at += len(rec)
        base += len(off) - 1
    want_off = np.concatenate(want_off + [np.asarray([at])])
    want_cro = np.concatenate(want_cro + [np.asarray([base])])
    assert db.n_rec == db.n_mapped == base
    assert np.array_equal(db.rec_off.cpu().numpy()[:base + 1], want_off)
    assert np.array_equal(db.ctg_rec_off.cpu().numpy(), want_cro)
    # every file alone through the single-file path gives the same pieces
    for k in (0, 3, 4):
        one = eng.ingest_bam(images[k])
        rec = files[k][1]
        assert one.records()[:len(rec)].tobytes() == rec and one.n_mapped == len(_host_index(rec, len(files[k][0]))[0]) - 1
    # errors: unsorted reference ids inside one file, a broken chain in one file, a flipped bit, a repeated reference name
    swapped = _rec(1, 0, "a", 50, rng=rng) + _rec(0, 0, "b", 50, rng=rng)
    with pytest.raises(_lib.FuzError) as ei:
        eng.ingest_bams([images[4], _bam_image(tmp_path, "sw", [("s0", 100), ("s1", 100)], swapped)])
    assert ei.value.code == _lib.FUZ_E_UNSORTED
    good = _rec(0, 0, "a", 50, rng=rng) * 3
    with pytest.raises(_lib.FuzError) as ei:
        eng.ingest_bams([_bam_image(tmp_path, "br", [("b0", 100)], good[:-3]), images[4]])
    assert ei.value.code == _lib.FUZ_E_BADRECORD
    with pytest.raises(_lib.FuzError) as ei:                      # refID 1 in a file that lists one reference
        eng.ingest_bams([images[3], _bam_image(tmp_path, "rf", [("r0", 100)], _rec(1, 0, "a", 50, rng=rng))])
    assert ei.value.code == _lib.FUZ_E_BADRECORD
    img = images[4].copy()
    n_blk = _lib.lib().fuz_host_bgzf_index(img.ctypes.data, len(img), 0, None, None, None, None)
    coff, csize, uoff = np.empty(n_blk, np.int64), np.empty(n_blk, np.int32), np.empty(n_blk + 1, np.int64)
    _lib.lib().fuz_host_bgzf_index(img.ctypes.data, len(img), n_blk, coff.ctypes.data, csize.ctypes.data, uoff.ctypes.data, None)
    img[coff[n_blk // 2] + csize[n_blk // 2] // 2] ^= 0x10           # inside the deflate stream of a middle block
    with pytest.raises(_lib.FuzError):
        eng.ingest_bams([images[3], img])
    with pytest.raises(_lib.FuzError) as ei:
        eng.ingest_bams([images[3], images[3]])
    assert ei.value.code == _lib.FUZ_E_ARG
    # the engine is usable afterwards
    db2 = eng.ingest_bams(images[3:5])
    assert db2.records().tobytes() == files[3][1] + files[4][1]
