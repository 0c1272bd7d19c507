"""BAM / BGZF / SAM-text plumbing for the phasing hot path (host side).

The reference never parses BAM: it pipes ``samtools view <bam> <ctg>`` and splits SAM
text (reference falcon_unzip/phasing.py:27,42-59).  Here the BAM container is decoded
natively (SAM spec section 4: BGZF blocks -> header -> alignment records) so that the
*uncompressed alignment records* can be handed to the CUDA kernels verbatim.  The SAM
text view of the same records (what ``samtools view`` would print) is produced by
:func:`sam_lines_from_records`; it is the input of the CPU oracle, so that both sides
of every parity test see the same records.

Record layout (all little endian), SAM spec 4.2::

    block_size:i32  refID:i32  pos:i32  l_read_name:u8  mapq:u8  bin:u16
    n_cigar_op:u16  flag:u16  l_seq:i32  next_refID:i32  next_pos:i32  tlen:i32
    read_name[l_read_name]  cigar:u32[n_cigar_op]  seq:u8[(l_seq+1)/2]  qual[l_seq]  aux

``cigar = len << 4 | op`` with op an index into ``MIDNSHP=X``; ``seq`` holds 4-bit
codes into ``=ACMGRSVTWYHKDBN`` (high nibble first).
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Iterable, Iterator, List, Sequence, Tuple

import numpy as np

CIGAR_OPS = "MIDNSHP=X"
SEQ_CODES = "=ACMGRSVTWYHKDBN"
_SEQ_ENC = {c: i for i, c in enumerate(SEQ_CODES)}
_OP_ENC = {c: i for i, c in enumerate(CIGAR_OPS)}
CORE_BYTES = 36  # block_size + 32-byte fixed core

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
_BGZF_MAX_PAYLOAD = 0xFF00


# --------------------------------------------------------------------------- records
def reg2bin(beg: int, end: int) -> int:
    """UCSC binning scheme (SAM spec 5.3)."""
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def pack_seq(seq: str) -> bytes:
    codes = np.frombuffer(seq.encode("ascii"), dtype=np.uint8)
    lut = np.full(256, 15, dtype=np.uint8)
    for c, i in _SEQ_ENC.items():
        lut[ord(c)] = i
        lut[ord(c.lower())] = i
    nib = lut[codes]
    if len(nib) & 1:
        nib = np.concatenate([nib, np.zeros(1, np.uint8)])
    return ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8).tobytes()


def unpack_seq(packed: bytes, l_seq: int) -> str:
    b = np.frombuffer(packed, dtype=np.uint8)
    nib = np.empty(len(b) * 2, dtype=np.uint8)
    nib[0::2] = b >> 4
    nib[1::2] = b & 15
    lut = np.frombuffer(SEQ_CODES.encode("ascii"), dtype=np.uint8)
    return lut[nib[:l_seq]].tobytes().decode("ascii")


def parse_cigar_string(cigar: str) -> List[Tuple[int, str]]:
    out, num = [], 0
    if cigar == "*":
        return out
    for ch in cigar:
        if ch.isdigit():
            num = num * 10 + ord(ch) - 48
        else:
            out.append((num, ch))
            num = 0
    return out


def encode_record(refid: int, pos: int, name: str, flag: int, mapq: int,
                  cigar: Sequence[Tuple[int, str]], seq: str, qual: bytes | None = None,
                  aux: bytes = b"") -> bytes:
    """One BAM alignment record including its leading block_size."""
    name_b = name.encode("ascii") + b"\0"
    if len(name_b) > 255:
        raise ValueError("read name too long for BAM")
    ref_len = sum(n for n, op in cigar if op in "MDN=X")
    l_seq = 0 if seq == "*" else len(seq)
    cig = b"".join(struct.pack("<I", (n << 4) | _OP_ENC[op]) for n, op in cigar)
    seq_b = b"" if l_seq == 0 else pack_seq(seq)
    if qual is None:
        qual = b"\xff" * l_seq
    body = struct.pack("<iiBBHHHiiii", refid, pos, len(name_b), mapq,
                       reg2bin(pos, pos + max(ref_len, 1)), len(cigar), flag, l_seq,
                       -1, -1, 0) + name_b + cig + seq_b + qual + aux
    return struct.pack("<i", len(body)) + body


def index_records(buf) -> np.ndarray:
    """Offsets (int64, n+1 entries) of the records in a concatenated record buffer."""
    mv = memoryview(buf)
    n = len(mv)
    offs = [0]
    o = 0
    while o < n:
        if o + 4 > n:
            raise ValueError("truncated BAM record stream")
        (bs,) = struct.unpack_from("<i", mv, o)
        if bs < 32 or o + 4 + bs > n:
            raise ValueError("corrupt BAM record at byte %d" % o)
        o += 4 + bs
        offs.append(o)
    return np.asarray(offs, dtype=np.int64)


def iter_records(buf, offs: np.ndarray | None = None) -> Iterator[dict]:
    """Decode records to dicts (slow path: tests, SAM rendering, small inputs)."""
    mv = memoryview(buf)
    if offs is None:
        offs = index_records(buf)
    for i in range(len(offs) - 1):
        o = int(offs[i])
        (bs, refid, pos, l_name, mapq, _bin, n_cig, flag, l_seq, nref, npos,
         tlen) = struct.unpack_from("<iiiBBHHHiiii", mv, o)
        p = o + CORE_BYTES
        name = bytes(mv[p:p + l_name - 1]).decode("ascii")
        p += l_name
        cig = np.frombuffer(mv[p:p + 4 * n_cig], dtype="<u4")
        p += 4 * n_cig
        seq = unpack_seq(bytes(mv[p:p + (l_seq + 1) // 2]), l_seq)
        yield dict(refid=refid, pos=pos, name=name, mapq=mapq, flag=flag, l_seq=l_seq,
                   cigar=[(int(c >> 4), CIGAR_OPS[int(c & 15)]) for c in cig], seq=seq,
                   next_refid=nref, next_pos=npos, tlen=tlen)


def sam_lines_from_records(buf, refs: Sequence[Tuple[str, int]],
                           refid: int | None = None) -> List[str]:
    """The text ``samtools view`` prints for these records (11 mandatory fields)."""
    out = []
    for r in iter_records(buf):
        if refid is not None and r["refid"] != refid:
            continue
        rname = refs[r["refid"]][0] if r["refid"] >= 0 else "*"
        cigar = "".join("%d%s" % c for c in r["cigar"]) or "*"
        seq = r["seq"] or "*"
        out.append("\t".join([r["name"], str(r["flag"]), rname, str(r["pos"] + 1),
                              str(r["mapq"]), cigar, "*", "0", "0", seq, "*"]))
    return out


def records_from_sam_lines(lines: Iterable[str], refs: Sequence[Tuple[str, int]]) -> bytes:
    """SAM text -> concatenated BAM records (header lines skipped like phasing.py:44-45)."""
    ref_index = {name: i for i, (name, _len) in enumerate(refs)}
    chunks = []
    for line in lines:
        f = line.strip().split()
        if not f or f[0][0] == "@":
            continue
        chunks.append(encode_record(ref_index.get(f[2], -1), int(f[3]) - 1, f[0], int(f[1]),
                                    int(f[4]) if f[4].isdigit() else 255,
                                    parse_cigar_string(f[5]), f[9]))
    return b"".join(chunks)


# --------------------------------------------------------------------------- BGZF
def _bgzf_block(payload: bytes, level: int = 1) -> bytes:
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    data = comp.compress(payload) + comp.flush()
    bsize = len(data) + 25
    return (b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00"
            + struct.pack("<H", bsize) + data
            + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload)))


def write_bam(path: str, refs: Sequence[Tuple[str, int]], records: bytes,
              header_text: str | None = None, level: int = 1) -> None:
    """Write a coordinate-sorted BAM (BGZF container + header + ``records``)."""
    if header_text is None:
        header_text = "@HD\tVN:1.5\tSO:coordinate\n" + "".join(
            "@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    ht = header_text.encode("ascii")
    hdr = b"BAM\1" + struct.pack("<i", len(ht)) + ht + struct.pack("<i", len(refs))
    for name, ln in refs:
        nb = name.encode("ascii") + b"\0"
        hdr += struct.pack("<i", len(nb)) + nb + struct.pack("<i", ln)
    with open(path, "wb") as f:
        f.write(_bgzf_block(hdr, level))
        mv = memoryview(records)
        for o in range(0, len(mv), _BGZF_MAX_PAYLOAD):
            f.write(_bgzf_block(bytes(mv[o:o + _BGZF_MAX_PAYLOAD]), level))
        f.write(_BGZF_EOF)


def bam_header_bytes(header_text: str, refs: Sequence[Tuple[str, int]]) -> bytes:
    ht = header_text.encode("latin-1")
    hdr = b"BAM\1" + struct.pack("<i", len(ht)) + ht + struct.pack("<i", len(refs))
    for name, ln in refs:
        nb = name.encode("ascii") + b"\0"
        hdr += struct.pack("<i", len(nb)) + nb + struct.pack("<i", ln)
    return hdr


_compress_pool = None


def _pool():
    """Threads for BGZF block compression (zlib releases the GIL; the blocks of a file are independent)."""
    global _compress_pool
    if _compress_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _compress_pool = ThreadPoolExecutor(max(1, min(16, len(os.sched_getaffinity(0)))))
    return _compress_pool


class BamWriter:
    """Incremental BAM writer: header in its own BGZF block, then the record bytes handed to write() cut into blocks of
    at most 65280 payload bytes (compressed on a thread pool, written in order), EOF marker on close()."""

    def __init__(self, path: str, header_text: str, refs: Sequence[Tuple[str, int]] = (), level: int = 6):
        self.path, self.level = path, level
        self._f = open(path, "wb")
        self._f.write(_bgzf_block(bam_header_bytes(header_text, refs), level))
        self._pending = bytearray()

    def write(self, records) -> None:
        self._pending += memoryview(records)
        n_full = len(self._pending) // _BGZF_MAX_PAYLOAD * _BGZF_MAX_PAYLOAD
        if n_full == 0:
            return
        mv = memoryview(self._pending)
        payloads = [bytes(mv[o:o + _BGZF_MAX_PAYLOAD]) for o in range(0, n_full, _BGZF_MAX_PAYLOAD)]
        del mv
        del self._pending[:n_full]
        level = self.level
        if len(payloads) > 2:
            blocks = _pool().map(lambda p: _bgzf_block(p, level), payloads)
        else:
            blocks = (_bgzf_block(p, level) for p in payloads)
        for b in blocks:
            self._f.write(b)

    def close(self) -> None:
        if self._f is None:
            return
        if self._pending:
            self._f.write(_bgzf_block(bytes(self._pending), self.level))
        self._f.write(_BGZF_EOF)
        self._f.close()
        self._f = None


def _bgzf_blocks(raw, path: str):
    """(compressed start, compressed end, isize, crc) of every BGZF block of the file image."""
    blocks = []
    o, n = 0, len(raw)
    while o < n:
        if raw[o:o + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("%s: not a BGZF block at byte %d" % (path, o))
        (xlen,) = struct.unpack_from("<H", raw, o + 10)
        x, xend, bsize = o + 12, o + 12 + xlen, None
        while x < xend:
            si1, si2, slen = raw[x], raw[x + 1], struct.unpack_from("<H", raw, x + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from("<H", raw, x + 4)[0] + 1
            x += 4 + slen
        if bsize is None:
            raise ValueError("%s: BGZF block without BC subfield" % path)
        crc, isize = struct.unpack_from("<II", raw, o + bsize - 8)
        blocks.append((xend, o + bsize - 8, isize, crc, o))
        o += bsize
    return blocks


def bgzf_inflate(path: str, threads: int = 0) -> memoryview:
    """Concatenated payload of every BGZF block of ``path``.  Blocks are independent deflate
    streams (<= 64 KiB each), so they are inflated by a thread pool straight into one buffer
    (zlib releases the GIL); threads = 0: all host cores (at most 32)."""
    with open(path, "rb") as f:
        raw = f.read()
    blocks = _bgzf_blocks(raw, path)
    offs = [0]
    for b in blocks:
        offs.append(offs[-1] + b[2])
    out = bytearray(offs[-1])
    view = memoryview(out)

    def work(lo: int, hi: int) -> None:
        for i in range(lo, hi):
            c0, c1, isize, crc, o = blocks[i]
            data = zlib.decompress(raw[c0:c1], -15) if isize else b""
            if len(data) != isize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
                raise ValueError("%s: BGZF block CRC/size mismatch at byte %d" % (path, o))
            view[offs[i]:offs[i] + isize] = data
    n_thr = threads or min(32, os.cpu_count() or 1)
    step = 32
    if n_thr <= 1 or len(blocks) <= step:
        work(0, len(blocks))
    else:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(n_thr) as pool:
            for fut in [pool.submit(work, lo, min(lo + step, len(blocks))) for lo in range(0, len(blocks), step)]:
                fut.result()
    return view


def parse_bam_header(data) -> Tuple[str, List[Tuple[str, int]], int]:
    """-> (header text, refs [(name, length)], offset of the first alignment record) from the
    start of an inflated BAM stream; raises IndexError/struct.error when `data` is too short."""
    if bytes(data[:4]) != b"BAM\1":
        raise ValueError("missing BAM magic")
    (l_text,) = struct.unpack_from("<i", data, 4)
    if 8 + l_text + 4 > len(data):
        raise IndexError("header text incomplete")
    text = bytes(data[8:8 + l_text]).split(b"\0", 1)[0].decode("latin-1")
    o = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, o)
        if o + 8 + l_name > len(data):
            raise IndexError("reference list incomplete")
        name = bytes(data[o + 4:o + 4 + l_name - 1]).decode("ascii")
        (l_ref,) = struct.unpack_from("<i", data, o + 4 + l_name)
        refs.append((name, l_ref))
        o += 8 + l_name
    return text, refs, o


def read_bam_header(raw, coff, csize) -> Tuple[str, List[Tuple[str, int]], int]:
    """Header of a BAM file image: inflates only as many leading BGZF blocks as the header
    spans (coff / csize: the block table of fuz_host_bgzf_index)."""
    data = b""
    for i in range(len(coff)):
        c0 = int(coff[i])
        data += zlib.decompress(bytes(raw[c0:c0 + int(csize[i])]), -15)
        try:
            return parse_bam_header(data)
        except (IndexError, struct.error):
            continue
    raise ValueError("BAM header is incomplete")


def read_bam_header_of_file(path: str, first: int = 1 << 20) -> Tuple[str, List[Tuple[str, int]]]:
    """(header text, refs) of a BAM file from its first bytes only: the leading BGZF blocks are inflated until the
    header parses; the prefix read grows when the header is longer than it (raw-read BAMs run to hundreds of GB)."""
    size = os.path.getsize(path)
    n = min(size, first)
    while True:
        with open(path, "rb") as f:
            raw = f.read(n)
        data, o = b"", 0
        while o + 18 <= len(raw):
            if raw[o:o + 4] != b"\x1f\x8b\x08\x04":
                raise ValueError("%s: not a BGZF block at byte %d" % (path, o))
            (xlen,) = struct.unpack_from("<H", raw, o + 10)
            x, xend, bsize = o + 12, o + 12 + xlen, None
            while x + 4 <= xend <= len(raw):
                slen = struct.unpack_from("<H", raw, x + 2)[0]
                if raw[x] == 66 and raw[x + 1] == 67:
                    bsize = struct.unpack_from("<H", raw, x + 4)[0] + 1
                x += 4 + slen
            if bsize is None or o + bsize > len(raw):
                break                                           # block cut by the prefix
            data += zlib.decompress(raw[xend:o + bsize - 8], -15)
            try:
                text, refs, _o = parse_bam_header(data)
                return text, refs
            except (IndexError, struct.error):
                pass
            o += bsize
        if n >= size:
            raise ValueError("%s: BAM header is incomplete" % path)
        n = min(size, n * 8)


def read_bam(path: str):
    """-> (header_text, refs [(name, length)], records buffer (memoryview, no copy))."""
    data = bgzf_inflate(path)
    if bytes(data[:4]) != b"BAM\1":
        raise ValueError("%s: missing BAM magic" % path)
    (l_text,) = struct.unpack_from("<i", data, 4)
    text = bytes(data[8:8 + l_text]).split(b"\0", 1)[0].decode("latin-1")
    o = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, o)
    o += 4
    refs = []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, o)
        name = bytes(data[o + 4:o + 4 + l_name - 1]).decode("ascii")
        (l_ref,) = struct.unpack_from("<i", data, o + 4 + l_name)
        refs.append((name, l_ref))
        o += 8 + l_name
    return text, refs, data[o:]


def is_bgzf(path: str) -> bool:
    with open(path, "rb") as f:
        return f.read(4) == b"\x1f\x8b\x08\x04"


# --------------------------------------------------------------------------- FASTA
def read_fasta(path: str) -> Iterator[Tuple[str, str]]:
    """Minimal stand-in for falcon_kit.FastaReader (reference phasing.py:3,490-494):
    yields (header line without '>', sequence).  A record starts at a line that begins with '>'; sequence
    lines are stripped and joined; anything before the first header is ignored."""
    with open(path) as f:
        text = f.read()
    for rec in ("\n" + text).split("\n>")[1:]:
        header, _, body = rec.partition("\n")
        yield header.rstrip("\r\n"), "".join(map(str.strip, body.split("\n")))


def read_fasta_bytes(path: str) -> Iterator[Tuple[str, bytes]]:
    """read_fasta with the sequences as bytes (no text decoding of megabases): same record and line rules."""
    with open(path, "rb") as f:
        data = f.read()
    # a record starts at a '>' that opens a line
    starts = [0] if data[:1] == b">" else []
    p = data.find(b"\n>")
    while p >= 0:
        starts.append(p + 1)
        p = data.find(b"\n>", p + 1)
    starts.append(len(data) + 1)
    for a, b in zip(starts[:-1], starts[1:]):
        nl = data.find(b"\n", a, b - 1)
        if nl < 0:
            yield data[a + 1:b - 1].rstrip(b"\r\n").decode("latin-1"), b""
            continue
        header = data[a + 1:nl].rstrip(b"\r\n").decode("latin-1")
        body = data[nl + 1:b - 1]
        # one sequence line per record (how assemblers write contigs) needs no split / join
        yield header, (body.strip() if body.find(b"\n") < 0 else b"".join(map(bytes.strip, body.split(b"\n"))))
