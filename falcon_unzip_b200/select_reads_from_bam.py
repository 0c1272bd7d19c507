"""Host mirror of falcon_unzip/select_reads_from_bam.py: the raw-read BAMs are split into
one BAM per contig, every read going to the contig `rawread_to_contigs` ranks first for it.

Same function, arguments, CLI flags and output files (`<sam_dir>/<ctg>.bam`) as the reference
(`select_reads_from_bam.py:8-92`).  What moved to the device: the BGZF inflate of the input BAMs and their record
index (`Engine.ingest_bam`, the BAM ingest of DESIGN.md section 4b), the QNAME of every record (gathered into
fixed-width rows, `Engine.name_rows`) and the partition of whole records by contig (`fuz_gather_records`, file order
kept inside a contig).  The read -> contig table is a sorted byte-string array probed with one vectorised binary
search per file instead of a dict lookup per record.  The output BAMs are compressed on the host (zlib), like the
reference's pysam writer; their compressed bytes depend on the zlib level, their content does not.

Deliberate differences: the header is merged on the text of the input headers (the @RG lines of the later files behind
those of the first, @PG lines dropped, `:43-53`) — pysam re-serialises a parsed header, tag order inside a line may
differ from ours, the set of lines does not; `sam_dir` is created when missing."""
from __future__ import annotations

import argparse
import os
import sys
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import bam, engine

MIN_READS = 20                      # a contig is written when MORE than this many reads pick it (:64)


def read_tables(rawread_to_contigs_fn: str, rawread_ids_fn: str):
    """read_partition: ctg -> set of read names with a rank-0 row for it; read_to_ctgs: read name -> [(score, ctg)]
    (`select_reads_from_bam.py:17-31`)."""
    read_partition: Dict[str, set] = {}
    read_to_ctgs: Dict[str, List[Tuple[int, str]]] = {}
    with open(rawread_ids_fn) as f:
        rid_to_oid = f.read().split("\n")
    with open(rawread_to_contigs_fn) as f:
        for row in f:
            row = row.strip().split()
            if int(row[3]) >= 1:                    # keep top one hits
                continue
            ctg_id = row[1]
            if ctg_id == "NA":
                continue
            o_id = rid_to_oid[int(row[0])]
            read_partition.setdefault(ctg_id, set()).add(o_id)
            read_to_ctgs.setdefault(o_id, []).append((int(row[4]), ctg_id))
    return read_partition, read_to_ctgs


def read_to_selected_ctg(read_partition, read_to_ctgs) -> Dict[str, str]:
    """read name -> the contig its records go to: the first of its sorted (score, ctg) list (`:75-77`), if more than
    MIN_READS reads were partitioned to that contig (`:60-65,78`)."""
    selected = {ctg for ctg, reads in read_partition.items() if len(reads) > MIN_READS}
    out = {}
    for o_id, lst in read_to_ctgs.items():
        ctg = min(lst)[1]
        if ctg in selected:
            out[o_id] = ctg
    return out


def merged_header_text(texts: Sequence[str]) -> str:
    """Header of every output BAM: the first input's, with the @RG lines of the other inputs appended to its own and
    every @PG line dropped (`:43-53`); record types in the order pysam writes them (HD, SQ, RG, then the rest)."""
    def lines_of(t):
        return [ln for ln in t.split("\n") if ln]
    first = lines_of(texts[0])
    rg = [ln for ln in first if ln.startswith("@RG")]
    if len(texts) > 1 and not rg:
        raise KeyError("RG")                        # header['RG'] of the reference
    for t in texts[1:]:
        more = [ln for ln in lines_of(t) if ln.startswith("@RG")]
        if not more:
            raise KeyError("RG")                    # samfile.header['RG']
        rg.extend(more)
    hd = [ln for ln in first if ln.startswith("@HD")]
    sq = [ln for ln in first if ln.startswith("@SQ")]
    rest = [ln for ln in first if ln[:3] not in ("@HD", "@SQ", "@RG", "@PG")]
    return "".join(ln + "\n" for ln in hd + sq + rg + rest)


def partition_records(eng, db, keys: np.ndarray, key_ctg: np.ndarray, n_ctg: int):
    """The records of a device BAM (window) -> (record bytes grouped by contig, file order inside a contig; byte range
    [n_ctg + 1] of every contig).  keys: sorted "S" array of read names, key_ctg: their contig index."""
    names = eng.name_rows(db)
    width = max(keys.dtype.itemsize, names.dtype.itemsize, 1)
    k, n = keys.astype("S%d" % width), names.astype("S%d" % width)
    if len(k) == 0 or len(n) == 0:
        return np.zeros(0, np.uint8), np.zeros(n_ctg + 1, np.int64)
    pos = np.minimum(np.searchsorted(k, n), len(k) - 1)
    hit = k[pos] == n
    rec = np.flatnonzero(hit)
    ctg_of = key_ctg[pos[rec]]
    order = np.argsort(ctg_of, kind="stable")
    data, off = eng.gather_records(db, rec[order])
    bounds = np.searchsorted(ctg_of[order], np.arange(n_ctg + 1), side="left")
    return data, off[bounds]


def partition_file(eng, image, keys: np.ndarray, key_ctg: np.ndarray, n_ctg: int):
    """One input BAM given as a whole file image (tests, small inputs)."""
    return partition_records(eng, eng.ingest_bam(image), keys, key_ctg, n_ctg)


WINDOW_BYTES = 512 << 20        # compressed bytes of an input BAM decoded per device call


def partition_windows(eng, path: str, keys: np.ndarray, key_ctg: np.ndarray, n_ctg: int, window_bytes: int = 0):
    """One input BAM of any size: yields (data, bounds) per window of BGZF blocks (Engine.ingest_bam_windows), so host memory
    and HBM hold one window at a time, like the reference's record loop holds one record (`:55-76`)."""
    for db in eng.ingest_bam_windows(path, window_bytes or WINDOW_BYTES):
        yield partition_records(eng, db, keys, key_ctg, n_ctg)


def select_reads_from_bam(input_bam_fofn_fn, rawread_to_contigs_fn, rawread_ids_fn, sam_dir, device: int = 0, level: int = 6,
                          rank: int = 0, world_size: int = 1, partition_fn=None, window_bytes: int = 0):
    """Write <sam_dir>/<ctg>.bam for every selected contig from the reads of the input BAMs (`:8-89`).

    One process per GPU (world_size > 1): the selected contigs are dealt to the ranks by their number of reads
    (longest-processing-time first); every rank decodes every input BAM on its GPU but gathers, compresses and writes
    only ITS contigs -- the output side (host BGZF compression) is what takes the time, and it divides by contig with
    no exchange between the ranks; a contig's BAM keeps the order of the input files.  partition_fn replaces the
    device call in the CPU tests."""
    print("rawread_ids_fn:", repr(rawread_ids_fn))
    print("rawread_to_contigs_fn:", repr(rawread_to_contigs_fn))
    read_partition, read_to_ctgs = read_tables(rawread_to_contigs_fn, rawread_ids_fn)
    print("num read_partitions:", len(read_partition))
    print("num read_to_ctgs:", len(read_to_ctgs))
    fofn_basedir = os.path.normpath(os.path.dirname(input_bam_fofn_fn))

    def abs_fn(maybe_rel_fn):
        return maybe_rel_fn if os.path.isabs(maybe_rel_fn) else os.path.join(fofn_basedir, maybe_rel_fn)
    with open(input_bam_fofn_fn) as f:
        fns = [abs_fn(row.strip()) for row in f]
    for ctg in sorted(read_partition):
        print("ctg, len:", ctg, len(read_partition[ctg]))
    target = read_to_selected_ctg(read_partition, read_to_ctgs)
    ctgs = sorted(set(target.values()))
    if world_size > 1:
        from . import shard
        mine = shard.assign_contigs([float(len(read_partition[c])) for c in ctgs], world_size)[rank]
        ctgs = [ctgs[i] for i in mine]
        keep = set(ctgs)
        target = {o: c for o, c in target.items() if c in keep}
    ctg_index = {c: i for i, c in enumerate(ctgs)}
    by_name = sorted((o.encode("latin-1"), ctg_index[c]) for o, c in target.items())
    keys = np.array([b for b, _c in by_name], dtype="S") if by_name else np.zeros(0, "S1")
    key_ctg = np.array([c for _b, c in by_name], dtype=np.int64)

    headers = [bam.read_bam_header_of_file(fn) for fn in fns]          # first bytes of every input only
    header_text = merged_header_text([h[0] for h in headers]) if headers else ""
    refs = headers[0][1] if headers else []
    os.makedirs(sam_dir, exist_ok=True)
    eng = engine.get_engine(device) if partition_fn is None else None
    outfile: Dict[str, bam.BamWriter] = {}

    def pieces(fn):
        if partition_fn is not None:                                     # CPU tests: a host stand-in for the device call
            yield partition_fn(eng, np.fromfile(fn, dtype=np.uint8), keys, key_ctg, len(ctgs))
        else:                                                            # one window of one input in memory at a time
            yield from partition_windows(eng, fn, keys, key_ctg, len(ctgs), window_bytes)
    try:
        for fn in fns:
            for data, bounds in pieces(fn):
                for i, ctg in enumerate(ctgs):
                    a, b = int(bounds[i]), int(bounds[i + 1])
                    if a == b:
                        continue
                    if ctg not in outfile:
                        samfile_fn = os.path.join(sam_dir, "%s.bam" % ctg)
                        print("samfile_fn:{!r}".format(samfile_fn), file=sys.stderr)
                        outfile[ctg] = bam.BamWriter(samfile_fn, header_text, refs, level=level)
                    outfile[ctg].write(data[a:b])
    finally:
        for w in outfile.values():
            w.close()
    return sorted(outfile)


def parse_args(argv):
    parser = argparse.ArgumentParser(description="Write ctg.sam files, based on BAM subreads.",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--rawread-to-contigs", type=str,
                        default="./2-asm-falcon/read_maps/dump_rawread_ids/rawread_to_contigs",
                        help="rawread_to_contigs file (from where?)")
    parser.add_argument("--rawread-ids", type=str, default="./2-asm-falcon/read_maps/dump_rawread_ids/rawread_ids",
                        help="rawread_ids file (from where?)")
    parser.add_argument("--sam-dir", type=str, default="./4-quiver/reads", help="Output directory for ctg.sam files")
    parser.add_argument("input_bam_fofn", type=str,
                        help="File of BAM filenames. Paths are relative to dir of FOFN, not CWD.")
    return parser.parse_args(argv[1:])


def main(argv=sys.argv):
    args = parse_args(argv)
    select_reads_from_bam(args.input_bam_fofn, args.rawread_to_contigs, args.rawread_ids, args.sam_dir)


if __name__ == "__main__":
    main()
