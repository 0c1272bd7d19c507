"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of falcon_unzip/select_reads_from_bam.py:8-89 — which raw-read
BAM record goes to which per-contig BAM, and the header those BAMs get.  Plain Python over the BAM reader of the host
package (record split only; no device code).  Pinned against the reference's own source, executed with a pysam
stand-in, by tests/test_select_oracle_vs_reference.py (oracle/ref_exec.py: load_select_reads).  Never imported by
falcon_unzip_b200/."""
from __future__ import annotations

import os
from typing import Dict, List, Tuple


def parse_header(text: str) -> dict:
    """SAM header text -> {'HD': {tag: value}, 'RG': [{...}, ...], ...} in the order of appearance (what the reference
    manipulates through pysam: header['RG'].extend(...), header.pop('PG'), select_reads_from_bam.py:47-50)."""
    out: dict = {}
    for ln in text.split("\n"):
        if not ln:
            continue
        f = ln.split("\t")
        typ = f[0][1:]
        if typ == "CO":
            out.setdefault("CO", []).append("\t".join(f[1:]))
            continue
        d = {}
        for kv in f[1:]:
            k, _, v = kv.partition(":")
            d[k] = v
        if typ == "HD":
            out["HD"] = d
        else:
            out.setdefault(typ, []).append(d)
    return out


def select(input_bam_fofn_fn: str, rawread_to_contigs_fn: str, rawread_ids_fn: str):
    """-> (header dict of the outputs, {ctg: [record bytes, ...] in output order})."""
    from falcon_unzip_b200 import bam
    read_partition: Dict[str, set] = {}
    read_to_ctgs: Dict[str, List[Tuple[int, str]]] = {}
    rid_to_oid = open(rawread_ids_fn).read().split("\n")                     # :17
    with open(rawread_to_contigs_fn) as f:                                   # :18-31
        for row in f:
            row = row.strip().split()
            if int(row[3]) >= 1:
                continue
            if row[1] == "NA":
                continue
            o_id = rid_to_oid[int(row[0])]
            read_partition.setdefault(row[1], set()).add(o_id)
            read_to_ctgs.setdefault(o_id, []).append((int(row[4]), row[1]))
    base = os.path.normpath(os.path.dirname(input_bam_fofn_fn))               # :36-41
    fns = [r.strip() if os.path.isabs(r.strip()) else os.path.join(base, r.strip()) for r in open(input_bam_fofn_fn)]
    header = None                                                            # :42-53
    for fn in fns:
        h = parse_header(bam.read_bam(fn)[0])
        if header is None:
            header = h
        else:
            header["RG"].extend(h["RG"])
    if header is not None:
        header.pop("PG", None)
    selected = {c for c in read_partition if len(read_partition[c]) > 20}    # :58-65
    out: Dict[str, List[bytes]] = {}
    for fn in fns:                                                           # :69-86
        _t, _refs, recs = bam.read_bam(fn)
        buf = bytes(recs)
        off = bam.index_records(buf)
        for i in range(len(off) - 1):
            rec = buf[off[i]:off[i + 1]]
            l_name = rec[12]
            name = rec[36:36 + l_name - 1].decode("latin-1")
            if name not in read_to_ctgs:
                continue
            ctg = sorted(read_to_ctgs[name])[0][1]
            if ctg not in selected:
                continue
            out.setdefault(ctg, []).append(rec)
    return header, out
