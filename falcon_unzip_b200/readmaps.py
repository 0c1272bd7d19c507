"""Drop-in mirrors of the id joins either side of the phasing / tracking path:

  get_phasing_readmap        reference falcon_unzip/phasing_readmap.py:8-51   phased_reads -> rid_to_phase.<ctg>
  get_rid_to_phase_all       reference falcon_unzip/unzip.py:303-314          concatenation -> rid_to_phase.all
  generate_read_to_hctg_map  reference falcon_unzip/get_read_hctg_map.py:12-59  contig edges -> read_to_contig_map

They are joins over a few 10^5 rows (I/O bound, no kernel); what makes them part of the parity contract is
that their row ORDER is the iteration order of CPython-2 dicts / sets (str keys in phasing_readmap.py:48,
tuple keys and sets of contig names in get_read_hctg_map.py:55-58), reproduced here with the table emulation of
py2compat.  Same arguments, same files, same bytes.
"""
from __future__ import annotations

import argparse
import logging
import os
import sys
from typing import Dict, Tuple

from . import py2compat


def fn(p):
    return getattr(p, "path", p)


# --------------------------------------------------------------------------- phasing_readmap.py
def get_phasing_readmap(args) -> None:
    """reference phasing_readmap.py:8-51 (args: phased_reads, read_map_dir, ctg_id, base_dir)."""
    the_ctg_id = args.ctg_id
    with open(os.path.join(args.read_map_dir, "dump_rawread_ids", "rawread_ids")) as f:
        rid_to_oid = f.read().split("\n")
    with open(os.path.join(args.read_map_dir, "dump_pread_ids", "pread_ids")) as f:
        pid_to_fid = f.read().split("\n")
    rid_to_phase: Dict[str, Tuple[int, int]] = {}
    with open(args.phased_reads) as f:
        for row in f:
            row = row.strip().split()
            rid_to_phase[row[6]] = (int(row[2]), int(row[3]))
    arid_to_phase: Dict[str, Tuple[int, int]] = {}
    with open(os.path.join(args.read_map_dir, "pread_to_contigs")) as f:
        for row in f:
            row = row.strip().split()
            if not row[1].startswith(the_ctg_id):
                continue
            if int(row[3]) != 0:                                   # not the best hit
                continue
            fid = pid_to_fid[int(row[0])]
            o_id = rid_to_oid[int(fid.split("/")[1]) // 10]        # py2 int division (:23)
            arid_to_phase["%09d" % int(row[0])] = rid_to_phase.get(o_id, (-1, 0))
    os.makedirs(args.base_dir or ".", exist_ok=True)
    out = os.path.join(args.base_dir, "rid_to_phase.%s" % the_ctg_id)
    with open(out + ".tmp", "w") as f:
        for arid in py2compat.str_dict_order(arid_to_phase):       # dict iteration order of CPython 2 (:48)
            phase = arid_to_phase[arid]
            f.write("%s %s %d %d\n" % (arid, the_ctg_id, phase[0], phase[1]))
    os.replace(out + ".tmp", out)


def parse_args_phasing_readmap(argv):
    parser = argparse.ArgumentParser(description="mapping internal daligner read id to phase block and phase",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--phased_reads", type=str, help="path to read vs. phase map", required=True)
    parser.add_argument("--read_map_dir", type=str, help="path to the read map directory", required=True)
    parser.add_argument("--ctg_id", type=str, help="contig identifier in the bam file", required=True)
    parser.add_argument("--base_dir", type=str, default="./", help="the output base_dir, default to current working directory")
    return parser.parse_args(argv[1:])


def main_phasing_readmap(argv=sys.argv):
    logging.basicConfig()
    get_phasing_readmap(parse_args_phasing_readmap(argv))


# --------------------------------------------------------------------------- unzip.py:303-314
def get_rid_to_phase_all(self) -> None:
    """reference unzip.py:303-314: the per-contig files in sorted path order, concatenated."""
    out_fn = fn(self.rid_to_phase_all)
    inputs_fn = sorted(fn(f) for f in self.inputs.values())
    chunks = []
    for fname in inputs_fn:
        with open(fname) as f:
            chunks.append(f.read())
    with open(out_fn, "w") as out:
        out.write("".join(chunks))


# --------------------------------------------------------------------------- get_read_hctg_map.py
def generate_read_to_hctg_map(self) -> None:
    """reference get_read_hctg_map.py:12-59."""
    with open(fn(self.pread_id_file)) as f:
        pread_did_to_rid = f.read().split("\n")
    with open(fn(self.rawread_id_file)) as f:
        rid_to_oid = f.read().split("\n")
    h_ctg_ids = set()
    with open(fn(self.h_ctg_ids)) as f:
        for row in f:
            h_ctg_ids.add(row.strip())
    # dict keyed by (pid, rid, oid) -> set of contig names, both in insertion order here
    pread_to_contigs: Dict[Tuple[int, int, str], Dict[str, None]] = {}
    for fname in (fn(self.p_ctg_edges), fn(self.h_ctg_edges)):
        with open(fname) as f:
            for row in f:
                row = row.strip().split()
                ctg = row[0]
                if len(ctg.split("_")) > 1 and ctg not in h_ctg_ids:
                    continue
                for node in (row[1], row[2]):
                    pid = int(node.split(":")[0])
                    rid = int(int(pread_did_to_rid[pid].split("/")[1]) / 10)
                    pread_to_contigs.setdefault((pid, rid, rid_to_oid[rid]), {})[ctg] = None
    keys = list(pread_to_contigs)
    hashes = [py2compat.tuple_hash([py2compat.int_hash(k[0]), py2compat.int_hash(k[1]), py2compat.str_hash(k[2])]) for k in keys]
    out_fn = fn(self.read_to_contig_map)
    os.makedirs(os.path.dirname(out_fn) or ".", exist_ok=True)
    with open(out_fn + ".tmp", "w") as f:
        for k in py2compat.table_order(keys, hashes):              # dict order of tuple keys (:55)
            for ctg in py2compat.str_dict_order(pread_to_contigs[k]):   # list(set) order (:57)
                f.write("%09d %09d %s %s\n" % (k[0], k[1], k[2], ctg))
    os.replace(out_fn + ".tmp", out_fn)


def get_read_hctg_map(asm_dir: str, hasm_dir: str, read_to_contig_map_fn: str) -> None:
    """reference get_read_hctg_map.py:61-85 without the one-task pypeFLOW workflow around it."""
    from types import SimpleNamespace
    generate_read_to_hctg_map(SimpleNamespace(
        rawread_id_file=os.path.join(asm_dir, "read_maps/dump_rawread_ids/rawread_ids"),
        pread_id_file=os.path.join(asm_dir, "read_maps/dump_pread_ids/pread_ids"),
        h_ctg_edges=os.path.join(hasm_dir, "all_h_ctg_edges"), p_ctg_edges=os.path.join(hasm_dir, "all_p_ctg_edges"),
        h_ctg_ids=os.path.join(hasm_dir, "all_h_ctg_ids"), read_to_contig_map=read_to_contig_map_fn))


def parse_args_get_read_hctg_map(argv):
    parser = argparse.ArgumentParser(description="generate `read_to_contig_map` (contig id -> internal p-read id -> internal "
                                                 "raw-read id -> original read id)", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument("--output", type=str, default="./4-quiver/read_maps/read_to_contig_map", help="output file")
    return parser.parse_args(argv[1:])


def main_get_read_hctg_map(argv=sys.argv):
    logging.basicConfig()
    args = parse_args_get_read_hctg_map(argv)
    get_read_hctg_map(asm_dir=os.path.abspath("2-asm-falcon"), hasm_dir=os.path.abspath("3-unzip"), read_to_contig_map_fn=args.output)
