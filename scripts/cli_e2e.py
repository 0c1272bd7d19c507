"""Whole user path for one configuration: BAM (BGZF) + FASTA on disk -> the six files per contig.
Prints where the time goes (inflate, record index / q_ids, GPU call, text formatting + writing).
Usage: cli_e2e.py [config] [contigs]"""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine, formats, phasing, synth  # noqa: E402


def main():
    cfg_name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    cfg = synth.CONFIGS[cfg_name]
    if len(sys.argv) > 2:
        import dataclasses
        cfg = dataclasses.replace(cfg, n_contigs=int(sys.argv[2]))
    sset = synth.generate_parallel(cfg)
    d = tempfile.mkdtemp(prefix="fuz_cli_")
    bam_fn, fa_fn = os.path.join(d, "in.bam"), os.path.join(d, "ref.fa")
    bam.write_bam(bam_fn, sset.refs, sset.records.tobytes())
    synth.write_fasta(fa_fn, sset)
    print("BAM %.1f MB (%.1f MB of records), %d contigs" % (os.path.getsize(bam_fn) / 1e6, len(sset.records) / 1e6, len(sset.refs)))
    engine.get_engine(0)
    acc = {}

    def timed(mod, name):
        f = getattr(mod, name)

        def w(*a, **k):
            t0 = time.perf_counter()
            try:
                return f(*a, **k)
            finally:
                acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        setattr(mod, name, w)
    timed(engine, "prepare_batch")
    timed(engine.Engine, "phase_host")
    timed(phasing, "write_contig_files")
    timed(formats, "contig_slices")
    for rep in range(2):
        t = [time.perf_counter()]
        _text, refs, recs = bam.read_bam(bam_fn); t.append(time.perf_counter())
        ref_seqs = {n.split()[0]: s.upper() for n, s in bam.read_fasta(fa_fn)}; t.append(time.perf_counter())
        records = np.frombuffer(recs, dtype=np.uint8)
        names = [r[0] for r in refs]
        out = os.path.join(d, "out%d" % rep)
        t0 = time.perf_counter()
        res, files = phasing.phase_contigs(records, names, [ref_seqs[n] for n in names], out)
        t.append(time.perf_counter())
        print("run %d: inflate+split %.3f s | fasta %.3f s | phase_contigs (index, q_ids, GPU, format, write) %.3f s | total %.3f s"
              % (rep, t[1] - t[0], t[2] - t[1], t[3] - t0, t[3] - t[0]))
        print("   inside phase_contigs: " + ", ".join("%s %.3f s" % kv for kv in acc.items()))
        acc.clear()
        print("   rows: sites %d vmap %d atable %d reads %d; aligned bases %.1f M -> %.2f G bases/s end to end"
              % (res.n_sites, res.n_vmap, res.n_atable, res.n_reads, res.aligned_bases / 1e6,
                 res.aligned_bases / (t[3] - t[0]) / 1e9))


if __name__ == "__main__":
    main()
