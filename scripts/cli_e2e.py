"""Whole user path for one configuration: BAM (BGZF) + FASTA on disk -> the six files per contig, once with the BAM
decoded on the device (phasing.phase_bam) and once with the host decoder (bam.read_bam + phasing.phase_contigs).
Prints one JSON line.   Usage: cli_e2e.py [config] [contigs]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine, phasing, synth  # noqa: E402


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
    engine.get_engine(0)
    out = {"config": cfg.name, "contigs": cfg.n_contigs, "bam_bytes": os.path.getsize(bam_fn), "record_bytes": int(len(sset.records))}

    def best(f, n=3):
        ts, r = [], None
        for k in range(n):
            t0 = time.perf_counter()
            r = f(k)
            ts.append(time.perf_counter() - t0)
        return min(ts), r
    t_dev, (res, files) = best(lambda k: phasing.phase_bam(bam_fn, fa_fn, os.path.join(d, "dev%d" % k)))

    def host_path(k):
        _text, refs, recs = bam.read_bam(bam_fn)
        ref_seqs = {n.split()[0]: s.upper() for n, s in bam.read_fasta(fa_fn)}
        names = [r[0] for r in refs]
        return phasing.phase_contigs(np.frombuffer(recs, dtype=np.uint8), names, [ref_seqs[n] for n in names], os.path.join(d, "host%d" % k))
    t_host, (res_h, files_h) = best(host_path, 2)
    same = all(open(files[n][k]).read() == open(files_h[n][k]).read() for n in files for k in files[n])
    out.update(aligned_bases=int(res.aligned_bases), files=sum(len(v) for v in files.values()), identical_to_host_decoded_path=bool(same),
               device_decode_s=t_dev, device_decode_bases_per_s=res.aligned_bases / t_dev,
               host_decode_s=t_host, host_decode_bases_per_s=res.aligned_bases / t_host, host_threads=min(32, os.cpu_count() or 1),
               rows={"sites": res.n_sites, "variant_map": res.n_vmap, "atable": res.n_atable, "phased_reads": res.n_reads})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
