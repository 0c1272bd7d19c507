"""One sorted BAM per contig (the reference's layout, unzip.py:90) -> rows: phasing.phase_bam over the list of files
against the same contigs in one BAM.  Usage: bench_bam_files.py [config] [contigs] [zlib level; 0 = stored blocks, for multi-GB cases]"""
import json
import os
import struct
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine, synth  # noqa: E402


def main():
    import dataclasses
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    if len(sys.argv) > 2:
        cfg = dataclasses.replace(cfg, n_contigs=int(sys.argv[2]))
    level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sset = synth.generate_parallel(cfg)
    d = tempfile.mkdtemp(prefix="fuz_files_")
    images = []
    for c, (name, ln) in enumerate(sset.refs):
        rec = np.frombuffer(sset.contig_records(c), np.uint8).copy()
        off = bam.index_records(rec.tobytes())
        rec[(off[:-1, None] + 4 + np.arange(4)[None, :])] = np.frombuffer(struct.pack("<i", 0), np.uint8)
        fn = os.path.join(d, "%s_sorted.bam" % name)
        bam.write_bam(fn, [(name, ln)], rec.tobytes(), level=level)
        images.append(np.fromfile(fn, dtype=np.uint8))
    one = os.path.join(d, "all.bam")
    bam.write_bam(one, sset.refs, sset.records.tobytes(), level=level)
    image = np.fromfile(one, dtype=np.uint8)
    eng = engine.get_engine(0)
    import torch

    def best(f, n=3):
        ts = []
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return min(ts), r
    t1, (r1, _i1) = best(lambda: eng.phase_bam(image))
    tm, (rm, _im) = best(lambda: eng.phase_bam(images))
    assert (r1.n_sites, r1.n_vmap, r1.n_atable, r1.n_reads) == (rm.n_sites, rm.n_vmap, rm.n_atable, rm.n_reads)
    print(json.dumps({"config": cfg.name, "files": len(images), "zlib_level": level, "bam_bytes": int(len(image)), "sites": int(r1.n_sites), "reads": int(r1.n_reads), "aligned_bases": int(r1.aligned_bases), "one_bam_ms": 1e3 * t1,
                      "per_contig_bams_ms": 1e3 * tm, "ms_per_file_overhead": 1e3 * (tm - t1) / len(images)}))


if __name__ == "__main__":
    main()
