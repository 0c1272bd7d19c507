"""Where the time of the user-visible call goes: phasing.phase_bam(BAM, FASTA, base_dir) on the C2 workload, BAM + FASTA on
disk -> the six files per contig on disk.  Wall time of a few calls, then cProfile.  Usage: prof_disk_to_disk.py [config]"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, phasing, synth  # noqa: E402


def main():
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    sset = synth.generate_parallel(cfg)
    d = tempfile.mkdtemp(prefix="fuz_d2d_")
    fn, fa = os.path.join(d, "in.bam"), os.path.join(d, "ref.fa")
    bam.write_bam(fn, sset.refs, sset.records.tobytes(), level=1)
    synth.write_fasta(fa, sset)
    for k in range(2):
        phasing.phase_bam(fn, fa, os.path.join(d, "warm%d" % k))
    ts = []
    for k in range(5):
        t0 = time.perf_counter()
        phasing.phase_bam(fn, fa, os.path.join(d, "out%d" % k))
        ts.append(1e3 * (time.perf_counter() - t0))
    print("phase_bam(bam, fasta, base_dir) wall ms:", " ".join("%.1f" % t for t in ts))
    pr = cProfile.Profile()
    pr.enable()
    for k in range(3):
        phasing.phase_bam(fn, fa, os.path.join(d, "prof%d" % k))
    pr.disable()
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(30)
    print(out.getvalue())


if __name__ == "__main__":
    main()
