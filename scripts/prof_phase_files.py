"""BAM + FASTA on disk -> the six files per contig (phasing.phase_bam): where the time goes."""
import cProfile
import os
import pstats
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine, phasing, synth  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
sset = synth.generate_parallel(cfg)
d = tempfile.mkdtemp(prefix="fuz_files_")
fn, fa = os.path.join(d, "in.bam"), os.path.join(d, "ref.fa")
bam.write_bam(fn, sset.refs, sset.records.tobytes())
synth.write_fasta(fa, sset)
engine.get_engine(0)
phasing.phase_bam(fn, fa, os.path.join(d, "warm"))
pr = cProfile.Profile()
pr.enable()
res, _ = phasing.phase_bam(fn, fa, os.path.join(d, "out"))
pr.disable()
print("aligned bases %.1f M, rows: sites %d vmap %d atable %d reads %d" % (res.aligned_bases / 1e6, res.n_sites, res.n_vmap, res.n_atable, res.n_reads))
pstats.Stats(pr).sort_stats("cumulative").print_stats(24)
