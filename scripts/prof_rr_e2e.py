"""Whole run_track_reads call (text in memory -> rawread_to_contigs on disk): where the time goes."""
import cProfile
import os
import pstats
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import engine, rr_hctg_track, synth_rr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
rr = synth_rr.generate_rr(n_reads=n, n_ctg=max(2, n // 3000), ctg_len=1_000_000, mean_len=10_000, n_files=8, seed=20240605)
d = tempfile.mkdtemp()
p = dict(phased=os.path.join(d, "all_phased_reads"), r2c=os.path.join(d, "read_to_contig_map"), ids=os.path.join(d, "rawread_ids"),
         out=os.path.join(d, "out", "rawread_to_contigs"))
open(p["phased"], "w").write("".join(l + "\n" for l in rr.phased_reads))
open(p["r2c"], "w").write("".join(l + "\n" for l in rr.read_to_contig_map))
open(p["ids"], "w").write(rr.rawread_ids)
blobs = {f: "".join(l + "\n" for l in rr.las_lines[f]).encode("ascii") for f in rr.las_lines}
rr_hctg_track.read_las_lines = lambda db_fn, fn: blobs[fn]
engine.get_engine(0)
run = lambda: rr_hctg_track.run_track_reads(None, p["phased"], p["r2c"], p["ids"], list(blobs), 2500, 40, "db", p["out"])
run()
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
print("lines", sum(len(v) for v in rr.las_lines.values()), "bytes", sum(len(b) for b in blobs.values()))
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
