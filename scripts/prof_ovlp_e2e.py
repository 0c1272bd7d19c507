"""Whole run_ovlp_filter call (LA4Falcon -mo text in memory -> output text): where the time goes."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import engine, ovlp_filter_with_phase as ofp, synth_rr  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
s = synth_rr.generate_ovlp(n_reads=n, n_ctg=max(2, n // 3300), ctg_len=480_000, mean_len=8000, n_files=8, seed=20240607,
                           dup_frac=float(sys.argv[2]) if len(sys.argv) > 2 else 0.03)
blobs = {f: ("\n".join(v) + "\n").encode() for f, v in s.las_lines.items()}
ofp.read_las_lines = lambda db, fn: blobs[fn]
ofp.arid2phase.clear()
ofp.arid2phase.update({r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows})
engine.get_engine(0)
run = lambda: ofp.run_ovlp_filter(list(blobs), "db", 120, 120, 1, 2500, 10)
run()
pr = cProfile.Profile()
pr.enable()
out = run()
pr.disable()
print("lines", sum(len(v) for v in s.las_lines.values()), "bytes", sum(len(b) for b in blobs.values()), "output bytes", len(out))
pstats.Stats(pr).sort_stats("cumulative").print_stats(20)
