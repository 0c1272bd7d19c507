"""Where the time of la4falcon.DeviceLines goes (cProfile + kernel profile)."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import engine, la4falcon, synth_rr  # noqa: E402

rr = synth_rr.generate_rr(n_reads=30000, n_ctg=10, ctg_len=1_000_000, mean_len=10_000, n_files=8, seed=20240605)
blobs = ["".join(l + "\n" for l in rr.las_lines[f]).encode("ascii") for f in sorted(rr.las_lines)]
print("bytes", sum(len(b) for b in blobs))
eng = engine.get_engine(0)
la4falcon.DeviceLines(blobs, False)
pr = cProfile.Profile()
pr.enable()
dl = la4falcon.DeviceLines(blobs, False)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
eng.profile(True)
dl = la4falcon.DeviceLines(blobs, False)
eng.sync()
print(eng.profile_report())
