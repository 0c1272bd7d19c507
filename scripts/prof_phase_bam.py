import os, sys, tempfile, time, cProfile, pstats
import numpy as np
sys.path.insert(0, '/root/repo')
from falcon_unzip_b200 import bam, engine, synth
import torch
cfg = synth.CONFIGS["c2"]
sset = synth.generate_parallel(cfg)
fn = os.path.join(tempfile.mkdtemp(), "in.bam")
bam.write_bam(fn, sset.refs, sset.records.tobytes(), level=1)
image = torch.from_numpy(np.fromfile(fn, dtype=np.uint8)).pin_memory().numpy()
eng = engine.get_engine(0)
for _ in range(2): eng.phase_bam(image)
pr = cProfile.Profile(); pr.enable()
for _ in range(3): eng.phase_bam(image)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
