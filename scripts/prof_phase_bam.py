"""Where the host time of Engine.phase_bam goes (BAM file image in pinned memory -> rows + QNAME rows on the host):
cProfile over a few calls on the C2 workload.  Usage: prof_phase_bam.py [config]"""
import cProfile
import io
import os
import pstats
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine, synth  # noqa: E402


def main():
    import torch
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    sset = synth.generate_parallel(cfg)
    fn = os.path.join(tempfile.mkdtemp(prefix="fuz_prof_"), "all.bam")
    bam.write_bam(fn, sset.refs, sset.records.tobytes())
    image = torch.from_numpy(np.fromfile(fn, dtype=np.uint8)).pin_memory().numpy()
    eng = engine.get_engine(0)
    for _ in range(3):
        eng.phase_bam(image)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, info = eng.phase_bam(image)
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    print("phase_bam wall ms:", " ".join("%.2f" % t for t in ts))
    t0 = time.perf_counter()
    n = sum(len(info.qnames(c)) for c in range(info.n_ctg))
    print("qnames -> str: %d names, %.2f ms" % (n, 1e3 * (time.perf_counter() - t0)))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5):
        eng.phase_bam(image)
    pr.disable()
    out = io.StringIO()
    pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(28)
    print(out.getvalue())


if __name__ == "__main__":
    main()
