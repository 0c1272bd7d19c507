"""One device ingest of a synthetic BAM (for ncu captures of k_bgzf_inflate / k_bam_*)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine, synth  # noqa: E402

cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
sset = synth.generate_parallel(cfg)
fn = os.path.join(tempfile.mkdtemp(prefix="fuz_bam_"), "in.bam")
bam.write_bam(fn, sset.refs, sset.records.tobytes(), level=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
eng = engine.get_engine(0)
for _ in range(2):
    db = eng.ingest_bam(np.fromfile(fn, dtype=np.uint8))
print("records", db.n_rec)
