"""BAM file image -> phased rows, everything after the PCIe copy on the device.
Times the ingest (H2D of the compressed image, k_bgzf_inflate, record index) and the whole
phase_bam call, next to the host path of the same library (zlib thread pool + host record walk).
Usage: bench_bam.py [config] [contigs] [zlib level]   -> one JSON line."""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from falcon_unzip_b200 import bam, engine, synth  # noqa: E402


def main():
    import dataclasses
    import torch
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
    if len(sys.argv) > 2:
        cfg = dataclasses.replace(cfg, n_contigs=int(sys.argv[2]))
    level = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sset = synth.generate_parallel(cfg)
    d = tempfile.mkdtemp(prefix="fuz_bam_")
    fn = os.path.join(d, "in.bam")
    bam.write_bam(fn, sset.refs, sset.records.tobytes(), level=level)
    image_t = torch.from_numpy(np.fromfile(fn, dtype=np.uint8)).pin_memory()
    image = image_t.numpy()
    eng = engine.get_engine(0)
    out = {"config": cfg.name, "contigs": cfg.n_contigs, "zlib_level": level, "bam_bytes": len(image), "record_bytes": len(sset.records)}

    def wall(f, n):
        ts = []
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        return min(ts), r
    t, db = wall(lambda: eng.ingest_bam(image), 4)
    out["ingest_ms"] = t * 1e3
    out["n_rec"] = db.n_rec
    assert np.array_equal(db.records(), np.asarray(sset.records)), "inflated stream differs"
    eng.ingest_bam(image, profile=True)
    eng.sync()
    out["kernels_us"] = {k: round(v * 1e3, 1) for k, v in eng.profile_report()}
    eng.profile(False)
    k_ms = out["kernels_us"].get("k_bgzf_inflate", 0) / 1e3
    if k_ms:
        out["inflate_GBps_out"] = len(sset.records) / k_ms / 1e6
        out["inflate_GBps_in"] = len(image) / k_ms / 1e6
    del db
    t, (res, _info) = wall(lambda: eng.phase_bam(image), 3)
    out["phase_bam_ms"] = t * 1e3
    out["aligned_bases"] = res.aligned_bases
    out["phase_bam_bases_per_s"] = res.aligned_bases / t
    # host path of the same library: zlib over all cores, sequential record walk
    t0 = time.perf_counter()
    _text, _refs, recs = bam.read_bam(fn)
    t1 = time.perf_counter()
    engine.index_records(np.frombuffer(recs, dtype=np.uint8))
    t2 = time.perf_counter()
    out["host_inflate_ms"] = (t1 - t0) * 1e3
    out["host_index_ms"] = (t2 - t1) * 1e3
    out["host_threads"] = min(32, os.cpu_count() or 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
