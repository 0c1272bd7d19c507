#!/usr/bin/env python
"""Per-kernel device time of one hot-path step, measured in the pipeline with CUDA events
between consecutive launches (warm caches, real clocks) -- complements the cold-cache,
serialised per-launch list of ncu.   python scripts/kprof.py [--config c2] [--steps 5]"""
import argparse
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--replicate", type=int, default=1)
    ap.add_argument("--contigs", type=int, default=0)
    ap.add_argument("--contig-len", type=int, default=0)
    a = ap.parse_args()
    cfg, sset = bench.make_workload(a.config, 0, a.replicate, a.contigs, a.contig_len)
    import torch
    from falcon_unzip_b200 import engine
    eng = engine.Engine(0)
    pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs], rec_off=sset.rec_off,
                              assign_qids=False)
    db = eng.upload(pb)
    do, st = eng._retry(engine.default_caps(int(pb.ctg_len.sum()), pb.n_rec), 0, lambda d: eng.phase_batch_async(db, d))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device)
    for _ in range(3):
        eng.phase_batch_async(db, do)
    eng.sync()
    agg = collections.OrderedDict()
    for _ in range(a.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        eng.profile(True)
        eng.phase_batch_async(db, do)
        for i, (name, ms) in enumerate(eng.profile_report()):
            agg.setdefault((i, name), []).append(ms)
        eng.profile(False)
    total = 0.0
    print("%-3s %-22s %9s" % ("#", "kernel", "mean us"))
    for (i, name), v in agg.items():
        m = 1e3 * sum(v) / len(v)
        total += m
        print("%-3d %-22s %9.1f" % (i, name, m))
    print("total %.1f us per step (%d launches), %.1f G aligned bases/s" % (
        total, len(agg), int(st.aligned_bases) / total / 1e3))


if __name__ == "__main__":
    main()
