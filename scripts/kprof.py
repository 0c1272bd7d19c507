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
    ap.add_argument("--contigs", type=int, default=0)
    ap.add_argument("--contig-len", type=int, default=0)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE")
    a = ap.parse_args()
    a.max_batch_mb = 1 << 20
    cfg = bench.workload_cfg(a)
    batches, _s = bench.build_workload(cfg, list(range(cfg.n_contigs)), set(), 1 << 62, pin=False,
                                       workers=min(cfg.n_contigs, os.cpu_count() or 1))
    import torch
    from falcon_unzip_b200 import engine
    eng = engine.Engine(0)
    for kv in a.opt:
        key, val = kv.split("=")
        eng.set_option(key, int(val))
    pb = batches[0]
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
