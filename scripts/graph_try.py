"""Experiment: one fuz_phase_batch call captured in a CUDA graph and replayed, against plain launches (C2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

def main():
    cfgname = sys.argv[1] if len(sys.argv) > 1 else "c2"
    class A: pass
    a = A(); a.config = cfgname; a.contigs = int(sys.argv[2]) if len(sys.argv) > 2 else 0; a.contig_len = 0
    cfg = bench.workload_cfg(a)
    batches, _s = bench.build_workload(cfg, list(range(cfg.n_contigs)), set(), 1 << 62, pin=False, workers=16)
    import torch
    from falcon_unzip_b200 import engine
    from falcon_unzip_b200._lib import lib
    eng = engine.Engine(0)
    stream = torch.cuda.Stream()
    lib().fuz_set_stream(eng.ctx, stream.cuda_stream)
    pb = batches[0]
    db = eng.upload(pb)
    do, st = eng._retry(engine.default_caps(int(np.asarray(pb.ctg_len, np.int64).sum()), pb.n_rec), 0, lambda d: eng.phase_batch_async(db, d))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def timed(f, n=20):
        ts = []
        for _ in range(n):
            with torch.cuda.stream(stream):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); f(); e1.record(stream)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts)), float(np.min(ts))
    for _ in range(5):
        eng.phase_batch_async(db, do)
    torch.cuda.synchronize()
    plain = timed(lambda: eng.phase_batch_async(db, do))
    print("plain launches: mean %.4f ms min %.4f" % plain, flush=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=stream):
            eng.phase_batch_async(db, do)
    torch.cuda.synchronize()
    graph = timed(lambda: g.replay())
    print("graph replay:   mean %.4f ms min %.4f" % graph, flush=True)
    st2 = eng.status()
    print("rows after replay", st2.n_sites, st2.n_vmap, st2.n_atable, st2.n_reads, "aligned", st2.aligned_bases)

main()
