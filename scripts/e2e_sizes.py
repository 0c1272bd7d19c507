import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from falcon_unzip_b200 import engine, synth
eng = engine.get_engine(0)
for name in ("tiny", "c1", "c2"):
    sset = synth.generate_parallel(synth.CONFIGS[name]) if name == "c2" else synth.generate(synth.CONFIGS[name])
    pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs], pin=True, assign_qids=False)
    caps = engine.default_caps(int(pb.ctg_len.sum()), pb.n_rec)
    host_out = engine.alloc_host_outputs(caps, pin=True)
    r = eng.phase_host(pb, caps, host_out)
    r = eng.phase_host(pb, caps, host_out)
    ts = []
    for _ in range(10):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = eng.phase_host(pb, caps, host_out)
        ts.append(time.perf_counter() - t0)
    print(name, "rec MB %.1f" % (len(pb.records) / 1e6), "e2e ms min %.3f med %.3f" % (1e3 * min(ts), 1e3 * float(np.median(ts))), "h2d", r.h2d_bytes, "d2h", r.d2h_bytes)
