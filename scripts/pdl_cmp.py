import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from falcon_unzip_b200 import engine, synth
sset = synth.generate_parallel(synth.CONFIGS['c2'])
eng = engine.get_engine(0)
pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs])
caps = engine.default_caps(int(pb.ctg_len.sum()), pb.n_rec)
db = eng.upload(pb); out = eng.alloc_outputs(caps)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for mode in (1, 0, 1, 0):
    eng.set_option("pdl", mode)
    for _ in range(3): eng.phase_batch_async(db, out)
    eng.sync()
    ts = []
    for _ in range(20):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.phase_batch_async(db, out); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b))
    print("pdl", mode, "ms mean %.4f min %.4f" % (np.mean(ts), np.min(ts)))
