"""Host-entry timing: whole-buffer copy (host_fetch 0) vs selective record fetch (1)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '/root/repo')
from falcon_unzip_b200 import engine, synth
cfg = synth.CONFIGS['c2']
sset = synth.generate_parallel(cfg)
eng = engine.get_engine(0)
pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs], pin=True)
caps = engine.default_caps(int(pb.ctg_len.sum()), pb.n_rec)
host_out = engine.alloc_host_outputs(caps, pin=True)
ref = None
for mode in (0, 1, 0, 1):
    eng.set_option("host_fetch", mode)
    r = eng.phase_host(pb, caps, host_out)
    ts = []
    for _ in range(8):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        r = eng.phase_host(pb, caps, host_out)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    arr = {k: v.copy() for k, v in r.arrays.items()}
    if ref is None: ref = arr
    same = all(np.array_equal(ref[k], arr[k]) for k in ref)
    print("mode", mode, "ms min %.3f med %.3f" % (1e3 * min(ts), 1e3 * float(np.median(ts))), "h2d", r.h2d_bytes, "same", same)
