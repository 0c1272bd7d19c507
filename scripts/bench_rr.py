#!/usr/bin/env python
"""Secondary measurement (BASELINE.json config 4 shape, reduced): overlap lines/s through the
raw-read -> haplotig tracking kernels (fuz_rr_track: filter + exact heapq replay + contig vote)
on device-resident int arrays, with the CPU oracle timed beside it on a sample.
    python scripts/bench_rr.py [--reads 60000] [--steps 10]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=60000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--bestn", type=int, default=40)
    a = ap.parse_args()
    from falcon_unzip_b200 import synth_rr
    n_ctg = max(2, a.reads // 3000)
    t0 = time.perf_counter()
    rr = synth_rr.generate_rr(n_reads=a.reads, n_ctg=n_ctg, ctg_len=1_000_000, mean_len=10_000, n_files=8, seed=20240605)
    n_lines = sum(len(v) for v in rr.las_lines.values())
    gen_s = time.perf_counter() - t0
    import torch
    from falcon_unzip_b200 import _lib, engine, rr_hctg_track as rrm
    from oracle import rr_oracle
    rid_to_ctg = {}
    for row in rr.read_to_contig_map:
        _p, rid, _o, ctg = row.split()
        rid_to_ctg.setdefault(rid, rrm.OrderedStrSet()).add(ctg)
    rid_to_phase = rr_oracle.phase_table(rr.phased_reads, rr.rawread_ids)
    tab = rrm._Tables(rid_to_ctg, rid_to_phase, len(rid_to_phase))
    files = sorted(rr.las_lines)
    blobs = ["".join(l + "\n" for l in rr.las_lines[f]).encode("ascii") for f in files]   # what LA4Falcon -m prints
    t0 = time.perf_counter()
    parts = [rrm._parse_lines(b) for b in blobs]
    parse_s = time.perf_counter() - t0
    # the same text parsed on the device (what run_track_reads does): upload + fuz_parse_la4falcon
    from falcon_unzip_b200 import la4falcon
    la4falcon.DeviceLines(blobs[:1], False)                      # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dl = la4falcon.DeviceLines(blobs, False)
    torch.cuda.synchronize()
    dev_parse_s = time.perf_counter() - t0
    assert dl.n == sum(len(p[0]) for p in parts)
    del dl
    q, t, ln, tl = (np.concatenate([p[k] for p in parts]) for k in range(4))
    fidx = np.concatenate([np.full(len(p[0]), i, np.int32) for i, p in enumerate(parts)])
    # one warm call through the product path (allocations, correctness of sizes)
    rrm._track_device(q, t, ln, tl, fidx, tab, 2500, a.bestn)
    eng = engine.get_engine()
    dev = eng.device
    up = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x, dtype=dt)).to(dev)
    d = dict(q=up(q, np.int32), t=up(t, np.int32), len=up(ln, np.int32), tlen=up(tl, np.int32), file=up(fidx, np.int32),
             in_map=up(tab.in_map, np.uint8), ph_ctg=up(tab.ph_ctg, np.int32), ph_block=up(tab.ph_block, np.int32),
             ph_phase=up(tab.ph_phase, np.int32), rc_off=up(tab.rc_off, np.int32), rc_ctg=up(tab.rc_ctg, np.int32))
    n_reads, b = tab.n_reads, a.bestn
    cap_votes = 8 * n_reads
    o = dict(keep=torch.zeros(len(q), dtype=torch.uint8, device=dev), hp_n=torch.zeros(n_reads, dtype=torch.int32, device=dev),
             hp_len=torch.zeros(n_reads * b, dtype=torch.int32, device=dev), hp_q=torch.zeros(n_reads * b, dtype=torch.int32, device=dev),
             vt_off=torch.zeros(n_reads + 1, dtype=torch.int32, device=dev), vt_ctg=torch.zeros(cap_votes, dtype=torch.int32, device=dev),
             vt_count=torch.zeros(cap_votes, dtype=torch.int32, device=dev), vt_score=torch.zeros(cap_votes, dtype=torch.int64, device=dev))
    ri = _lib.RRInput()
    ri.n_ovl, ri.n_reads, ri.min_len, ri.bestn, ri.n_ctg = len(q), n_reads, 2500, b, len(tab.ctg_names)
    for k in d:
        setattr(ri, "d_" + k, d[k].data_ptr())
    ro = _lib.RROutputs()
    ro.cap_votes = cap_votes
    for k in o:
        setattr(ro, "d_" + k, o[k].data_ptr())
    torch.cuda.synchronize()
    lib = _lib.lib()
    for _ in range(3):
        _lib.check(eng.ctx, lib.fuz_rr_track(eng.ctx, C.byref(ri), C.byref(ro)))
    eng.sync()
    eng.profile(True)
    _lib.check(eng.ctx, lib.fuz_rr_track(eng.ctx, C.byref(ri), C.byref(ro)))
    prof = eng.profile_report()
    eng.profile(False)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        _lib.check(eng.ctx, lib.fuz_rr_track(eng.ctx, C.byref(ri), C.byref(ro)))
    eng.sync()
    ms = 1e3 * (time.perf_counter() - t0) / a.steps
    st = eng.status()
    # CPU oracle on the first LAS file (1/8 of the lines)
    f0 = files[0]
    t0 = time.perf_counter()
    rr_oracle.tr_stage1(rr.las_lines[f0], 2500, b, rr_oracle.get_rid_to_ctg(rr.read_to_contig_map), rid_to_phase)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"metric": "overlap_lines_per_sec_rr_hctg_track", "value": n_lines / (ms / 1e3), "unit": "overlap lines/s",
                      "ms_per_step": ms, "n_lines": n_lines, "n_reads": n_reads, "kept_lines": int(st.reserved[3]),
                      "vote_rows": int(st.reserved[1]), "bestn": b, "kernels_us": {k: round(1e3 * v, 1) for k, v in prof},
                      "host_parse_lines_per_sec": n_lines / parse_s, "device_parse_lines_per_sec_incl_upload": n_lines / dev_parse_s,
                      "text_bytes": sum(len(b) for b in blobs),
                      "cpu_baseline": {"value": len(rr.las_lines[f0]) / cpu_s, "unit": "overlap lines/s", "cores": 1, "kind": "port",
                                       "sample": "tr_stage1 of oracle/rr_oracle.py on 1 of 8 LAS files (%d lines, %.1f s)" % (
                                           len(rr.las_lines[f0]), cpu_s)},
                      "generate_s": round(gen_s, 1)}))


if __name__ == "__main__":
    main()
