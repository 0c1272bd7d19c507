#!/usr/bin/env python
"""Secondary measurement: overlap lines/s through the overlap filter with phase
(fuz_ovlp_filter: phase test, per-read end counts, contained set, per-read best-n selection) next to
the CPU oracle on a sample, plus the host stages around it (text parse, output text).
    python scripts/bench_ovlp.py [--reads 40000] [--steps 10]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=40000)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    from falcon_unzip_b200 import synth_rr
    t0 = time.perf_counter()
    n_ctg = max(2, a.reads // 3300)
    s = synth_rr.generate_ovlp(n_reads=a.reads, n_ctg=n_ctg, ctg_len=480_000, mean_len=8000, n_files=8, seed=20240607)
    gen_s = time.perf_counter() - t0
    n_lines = sum(len(v) for v in s.las_lines.values())
    import torch
    from falcon_unzip_b200 import engine, ovlp_filter_with_phase as ofp
    from oracle import ovlp_oracle
    a2p = {r.split()[0]: tuple(r.split()[1:4]) for r in s.rid_phase_rows}
    p = dict(max_diff=120, max_cov=120, min_cov=1, min_len=2500, bestn=10)      # the reference's own call (unzip.py:153)
    blobs = [("\n".join(v) + "\n").encode() for v in s.las_lines.values()]
    t0 = time.perf_counter()
    L = ofp.Lines(blobs)
    parse_s = time.perf_counter() - t0
    from falcon_unzip_b200 import la4falcon
    la4falcon.DeviceLines(blobs[:1], True)                       # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    DL = la4falcon.DeviceLines(blobs, True)                      # upload + fuz_parse_la4falcon (what run_ovlp_filter does)
    torch.cuda.synchronize()
    dev_parse_s = time.perf_counter() - t0
    assert DL.n == L.n
    del DL
    tab = ofp.PhaseTable(a2p, n_reads=s.n_reads)
    eng = engine.get_engine()
    t0 = time.perf_counter()
    dcols = ofp._upload(L, tab)
    torch.cuda.synchronize()
    h2d_ms = 1e3 * (time.perf_counter() - t0)
    run = lambda: ofp._device_filter(L, tab, p["max_diff"], p["max_cov"], p["min_cov"], p["min_len"], p["bestn"], 3, d=dcols)
    r = run()                                                # warm-up (arena, capacities)
    eng.profile(True)
    run()
    prof = eng.profile_report()
    eng.profile(False)
    k_ms = sum(v for _k, v in prof)
    kern = {}
    for k, v in prof:
        kern[k] = round(kern.get(k, 0.0) + 1e3 * v, 1)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        r = run()
    torch.cuda.synchronize()
    call_ms = 1e3 * (time.perf_counter() - t0) / a.steps      # device-resident columns; includes output allocation and D2H of the results
    sel = ofp._resolve_ties(L, tab, r, p["bestn"])
    t0 = time.perf_counter()
    text = ofp._format(L, tab, sel)
    fmt_s = time.perf_counter() - t0
    # CPU oracle: the whole filter on the first LAS file only (1/8 of the lines), extrapolated per line
    f0 = next(iter(s.las_lines))
    t0 = time.perf_counter()
    ovlp_oracle.run_filter([(f0, s.las_lines[f0])], a2p, **p)
    cpu_s = time.perf_counter() - t0
    print(json.dumps({"metric": "overlap_lines_per_sec_ovlp_filter_with_phase", "value": n_lines / (k_ms / 1e3), "unit": "overlap lines/s",
                      "kernels_ms": k_ms, "call_ms_resident_inputs": call_ms, "h2d_ms_pageable_columns": h2d_ms, "n_lines": n_lines, "n_reads": s.n_reads,
                      "groups": int(r["n_groups"]), "selected_lines": int(len(sel)), "tie_groups": int(r["grp_tie"].sum()),
                      "ignored_reads": int(r["ignore"].sum()), "contained_reads": int(r["contained"].sum()),
                      "kernels_us": kern, "params": p,
                      "host_parse_lines_per_sec": n_lines / parse_s, "device_parse_lines_per_sec_incl_upload": n_lines / dev_parse_s,
                      "text_bytes": sum(len(b) for b in blobs), "host_format_lines_per_sec": len(sel) / max(fmt_s, 1e-9),
                      "output_bytes": len(text),
                      "cpu_baseline": {"value": len(s.las_lines[f0]) / cpu_s, "unit": "overlap lines/s", "cores": 1, "kind": "port",
                                       "sample": "all three stages of oracle/ovlp_oracle.py on 1 of 8 LAS files (%d lines, %.1f s)" % (
                                           len(s.las_lines[f0]), cpu_s)},
                      "generate_s": round(gen_s, 1)}))


if __name__ == "__main__":
    main()
