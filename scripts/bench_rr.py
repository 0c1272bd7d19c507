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


# --------------------------------------------------------------------------- bench.py --config c4
def _sample_check(rr, keep, hp_n, hp_len, hp_q, bestn, targets):
    """Device result of a few targets against a direct numpy evaluation of rr_hctg_track.py:59-63,97-105: the kept
    (overlap_len, q) tuples of a target, the bestn largest of them as a multiset."""
    kept = np.flatnonzero(keep)
    tk = rr.t[kept]
    order = np.argsort(tk, kind="stable")
    lo = np.searchsorted(tk[order], targets, "left")
    hi = np.searchsorted(tk[order], targets, "right")
    for t, a, b in zip(targets.tolist(), lo.tolist(), hi.tolist()):
        idx = kept[order[a:b]]
        want = sorted(zip(rr.len[idx].tolist(), rr.q[idx].tolist()))[-bestn:] if bestn else []
        n = int(hp_n[t])
        got = sorted(zip(hp_len[t, :n].tolist(), hp_q[t, :n].tolist()))
        if got != want:
            raise AssertionError("bench parity (c4): target %d keeps %r, expected %r" % (t, got[:5], want[:5]))
    return len(targets)


def main_from_bench(args):
    """bench.py --config c4: BASELINE.json configs[3], overlap lines/s through the tracking kernels (filter + heapq replay +
    contig vote).  N > 1: the LAS files are dealt to the ranks in contiguous blocks, every rank filters its lines, ONE
    all-gather (NCCL) of the kept lines, then every rank replays and votes for the targets t % N == rank."""
    import bench
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from falcon_unzip_b200 import synth_rr
    scale = (args.contigs / 50.0) if args.contigs else 1.0           # --contigs N: N of the 50 contigs (reduced, named)
    n_ctg = args.contigs or 50
    bestn, min_len, n_files = 40, 2500, 48
    t0 = time.perf_counter()
    rr = synth_rr.generate_rr_arrays(total_len=int(100_000_000 * scale), n_ctg=n_ctg, n_files=n_files)
    gen_s = time.perf_counter() - t0
    n_lines = len(rr.q)
    workload = {"workload": "BASELINE.json configs[3]: rr_hctg_track raw-read-to-haplotig tracking, %d Mb primary+haplotigs (%d contigs), "
                            "60x raw reads of 10 kb: %d reads, %d overlap lines in %d LAS files, bestn %d, min_len %d%s"
                            % (int(100 * scale), n_ctg, rr.n_reads, n_lines, n_files, bestn, min_len, "" if not args.contigs else " (reduced: --contigs)"),
                "parallelism": "LAS files dealt to %d GPU(s) in contiguous blocks; filter per rank, one NCCL all-gather of the kept lines, "
                               "replay + vote for the targets t %% N == rank" % world, "seed": 20240605,
                "l2": "inputs (%.0f MB of overlap columns) exceed the 126 MB L2 and a 256 MiB buffer is overwritten between timed steps" % (16 * n_lines / 1e6)}
    if args.impl == "reference":
        if rank != 0:
            return
        from oracle import rr_oracle
        n_s = min(n_lines, 400_000)
        lines = synth_rr.rr_text_lines(rr, 0, n_s)
        rid_to_ctg = {"%09d" % r: set(str(c) for c in rr.rc_ctg[rr.rc_off[r]:rr.rc_off[r + 1]].tolist()) for r in np.flatnonzero(rr.in_map).tolist()}
        rid_to_phase = [(str(rr.ph_ctg[r]), int(rr.ph_block[r]), int(rr.ph_phase[r])) if rr.ph_ctg[r] >= 0 else None for r in range(rr.n_reads)]
        times = []
        for _ in range(max(1, args.steps)):
            t0 = time.perf_counter()
            rr_oracle.tr_stage1(lines, min_len, bestn, rid_to_ctg, rid_to_phase)
            times.append(time.perf_counter() - t0)
        v = n_s / float(np.mean(times))
        print(json.dumps({"impl": "reference", "metric": "overlap_lines_per_sec_rr_hctg_track", "value": v, "unit": "overlap lines/s",
                          "n_gpus": args.gpus, "steps": max(1, args.steps), "warmup": 0, "ms_per_step": 1e3 * float(np.mean(times)),
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                          "config": workload,
                          "cpu_baseline": {"value": v, "unit": "overlap lines/s", "cores": 1, "kind": "port",
                                           "sample": "tr_stage1 of oracle/rr_oracle.py (restatement of rr_hctg_track.py:31-65 in CPython %d.%d) on the "
                                                     "first %d LA4Falcon lines of the workload" % (sys.version_info[0], sys.version_info[1], n_s)},
                          "e2e": {"value": v, "unit": "overlap lines/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    import torch
    import torch.distributed as dist
    from falcon_unzip_b200 import _lib, engine, rr_hctg_track as rrm, shard
    binding = shard.bind_to_gpu_cpus(local_rank) if world > 1 and not args.no_bind else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    eng = engine.Engine(local_rank)
    if os.environ.get("FUZ_GRID_RR"):                      # A/B of the launch shapes of the tracking kernels (option grid_rr)
        eng.set_option("grid_rr", int(os.environ["FUZ_GRID_RR"]))
    stream = torch.cuda.Stream(device=dev)
    lib = _lib.lib()
    lib.fuz_set_stream(eng.ctx, stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # this rank's lines: the files [f0, f1)
    f0, f1 = rank * n_files // world, (rank + 1) * n_files // world
    i0, i1 = int(np.searchsorted(rr.file, f0, "left")), int(np.searchsorted(rr.file, f1, "left"))
    cols = np.stack([rr.q[i0:i1], rr.t[i0:i1], rr.len[i0:i1], rr.tlen[i0:i1], rr.file[i0:i1]]).astype(np.int32)
    h_cols = torch.from_numpy(cols).pin_memory()
    up = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x, dtype=dt)).to(dev)
    tabs = dict(in_map=up(rr.in_map, np.uint8), ph_ctg=up(rr.ph_ctg, np.int32), ph_block=up(rr.ph_block, np.int32),
                ph_phase=up(rr.ph_phase, np.int32), rc_off=up(rr.rc_off, np.int32), rc_ctg=up(rr.rc_ctg, np.int32))
    n_reads, b = rr.n_reads, bestn
    cap_votes = 4 * n_reads
    o = dict(hp_n=torch.zeros(n_reads, dtype=torch.int32, device=dev), hp_len=torch.zeros(n_reads * b, dtype=torch.int32, device=dev),
             hp_q=torch.zeros(n_reads * b, dtype=torch.int32, device=dev), vt_off=torch.zeros(n_reads + 1, dtype=torch.int32, device=dev),
             vt_ctg=torch.zeros(cap_votes, dtype=torch.int32, device=dev), vt_count=torch.zeros(cap_votes, dtype=torch.int32, device=dev),
             vt_score=torch.zeros(cap_votes, dtype=torch.int64, device=dev))

    def track(d_cols, keep, filter_only):
        ri = _lib.RRInput()
        ri.n_ovl, ri.n_reads, ri.min_len, ri.bestn, ri.n_ctg = d_cols.shape[1], n_reads, min_len, bestn, rr.n_ctg_names
        for k, row in zip(("q", "t", "len", "tlen", "file"), range(5)):
            setattr(ri, "d_" + k, d_cols[row].data_ptr())
        for k in tabs:
            setattr(ri, "d_" + k, tabs[k].data_ptr())
        ro = _lib.RROutputs()
        ro.cap_votes = cap_votes
        ro.d_keep = keep.data_ptr()
        for k in o:
            setattr(ro, "d_" + k, o[k].data_ptr())
        eng.set_option("rr_filter_only", 1 if filter_only else 0)
        _lib.check(eng.ctx, lib.fuz_rr_track(eng.ctx, C.byref(ri), C.byref(ro)))

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    d_cols = h_cols.to(dev)
    keep = torch.zeros(max(d_cols.shape[1], 1), dtype=torch.uint8, device=dev)
    state = {}

    def step_resident():
        with torch.cuda.stream(stream):
            if world == 1:
                track(d_cols, keep, False)
                return
            track(d_cols, keep, True)                               # map: the overlap filter on this rank's files
            kept = d_cols[:, keep[:d_cols.shape[1]].bool()].contiguous()
            n_loc = torch.tensor([kept.shape[1]], dtype=torch.int64, device=dev)
            sizes = [torch.zeros_like(n_loc) for _ in range(world)]
            dist.all_gather(sizes, n_loc)
            sizes = [int(s.item()) for s in sizes]
            pad = max(max(sizes), 1)
            send = torch.zeros((5, pad), dtype=torch.int32, device=dev)
            send[:, :kept.shape[1]] = kept
            recv = [torch.zeros_like(send) for _ in range(world)]
            dist.all_gather(recv, send)                             # the one exchange of the path
            allk = torch.cat([r[:, :n] for r, n in zip(recv, sizes)], dim=1)          # rank blocks = file blocks: (file, line) order
            own = allk[:, (allk[1] % world) == rank].contiguous()
            k2 = torch.zeros(max(own.shape[1], 1), dtype=torch.uint8, device=dev)
            track(own, k2, False)                                   # merge + vote for this rank's targets
            state["own"], state["kept_total"] = own, allk.shape[1]

    for _ in range(max(args.warmup, 1)):
        step_resident()
    barrier()
    launches0 = eng.launch_count()
    evs = []
    sampler = bench.ClockSampler(local_rank)
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
        step_resident()
        with torch.cuda.stream(stream):
            e1.record(stream)
        evs.append((e0, e1))
    barrier()
    launches = eng.launch_count() - launches0
    st = eng.status()
    ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    # ---- per-kernel profile of one call (rank 0) for the roofline block
    prof = {}
    if world == 1:
        eng.profile(True)
        with torch.cuda.stream(stream):
            track(d_cols, keep, False)
        prof = {k: round(1e3 * v, 1) for k, v in eng.profile_report()}
        eng.profile(False)
    # ---- parity sample (N = 1: every rank would hold only its targets)
    checked = 0
    if world == 1 and not args.no_parity:
        torch.cuda.synchronize(dev)
        hp_n = o["hp_n"].cpu().numpy()
        with_lines = np.flatnonzero(hp_n > 0)
        targets = with_lines[np.linspace(0, len(with_lines) - 1, 200).astype(np.int64)] if len(with_lines) else with_lines
        checked = _sample_check(rr, keep[:n_lines].cpu().numpy().astype(bool), hp_n, o["hp_len"].cpu().numpy().reshape(n_reads, b),
                                o["hp_q"].cpu().numpy().reshape(n_reads, b), bestn, targets)
    # ---- end to end: pinned host columns in, rows of the vote on the host (per rank: its files / its targets)
    e2e_steps = args.e2e_steps or min(args.steps, 3)
    h2d = d2h = 0

    def step_e2e():
        nonlocal d_cols, h2d, d2h
        with torch.cuda.stream(stream):
            d_cols = h_cols.to(dev, non_blocking=True)
        step_resident()
        with torch.cuda.stream(stream):
            n_votes = int(o["vt_off"][n_reads].item()) if world == 1 else int(o["vt_off"][n_reads].item())
            res = [o["hp_n"].cpu(), o["vt_off"].cpu(), o["vt_ctg"][:n_votes].cpu(), o["vt_count"][:n_votes].cpu(), o["vt_score"][:n_votes].cpu()]
        h2d, d2h = h_cols.numel() * 4, sum(r.numel() * r.element_size() for r in res)
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop()
    barrier()
    vals = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(i1 - i0), float(h2d), float(d2h), float(launches), float(checked)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms_max, e2e_max = [float(x) for x in vals.tolist()]
        tot_lines, h2d_all, d2h_all, launches_all, checked_all = [float(x) for x in sums.tolist()]
        n_targets = int((o["hp_n"] > 0).sum().item())
        table_bytes = sum(t.numel() * t.element_size() for t in tabs.values())
        alg = 16 * (i1 - i0) + 8 * n_targets * bestn + table_bytes
        peak, peak_src = bench.measured_peak()
        achieved = alg / (ms / 1e3) / 1e9
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import rr_oracle
            n_s = min(n_lines, 300_000)
            lines = synth_rr.rr_text_lines(rr, 0, n_s)
            rid_to_ctg = {"%09d" % r: set(str(c) for c in rr.rc_ctg[rr.rc_off[r]:rr.rc_off[r + 1]].tolist()) for r in np.flatnonzero(rr.in_map).tolist()}
            rid_to_phase = [(str(rr.ph_ctg[r]), int(rr.ph_block[r]), int(rr.ph_phase[r])) if rr.ph_ctg[r] >= 0 else None for r in range(rr.n_reads)]
            t0 = time.perf_counter()
            rr_oracle.tr_stage1(lines, min_len, bestn, rid_to_ctg, rid_to_phase)
            dt = time.perf_counter() - t0
            cpu = {"value": n_s / dt, "unit": "overlap lines/s", "cores": 1, "kind": "port",
                   "sample": "tr_stage1 of oracle/rr_oracle.py (restatement of rr_hctg_track.py:31-65) on the first %d lines (%.1f s)" % (n_s, dt)}
        line = {"metric": "overlap_lines_per_sec_rr_hctg_track", "value": tot_lines / (ms_max / 1e3), "unit": "overlap lines/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 1), "ms_per_step": ms_max, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload,
                "parity_checked": int(checked_all),
                "parity": "kept (overlap_len, q) multiset of %d sample targets equals a direct numpy evaluation of rr_hctg_track.py:59-63,97-105 "
                          "(heap ARRAY order and vote rows: tests/test_gpu_rr.py against the oracle)" % int(checked_all),
                "e2e": {"value": tot_lines / (e2e_max / 1e3), "unit": "overlap lines/s", "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                        "ms_per_step": e2e_max, "steps": e2e_steps,
                        "api": "fuz_rr_track through the C ABI: pinned host int32 columns (q, t, len, tlen, file = the parsed LA4Falcon -m lines) in, "
                               "heap sizes + vote rows on the host out"},
                "gpu_launches": int(launches_all),
                "roofline": {"bound": "hbm", "kernel": "fuz_rr_track of rank 0 (k_rr_filter + grouping + k_rr_replay + k_rr_vote), CUDA events around the call",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "algorithmic_bytes_per_launch": int(alg), "kernel_ms": ms, "peak_source": peak_src, "kernels_us": prof,
                             "note": "SURVEY.md 8(d) R1-R3: 16 B per overlap line + 8 B x targets x bestn + the id tables once"},
                "cpu_baseline": cpu, "clocks": clocks,
                "rows_rank0": {"kept_lines": int(st.reserved[3]), "vote_rows": int(st.reserved[1]), "targets": n_targets},
                "setup_s": {"generate": round(gen_s, 1)},
                **({"host_binding": binding or {"bound": False}} if world > 1 else {})}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
