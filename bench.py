#!/usr/bin/env python
"""bench.py -- aligned bases/s through pileup + het-call + association + phasing +
read assignment (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2] [--impl reference]

A step = one pass of the whole hot path (fuz_phase_batch: all four stages, every contig of
the workload) over device-resident decoded BAM records.  Workload at N=1: BASELINE.json
configs[1] (synthetic E. coli-scale diploid, 20 contigs x 250 kb, 40x); with N ranks every
rank phases its own, differently seeded copy of that workload (contigs are independent:
weak scaling, no data-path collective).  Prints ONE JSON line on rank 0.

--impl reference times the CPU restatement of the reference's algorithm (oracle/, C port;
the reference itself is Python 2 + samtools and cannot run on the GPU box) with one thread
per contig on all host cores, on the same workload.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned_bases_per_sec_pileup_hetcall_phasing"
UNIT = "aligned bases/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c2")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default min(steps, 5))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-bam", action="store_true", help="skip the BAM-ingest leg (BGZF file image -> rows, N=1 only)")
    ap.add_argument("--replicate", type=int, default=1, help="repeat the workload's contigs (named in config)")
    ap.add_argument("--contigs", type=int, default=0, help="use only the first N contigs of the config (named in config)")
    ap.add_argument("--contig-len", type=int, default=0, help="override the contig length of the config (stress cases; named in config)")
    ap.add_argument("--no-bind", action="store_true", help="N>1: do not bind the rank to the CPUs next to its GPU")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="library option for experiments (fuz_set_option), e.g. pdl=0; recorded in config")
    return ap.parse_args()


def make_workload(name: str, rank: int, replicate: int = 1, contigs: int = 0, contig_len: int = 0):
    from falcon_unzip_b200 import synth
    cfg = synth.CONFIGS[name]
    if contigs:
        cfg = dataclasses.replace(cfg, n_contigs=contigs)
    if contig_len:
        cfg = dataclasses.replace(cfg, contig_len=contig_len)
    cfg = dataclasses.replace(cfg, first_contig=rank * cfg.n_contigs * replicate, n_contigs=cfg.n_contigs * replicate)
    return cfg, synth.generate_parallel(cfg)


def algorithmic_bytes(sset) -> dict:
    """bytes the dominant kernel (pileup + het test) must move."""
    rec, off = sset.records, sset.rec_off[:-1]
    n_cig = rec[off[:, None] + 16 + np.arange(2)[None, :]].copy().view("<u2").reshape(-1).astype(np.int64)
    l_seq = rec[off[:, None] + 20 + np.arange(4)[None, :]].copy().view("<i4").reshape(-1).astype(np.int64)
    total_len = int(sum(l for _n, l in sset.refs))
    rec_bytes = int((36 + 4 * n_cig + (l_seq + 1) // 2).sum())
    return dict(records=rec_bytes, counts_write=16 * total_len, counts_read=16 * total_len,
                total=rec_bytes + 32 * total_len)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every
    millisecond or so (the timed region is a few milliseconds long; nvidia-smi's own loop mode
    cannot sample that fast and is only the fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.p = None
        self.thread = None
        self.sm, self.smax, self.reasons, self.run = [], [], set(), True
        try:
            import threading
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            bits = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
            reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            self.smax.append(float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)))

            def poll():
                while self.run:
                    try:
                        self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                        r = int(reasons_fn(h))
                        for nm, bit in bits.items():
                            if r & bit:
                                self.reasons.add(nm)
                    except Exception:
                        break
                    time.sleep(0.0005)
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self) -> dict:
        if self.thread is not None:
            self.run = False
            self.thread.join(timeout=2)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.smax) if self.smax else None,
                    "samples": len(self.sm), "reasons": sorted(self.reasons), "source": "nvml"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, smax, reasons = [], [], set()
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "pileup_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


# --------------------------------------------------------------------------- CPU legs
def oracle_contig(args):
    """All four stages of one contig through the C port (array level, no text)."""
    from oracle import c_oracle
    recs, off = args
    qid, _names = c_oracle.assign_qids(c_oracle.record_names(recs, off))
    h = c_oracle.het_call(recs, off, qid)
    t = c_oracle.association_table(h["vm_pos"] + 1, h["vm_allele"], h["vm_qid"])
    b = c_oracle.phased_blocks(t["pos1"], t["pos2"], t["b"], t["ct"])
    r = c_oracle.phased_reads(h["vm_pos"] + 1, h["vm_allele"], h["vm_qid"], b["pid"], b["pos"], b["h"])
    return h["aligned_bases"], len(h["site_pos"]), len(r["qid"])


def contig_inputs(sset):
    out = []
    for c in range(len(sset.refs)):
        idx = np.flatnonzero(sset.rec_ctg == c)
        lo, hi = sset.rec_off[idx[0]], sset.rec_off[idx[-1] + 1]
        out.append((sset.records[lo:hi], sset.rec_off[idx[0]:idx[-1] + 2] - lo))
    return out


def cpu_run(sset, threads: int, contigs=None):
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle
    c_oracle.lib()
    items = contig_inputs(sset)
    if contigs is not None:
        items = items[:contigs]
    t0 = time.perf_counter()
    if threads <= 1:
        res = [oracle_contig(x) for x in items]
    else:
        with ThreadPoolExecutor(threads) as ex:     # ctypes releases the GIL inside the C port
            res = list(ex.map(oracle_contig, items))
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt, len(items)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, sset = make_workload(args.config, 0, args.replicate, args.contigs, args.contig_len)
    cores = os.cpu_count() or 1
    threads = min(cores, len(sset.refs))
    for _ in range(min(args.warmup, 1)):
        cpu_run(sset, threads)
    times, bases = [], 0
    for _ in range(max(1, args.steps)):
        bases, dt, n = cpu_run(sset, threads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = bases / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": min(args.warmup, 1), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(cfg, sset, args, bases),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "whole workload per step, one thread per contig (C restatement oracle/phasing_oracle.c; "
                                       "the reference itself is CPython 2 + samtools, ~1.25e6 bases/s/core per SURVEY.md section 6)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(cfg, sset, args, aligned_bases):
    return {"workload": "BASELINE.json configs[1]: synthetic E. coli-scale diploid, %d contigs x %d bp, %.0fx %d bp reads, "
                        "%.1f%% het, %.0f%% error" % (cfg.n_contigs, cfg.contig_len, cfg.coverage, cfg.mean_read_len,
                                                      100 * cfg.het_rate, 100 * cfg.error_rate)
            if args.config == "c2" and args.replicate == 1 and not args.contigs and not args.contig_len else
            "synthetic %s x%d: %d contigs x %d bp, %.0fx" % (args.config, args.replicate, cfg.n_contigs, cfg.contig_len, cfg.coverage),
            "contigs_per_gpu": cfg.n_contigs, "records_per_gpu": int(len(sset.rec_off) - 1),
            "aligned_bases_per_gpu_step": int(aligned_bases), "record_bytes_per_gpu": int(len(sset.records)),
            "seed": cfg.seed, "parallelism": "contig-sharded x%d, no collective" % args.gpus,
            "l2": "inputs (%.0f MB of records per GPU) exceed the 126 MB L2 and a 256 MiB buffer is overwritten between timed steps"
                  % (len(sset.records) / 1e6),
            **({"options": list(args.opt)} if args.opt else {})}


def bam_ingest_leg(eng, sset, aligned, torch, reps: int = 3):
    """BAM file image (BGZF, zlib level 1) -> rows with everything after the PCIe copy on the device."""
    import tempfile
    from falcon_unzip_b200 import bam
    with tempfile.TemporaryDirectory(prefix="fuz_bench_") as d:
        fn = os.path.join(d, "in.bam")
        bam.write_bam(fn, sset.refs, sset.records.tobytes(), level=1)
        image_t = torch.from_numpy(np.fromfile(fn, dtype=np.uint8)).pin_memory()
    image = image_t.numpy()

    def wall(f):
        best, r = None, None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = f()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return best, r
    t_ing, dbam = wall(lambda: eng.ingest_bam(image))
    n_rec = dbam.n_rec
    eng.ingest_bam(image, profile=True)
    eng.sync()
    k = {name: ms for name, ms in eng.profile_report()}
    eng.profile(False)
    del dbam
    t_all, (res, _info) = wall(lambda: eng.phase_bam(image))
    assert res.aligned_bases == aligned, "BAM path and record path disagree"
    # the layout the reference leaves: one sorted BAM per contig (unzip.py:90-91), each with its own header and refID 0
    import struct
    from falcon_unzip_b200 import bam as bam_mod
    images = []
    with tempfile.TemporaryDirectory(prefix="fuz_bench_") as d:
        for c, (name, ln) in enumerate(sset.refs):
            rec = np.frombuffer(sset.contig_records(c), np.uint8).copy()
            off = bam_mod.index_records(rec.tobytes())
            rec[(off[:-1, None] + 4 + np.arange(4)[None, :])] = np.frombuffer(struct.pack("<i", 0), np.uint8)
            fn = os.path.join(d, "%s_sorted.bam" % name)
            bam_mod.write_bam(fn, [(name, ln)], rec.tobytes(), level=1)
            images.append(torch.from_numpy(np.fromfile(fn, dtype=np.uint8)).pin_memory())
    t_files, (res_f, _i) = wall(lambda: eng.phase_bam([t.numpy() for t in images]))
    assert (res_f.aligned_bases, res_f.n_sites, res_f.n_vmap, res_f.n_atable, res_f.n_reads) == \
        (res.aligned_bases, res.n_sites, res.n_vmap, res.n_atable, res.n_reads), "per-contig BAM files and one BAM disagree"
    return {"bam_bytes": int(len(image)), "inflated_bytes": int(len(sset.records)), "zlib_level": 1, "records": int(n_rec),
            "ingest_ms": 1e3 * t_ing, "k_bgzf_inflate_ms": k.get("k_bgzf_inflate"),
            "record_index_ms": sum(v for n, v in k.items() if n.startswith("k_bam_")),
            "inflate_out_GBps": len(sset.records) / k["k_bgzf_inflate"] / 1e6 if k.get("k_bgzf_inflate") else None,
            "phase_bam_ms": 1e3 * t_all, "phase_bam_value": aligned / t_all, "unit": UNIT,
            "per_contig_files": {"files": len(images), "phase_bam_ms": 1e3 * t_files, "phase_bam_value": aligned / t_files,
                                 "api": "Engine.phase_bam(list of file images): one fuz_bgzf_inflate over the blocks of all files + "
                                        "fuz_bam_index_files + fuz_phase_batch"},
            "api": "Engine.phase_bam: fuz_host_bgzf_index + fuz_bgzf_inflate + fuz_bam_index_records + fuz_phase_batch "
                   "(pinned BAM file image in, host row arrays and fixed-width QNAME rows out; best of %d)" % reps}


# --------------------------------------------------------------------------- GPU arm
def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg, sset = make_workload(args.config, rank, args.replicate, args.contigs, args.contig_len)      # before CUDA init (fork)
    alg = algorithmic_bytes(sset)

    import torch
    import torch.distributed as dist
    from falcon_unzip_b200 import engine, shard
    # one process per GPU: stay on the CPUs next to the GPU so that the page-locked records land in that socket's memory
    binding = shard.bind_to_gpu_cpus(local_rank) if world > 1 and not args.no_bind else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    eng = engine.Engine(local_rank)
    for kv in args.opt:
        key, val = kv.split("=")
        eng.set_option(key, int(val))
    stream = torch.cuda.Stream(device=dev)
    from falcon_unzip_b200._lib import lib
    lib().fuz_set_stream(eng.ctx, stream.cuda_stream)

    # q_ids (phasing.py:47-54) are NOT precomputed: both timed legs assign them on the device
    pb = engine.prepare_batch(sset.records, [r[0] for r in sset.refs], [r[1] for r in sset.refs], rec_off=sset.rec_off,
                              pin=True, assign_qids=False)
    db = eng.upload(pb)
    caps = engine.default_caps(int(pb.ctg_len.sum()), pb.n_rec)
    # size the outputs once (capacity retry outside the timed region)
    do, st = eng._retry(caps, 0, lambda d: eng.phase_batch_async(db, d))
    caps = do.caps
    aligned = int(st.aligned_bases)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_step(timed):
        with torch.cuda.stream(stream):
            flush.fill_(1)                                  # evict L2 between steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            eng.phase_batch_async(db, do)
            e1.record(stream)
        return e0, e1

    for _ in range(args.warmup):
        one_step(False)
    barrier()
    eng.kernel_timing(True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    t_wall0 = time.perf_counter()
    evs = [one_step(True) for _ in range(args.steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - launches0
    st = eng.status()
    step_ms = [a.elapsed_time(b) for a, b in evs]
    ms_per_step = float(np.mean(step_ms))
    k_ms, k_n = eng.get_kernel_timing()
    eng.kernel_timing(False)

    # ---- end to end through the host-buffer C-ABI call (pinned host records in, host rows out)
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    host_out = engine.alloc_host_outputs(caps, pin=True)
    r = eng.phase_host(pb, caps, host_out)                      # warm-up (staging allocation)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = eng.phase_host(pb, caps, host_out)
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop()                                     # sampled over both timed regions (device-resident + e2e)
    barrier()

    # ---- BAM ingest leg (N=1): the workload as a BGZF-compressed BAM file image in pinned host memory ->
    # device inflate + record index + q_ids + the four stages (Engine.phase_bam); reported beside the metric
    bam_leg = None
    if world == 1 and not args.no_bam:
        bam_leg = bam_ingest_leg(eng, sset, aligned, torch)
        barrier()

    # ---- aggregate over ranks: units summed, time = max over ranks
    vals = torch.tensor([ms_per_step, e2e_ms, k_ms / max(k_n, 1)], dtype=torch.float64, device=dev)
    units = torch.tensor([float(aligned)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max, kern_ms = [float(x) for x in vals.tolist()]
    per_rank = None
    if world > 1:                                               # diagnostics: e2e time and CPU binding of every rank
        mine = torch.tensor([e2e_ms, float(binding["first_cpu"]) if binding and binding.get("bound") else -1.0],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"e2e_ms": [round(float(t[0]), 3) for t in allr], "first_cpu": [int(t[1]) for t in allr]}
    total_units = float(units.item())

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        bases, dt, n = cpu_run(sset, 1)
        cpu = {"value": bases / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "whole per-GPU workload once (%d contigs, %d aligned bases, %.1f s) through oracle/phasing_oracle.c, "
                         "single thread; the reference's own CPython loop measured ~1.25e6 bases/s/core (SURVEY.md section 6)"
                         % (n, bases, dt)}
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = alg["total"] / (kern_ms / 1e3) / 1e9
        line = {"metric": METRIC, "value": total_units / (ms_max / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": workload_config(cfg, sset, args, aligned),
                "e2e": {"value": total_units / (e2e_ms_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(r.h2d_bytes),
                        "d2h_bytes_per_step": int(r.d2h_bytes), "ms_per_step": e2e_ms_max, "steps": e2e_steps,
                        "api": "fuz_phase_batch_host (pinned host BAM records in, host row arrays out; header/name/CIGAR/SEQ of each record cross PCIe, QUAL and tags do not)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "k_project + k_pileup_gather (pileup + het test, timed as one group)", "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": ncu_traffic(),
                             "algorithmic_bytes_per_launch": alg["total"], "kernel_ms": kern_ms, "peak_source": peak_src,
                             "note": "algorithmic bytes per SURVEY.md 8(d): records once (36+4*n_cigar+ceil(l_seq/2)) + 32 B/position "
                                     "of pileup counts; the kernels keep the counts in registers and move a 4-bit reference-aligned projection "
                                     "(0.5 B/base written + read) instead"},
                "cpu_baseline": cpu,
                "clocks": clocks,
                **({"bam_ingest": bam_leg} if bam_leg else {}),
                **({"host_binding": {**(binding or {"bound": False, "why": "--no-bind"}), "per_rank": per_rank}} if world > 1 else {}),
                "rows": {"sites": int(st.n_sites), "variant_map": int(st.n_vmap), "atable": int(st.n_atable),
                         "phased_reads": int(st.n_reads), "accepted_records": int(st.n_accepted)},
                "wall_ms_per_step_incl_l2_flush": 1e3 * t_wall / max(args.steps, 1)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
