#!/usr/bin/env python
"""bench.py -- aligned bases/s through pileup + het-call + association + phasing +
read assignment (BASELINE.json metric) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c5|c3|c2|c4] [--impl reference]

A step = one pass of the whole hot path over ONE FIXED contig set: BASELINE.json configs[4] by default
(c5: synthetic human-chr1-scale diploid, 125 contigs x 2 Mb = 250 Mb, 60x, 15 kb reads -- the configuration the
metric "... at 1/2/4/8 B200" is quoted on); --config c3 (2000 x 67.5 kb, 50x), c2 (20 x 250 kb, 40x) and c4
(raw-read tracking, rr_hctg_track) are the other configs.  The contig set is dealt to the N ranks by
shard.assign_contigs (longest-processing-time first, reference unzip.py:231-281: one phasing job per contig), every
rank generates and phases ITS contigs only: STRONG scaling, no data-path collective.  A rank cuts its contigs into
device batches (engine.BatchPacker) and runs fuz_phase_batch over each: all four stages, decoded BAM records
resident in HBM.  Time = max over ranks, value = aligned bases of the whole set / that time.  Rank 0 prints ONE
JSON line.  Inside the run, a sample of every rank's contigs is checked byte for byte (the six output files) against
the CPU oracle ("parity_checked").

--impl reference times the CPU restatement of the reference's algorithm (oracle/, C port; the reference itself is
Python 2 + samtools and cannot run on the GPU box) with one thread per contig on all host cores, each step a bounded
sample of the same workload.  When the reference tree is mounted ($FALCON_UNZIP_REF, build container only) the
reference's OWN source is timed as well (oracle/ref_exec) and reported as cpu_baseline.reference_python.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "aligned_bases_per_sec_pileup_hetcall_phasing"
UNIT = "aligned bases/s"
CONFIG_TEXT = {
    "c2": "BASELINE.json configs[1]: synthetic E. coli-scale diploid",
    "c3": "BASELINE.json configs[2]: synthetic Arabidopsis-scale diploid",
    "c5": "BASELINE.json configs[4]: synthetic human-chr1-scale diploid",
    "c1": "BASELINE.json configs[0]: synthetic diploid single contig",
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="c5")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the host-buffer leg (default min(steps, 3))")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench oracle check (profiling runs)")
    ap.add_argument("--bam", action="store_true", help="add the BAM-ingest leg (BGZF file image -> rows; N=1, small configs)")
    ap.add_argument("--no-bam", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary C2 measurement (N=1)")
    ap.add_argument("--contigs", type=int, default=0, help="use only the first N contigs of the config (named in config)")
    ap.add_argument("--contig-len", type=int, default=0, help="override the contig length of the config (stress cases; named in config)")
    ap.add_argument("--max-batch-mb", type=int, default=6144, help="record bytes per device batch")
    ap.add_argument("--no-bind", action="store_true", help="N>1: do not bind the rank to the CPUs next to its GPU")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="library option for experiments (fuz_set_option), e.g. pdl=0; recorded in config")
    return ap.parse_args()


# --------------------------------------------------------------------------- workload
def workload_cfg(args):
    from falcon_unzip_b200 import synth
    cfg = synth.CONFIGS[args.config]
    if args.contigs:
        cfg = dataclasses.replace(cfg, n_contigs=args.contigs)
    if args.contig_len:
        cfg = dataclasses.replace(cfg, contig_len=args.contig_len)
    return cfg


def rank_contigs(cfg, rank: int, world: int):
    """The rank's share of the fixed contig set: shard.assign_contigs on the expected aligned bases per contig."""
    from falcon_unzip_b200 import shard
    return shard.assign_contigs([float(cfg.contig_len) * cfg.coverage] * cfg.n_contigs, world)[rank]


def sample_ids(cfg, ids, budget_bases: float = 2.5e8, most: int = 16):
    """Contigs of this rank that are checked against the oracle: the first few, bounded by aligned bases."""
    per = max(float(cfg.contig_len) * cfg.coverage, 1.0)
    k = int(max(1, min(most, budget_bases // per, len(ids))))
    return list(ids[:k])


def build_workload(cfg, ids, keep_ids, max_batch_bytes: int, pin: bool, workers: int):
    """Generate the contigs `ids` (synth.generate_contigs_fast: libfuz_synth.so on a pool of threads) and pack them, in order, into
    PreparedBatches.  -> (batches, samples {contig id: dict(batch, local, name, ref_seq, records, rec_off)})."""
    from falcon_unzip_b200 import engine, synth
    pk = engine.BatchPacker(max_batch_bytes)
    batches, samples = [], {}
    parts, names, lens = [], [], []

    def flush():
        if parts:
            batches.append(engine.build_batch(parts, names, lens, pin=pin, assign_qids=False))
            parts.clear(); names.clear(); lens.clear()
            pk.reset()
    # full-size workloads come from the native generator (threads); configs with planted quirks from the numpy one (processes)
    gen = synth.generate_contigs_fast(cfg, ids, workers) if synth.fast_supported(cfg) else synth.generate_contigs(cfg, ids, workers)
    for ci, part in gen:
        name, length = part.refs[0]
        nb, nr = len(part.records), len(part.rec_off) - 1
        if not pk.fits(nb, nr, length):
            flush()
        pk.add(nb, nr, length)
        parts.append((part.records, part.rec_off)); names.append(name); lens.append(length)
        if ci in keep_ids:
            samples[ci] = dict(batch=len(batches), local=len(parts) - 1, name=name, ref_seq=part.ref_seqs[0],
                               records=part.records, rec_off=part.rec_off)
    flush()
    return batches, samples


def algorithmic_bytes(pb) -> dict:
    """bytes the dominant kernel group (pileup + het test) must move for one batch."""
    rec, off = pb.records, pb.rec_off[:-1]
    n_cig = l_seq = np.zeros(0, np.int64)
    if len(off):
        n_cig = rec[off[:, None] + 16 + np.arange(2)[None, :]].copy().view("<u2").reshape(-1).astype(np.int64)
        l_seq = rec[off[:, None] + 20 + np.arange(4)[None, :]].copy().view("<i4").reshape(-1).astype(np.int64)
    total_len = int(np.asarray(pb.ctg_len, np.int64).sum())
    rec_bytes = int((36 + 4 * n_cig + (l_seq + 1) // 2).sum())
    return dict(records=rec_bytes, counts_write=16 * total_len, counts_read=16 * total_len,
                total=rec_bytes + 32 * total_len)


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: NVML polled from a thread every
    millisecond or so (nvidia-smi's own loop mode cannot sample that fast and is only the fallback)."""
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


def ncu_traffic(aligned_bases: int):
    """DRAM bytes per launch group of the dominant kernels: the committed ncu capture (profiles/pileup_traffic.json,
    dram__bytes_read.sum + dram__bytes_write.sum per aligned base, dated) scaled to this launch; (None, why) without it."""
    p = os.path.join(ROOT, "profiles", "pileup_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        per = d.get("dram_bytes_per_aligned_base")
        if per:
            return int(per * aligned_bases), "ncu capture %s (%s), %.3f B/aligned base scaled to this launch" % (
                d.get("capture", "?"), d.get("date", "?"), per)
    return None, "no ncu capture committed for this kernel version"


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


def cpu_run(items, threads: int):
    from concurrent.futures import ThreadPoolExecutor
    from oracle import c_oracle
    c_oracle.lib()
    t0 = time.perf_counter()
    if threads <= 1:
        res = [oracle_contig(x) for x in items]
    else:
        with ThreadPoolExecutor(threads) as ex:     # ctypes releases the GIL inside the C port
            res = list(ex.map(oracle_contig, items))
    dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt, len(items)


def reference_python_leg(cfg, budget_s: float = 40.0):
    """The reference's OWN phasing.py (patched for Python 3, oracle/ref_exec) on one core: only where the reference tree is
    mounted ($FALCON_UNZIP_REF / /root/reference: the build container).  ~1.3e6 bases/s, so a bounded contig."""
    from oracle import ref_exec
    if not ref_exec.available():
        return None
    from falcon_unzip_b200 import bam, synth
    length = int(min(cfg.contig_len, max(20_000, budget_s * 1.25e6 / cfg.coverage)))
    small = dataclasses.replace(cfg, n_contigs=1, contig_len=length, first_contig=0)
    sset = synth.generate(small)
    name = sset.refs[0][0]
    with tempfile.TemporaryDirectory(prefix="fuz_ref_") as d:
        sam = os.path.join(d, "in.sam")
        with open(sam, "w") as f:
            f.write("\n".join(bam.sam_lines_from_records(sset.contig_records(0), sset.refs)) + "\n")
        t0 = time.perf_counter()
        ref_exec.run_phasing_stages(sam, name, sset.ref_seqs[0], os.path.join(d, "ref"))
        dt = time.perf_counter() - t0
    bases, _dt, _n = cpu_run([(sset.records, sset.rec_off)], 1)
    return {"value": bases / dt, "unit": UNIT, "cores": 1, "kind": "reference",
            "interpreter": "CPython %d.%d (reference source patched per SURVEY.md App. C; Python 2 is not installed)" % sys.version_info[:2],
            "sample": "one contig of the config's shape cut to %d bp (%d aligned bases, %.1f s): make_het_call + "
                      "generate_association_table + get_phased_blocks + get_phased_reads from SAM text" % (length, bases, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = workload_cfg(args)
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, cfg.n_contigs))
    per = float(cfg.contig_len) * cfg.coverage
    # bounded sample: one contig per thread, fewer if a step would take more than ~6 s at ~2e7 bases/s/thread
    n_sample = int(max(1, min(threads * max(1, int(6.0 * 2e7 // per)), cfg.n_contigs)))
    ids = list(range(n_sample))
    batches, _s = build_workload(cfg, ids, set(), 1 << 62, pin=False, workers=min(cores, n_sample))
    assert len(batches) == 1
    pb = batches[0]
    from falcon_unzip_b200 import engine
    items = []
    for c in range(pb.n_ctg):
        rec, off, _cro = engine.sub_batch(pb.records, pb.rec_off, pb.ctg_rec_off, c, c + 1)
        items.append((rec, off))
    warm = min(args.warmup, 1)
    for _ in range(warm):
        cpu_run(items, threads)
    times, bases = [], 0
    for _ in range(max(1, args.steps)):
        bases, dt, _n = cpu_run(items, threads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = bases / (ms / 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": max(1, args.steps), "warmup": warm, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(cfg, args, 1, [n_sample], [pb.n_rec], [len(pb.records)], bases, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "each step = the first %d of the %d contigs of the workload (%d aligned bases), one thread per contig, "
                                       "C restatement oracle/phasing_oracle.c; the reference itself is CPython 2 + samtools "
                                       "(~1.25e6 bases/s/core, SURVEY.md section 6)" % (n_sample, cfg.n_contigs, bases)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    try:
        ref = reference_python_leg(cfg)
        if ref:
            line["cpu_baseline"]["reference_python"] = ref
    except Exception as e:                                   # noqa: BLE001 -- optional leg
        line["cpu_baseline"]["reference_python"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    print(json.dumps(line))


def workload_config(cfg, args, world, contigs_per_rank, recs_per_rank, bytes_per_rank, aligned_total, n_batches):
    full = not args.contigs and not args.contig_len
    head = CONFIG_TEXT.get(args.config, args.config) if full else "synthetic %s shape (reduced: --contigs/--contig-len)" % args.config
    return {"workload": "%s, %d contigs x %d bp, %.0fx %d bp reads, %.1f%% het, %.0f%% error"
                        % (head, cfg.n_contigs, cfg.contig_len, cfg.coverage, cfg.mean_read_len, 100 * cfg.het_rate, 100 * cfg.error_rate),
            "contigs_total": cfg.n_contigs, "contigs_per_gpu": contigs_per_rank, "records_per_gpu": recs_per_rank,
            "record_bytes_per_gpu": bytes_per_rank, "aligned_bases_per_step": int(aligned_total),
            "device_batches_per_gpu": n_batches, "seed": cfg.seed,
            "parallelism": "fixed contig set dealt to %d GPU(s) by shard.assign_contigs (LPT), no collective" % world,
            "l2": "inputs (%.0f MB of records on the lightest GPU) exceed the 126 MB L2 and a 256 MiB buffer is overwritten between timed steps"
                  % (min(bytes_per_rank) / 1e6),
            **({"options": list(args.opt)} if args.opt else {})}


# --------------------------------------------------------------------------- in-bench parity
def parity_check(eng, engine, dbs, dos, samples) -> int:
    """The six files of every sample contig from the device rows == the oracle's files, byte for byte."""
    from falcon_unzip_b200 import formats, phasing
    from oracle import c_oracle
    checked = 0
    by_batch = {}
    for ci, s in samples.items():
        by_batch.setdefault(s["batch"], []).append((ci, s))
    with tempfile.TemporaryDirectory(prefix="fuz_parity_") as d:
        for b, items in sorted(by_batch.items()):
            db, do = dbs[b], dos[b]
            eng.phase_batch_async(db, do)                       # the status block holds the last batch only: run this one again
            st = eng.status()
            arrays = do.fetch(st)
            pb = db.pb
            pb.ctg_nq = db.qid_nq.cpu().numpy()[:pb.n_ctg]
            pb.name_first = db.qid_first.cpu().numpy()[:int(pb.ctg_nq.sum())].copy()
            res = engine.PhaseResult(arrays, int(st.n_sites), int(st.n_vmap), int(st.n_atable), int(st.n_reads),
                                     int(st.n_accepted), int(st.aligned_bases))
            sl = formats.contig_slices(res, pb.n_ctg)
            for ci, s in items:
                got = phasing.write_contig_files(res, sl, s["local"], s["name"], s["ref_seq"], pb.qnames(s["local"]),
                                                 os.path.join(d, "gpu"))
                want = c_oracle.run_phasing_stages(s["records"].tobytes(), s["name"], s["ref_seq"], os.path.join(d, "oracle"))
                for k in want:
                    with open(want[k], "rb") as fw, open(got[k], "rb") as fg:
                        if fw.read() != fg.read():
                            raise AssertionError("bench parity: contig %s file %s differs from the oracle" % (s["name"], k))
                checked += 1
    return checked


def bam_ingest_leg(eng, pb, aligned, torch, reps: int = 3):
    """BAM file image (BGZF, zlib level 1) -> rows with everything after the PCIe copy on the device."""
    from falcon_unzip_b200 import bam
    refs = list(zip(pb.ctg_names, [int(x) for x in pb.ctg_len]))
    with tempfile.TemporaryDirectory(prefix="fuz_bench_") as d:
        fn = os.path.join(d, "in.bam")
        rec = pb.records.copy()                                 # refID of a record = contig index inside the file
        refid = np.repeat(np.arange(pb.n_ctg, dtype="<i4"), np.diff(pb.ctg_rec_off))
        rec[(pb.rec_off[:-1, None] + 4 + np.arange(4)[None, :])] = refid.view(np.uint8).reshape(-1, 4)
        bam.write_bam(fn, refs, rec.tobytes(), level=1)
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
    return {"bam_bytes": int(len(image)), "inflated_bytes": int(len(pb.records)), "zlib_level": 1, "records": int(n_rec),
            "ingest_ms": 1e3 * t_ing, "k_bgzf_inflate_ms": k.get("k_bgzf_inflate"),
            "record_index_ms": sum(v for n, v in k.items() if n.startswith("k_bam_")),
            "inflate_out_GBps": len(pb.records) / k["k_bgzf_inflate"] / 1e6 if k.get("k_bgzf_inflate") else None,
            "phase_bam_ms": 1e3 * t_all, "phase_bam_value": aligned / t_all, "unit": UNIT,
            "api": "Engine.phase_bam: fuz_host_bgzf_index + fuz_bgzf_inflate + fuz_bam_index_records + fuz_phase_batch "
                   "(pinned BAM file image in, host row arrays and fixed-width QNAME rows out; best of %d)" % reps}


def disk_to_disk_leg(pb, samples, aligned, device: int, reps: int = 3):
    """SURVEY.md 8(d) "also report": the user-visible call, files in -> files out.  A coordinate-sorted BAM (BGZF, zlib level 1)
    and the FASTA on disk -> phasing.phase_bam -> the six files of every contig on disk (device BGZF inflate + record index
    + four stages + host text formatting + file writes), best of `reps`."""
    from falcon_unzip_b200 import bam, phasing
    refs = list(zip(pb.ctg_names, [int(x) for x in pb.ctg_len]))
    with tempfile.TemporaryDirectory(prefix="fuz_d2d_") as d:
        fn, fa = os.path.join(d, "in.bam"), os.path.join(d, "ref.fa")
        rec = pb.records.copy()                                 # refID of a record = contig index inside the file
        refid = np.repeat(np.arange(pb.n_ctg, dtype="<i4"), np.diff(pb.ctg_rec_off))
        rec[(pb.rec_off[:-1, None] + 4 + np.arange(4)[None, :])] = refid.view(np.uint8).reshape(-1, 4)
        bam.write_bam(fn, refs, rec.tobytes(), level=1)
        seq_of = {s["name"]: s["ref_seq"] for s in samples.values()}
        with open(fa, "w") as f:
            for name, _l in refs:
                f.write(">%s\n%s\n" % (name, seq_of[name]))
        best, n_files = None, 0
        for k in range(reps):
            t0 = time.perf_counter()
            res, files = phasing.phase_bam(fn, fa, os.path.join(d, "out%d" % k), device=device)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            n_files = sum(len(v) for v in files.values())
            assert res.aligned_bases == aligned, "disk path and record path disagree"
        return {"value": aligned / best, "unit": UNIT, "ms": 1e3 * best, "bam_bytes": os.path.getsize(fn), "files_written": n_files,
                "api": "phasing.phase_bam(bam, fasta, base_dir): BGZF BAM + FASTA on disk -> six files per contig on disk (best of %d)" % reps}


# --------------------------------------------------------------------------- GPU arm
def device_leg(eng, engine, torch, stream, dev, batches, steps, warmup, barrier, flush):
    """Upload the batches, size the outputs, run warmup + timed steps.  -> dict of measurements + device objects."""
    dbs, dos, aligned, rows = [], [], 0, dict(sites=0, variant_map=0, atable=0, phased_reads=0, accepted_records=0)
    for pb in batches:
        db = eng.upload(pb)
        caps = engine.default_caps(int(np.asarray(pb.ctg_len, np.int64).sum()), pb.n_rec)
        do, st = eng._retry(caps, 0, lambda d, db=db: eng.phase_batch_async(db, d))     # capacity retry outside the timed region
        dbs.append(db); dos.append(do)
        aligned += int(st.aligned_bases)
        for k, v in (("sites", st.n_sites), ("variant_map", st.n_vmap), ("atable", st.n_atable), ("phased_reads", st.n_reads),
                     ("accepted_records", st.n_accepted)):
            rows[k] += int(v)

    def one_step():
        with torch.cuda.stream(stream):
            flush.fill_(1)                                  # evict L2 between steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for db, do in zip(dbs, dos):
                eng.phase_batch_async(db, do)
            e1.record(stream)
        return e0, e1
    for _ in range(warmup):
        one_step()
    barrier()
    eng.kernel_timing(True)
    launches0 = eng.launch_count()
    barrier()
    t_wall0 = time.perf_counter()
    evs = [one_step() for _ in range(steps)]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count() - launches0
    eng.status()                                             # raises on a device-side error of the last batch
    ms_per_step = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    k_ms, _k_n = eng.get_kernel_timing()
    eng.kernel_timing(False)
    return dict(dbs=dbs, dos=dos, aligned=aligned, rows=rows, ms_per_step=ms_per_step, kern_ms_per_step=k_ms / max(steps, 1),
                launches=launches, wall_ms=1e3 * t_wall / max(steps, 1))


def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = workload_cfg(args)
    ids = rank_contigs(cfg, rank, world)
    keep = set() if args.no_parity and args.no_cpu_baseline else set(sample_ids(cfg, ids))
    workers = max(1, min(len(ids), (os.cpu_count() or 1) // max(world, 1)))
    t_gen0 = time.perf_counter()
    # everything that forks happens before CUDA is initialised
    batches, samples = build_workload(cfg, ids, keep, args.max_batch_mb << 20, pin=False, workers=workers)
    second, second_samples = None, {}
    if world == 1 and not args.no_secondary and args.config != "c2" and not args.contigs and not args.contig_len:
        from falcon_unzip_b200 import synth
        c2 = synth.CONFIGS["c2"]
        second, second_samples = build_workload(c2, list(range(c2.n_contigs)), set(range(c2.n_contigs)), 1 << 62, pin=False,
                                                workers=min(c2.n_contigs, os.cpu_count() or 1))
    t_gen = time.perf_counter() - t_gen0
    alg = [algorithmic_bytes(pb) for pb in batches]

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
    for pb in batches + (second or []):
        engine.pin_batch(pb)                                 # after the binding: first touch places the pages
    eng = engine.Engine(local_rank)
    for kv in args.opt:
        key, val = kv.split("=")
        eng.set_option(key, int(val))
    stream = torch.cuda.Stream(device=dev)
    from falcon_unzip_b200._lib import lib
    lib().fuz_set_stream(eng.ctx, stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    m = device_leg(eng, engine, torch, stream, dev, batches, args.steps, args.warmup, barrier, flush)
    aligned = m["aligned"]

    # ---- in-bench parity: the six files of the sample contigs against the oracle (every rank checks its own sample)
    checked = 0
    if not args.no_parity:
        checked = parity_check(eng, engine, m["dbs"], m["dos"], samples)
    barrier()

    # ---- end to end through the host-buffer C-ABI call (pinned host records in, host rows out), batch after batch,
    # then the host-side gather of what every rank produced (row counts: the files themselves stay with the rank)
    e2e_steps = args.e2e_steps or min(args.steps, 3)
    caps = {k: max(do.caps[k] for do in m["dos"]) for k in ("sites", "vmap", "atable", "reads")}
    host_out = engine.alloc_host_outputs(caps, pin=True)
    h2d = d2h = 0

    def e2e_step():
        nonlocal h2d, d2h
        h2d = d2h = 0
        tot = np.zeros(5, np.int64)
        for pb in batches:
            r = eng.phase_host(pb, caps, host_out)
            h2d += int(r.h2d_bytes); d2h += int(r.d2h_bytes)
            tot += np.asarray([r.n_sites, r.n_vmap, r.n_atable, r.n_reads, r.aligned_bases], np.int64)
        if world > 1:
            t = torch.from_numpy(tot).to(dev)
            parts = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(parts, t)
            tot = torch.stack(parts).sum(0).cpu().numpy()
        return tot
    e2e_step()                                                  # warm-up (staging allocation)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        tot = e2e_step()
    torch.cuda.synchronize(dev)
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    clocks = sampler.stop()                                     # sampled over both timed regions (device-resident + e2e)
    barrier()

    # ---- aggregate over ranks: units summed, time = max over ranks
    vals = torch.tensor([m["ms_per_step"], e2e_ms, m["kern_ms_per_step"]], dtype=torch.float64, device=dev)
    units = torch.tensor([float(aligned), float(checked), float(h2d), float(d2h), float(m["launches"])], dtype=torch.float64, device=dev)
    mine = torch.tensor([m["ms_per_step"], e2e_ms, float(len(ids)), float(sum(pb.n_rec for pb in batches)),
                         float(sum(len(pb.records) for pb in batches)), float(len(batches)),
                         float(binding["first_cpu"]) if binding and binding.get("bound") else -1.0],
                        dtype=torch.float64, device=dev)
    allr = [mine]
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
    ms_max, e2e_ms_max, _k = [float(x) for x in vals.tolist()]
    total_units, total_checked, h2d_all, d2h_all, launches_all = [float(x) for x in units.tolist()]
    per_rank = {"ms_per_step": [round(float(t[0]), 4) for t in allr], "e2e_ms": [round(float(t[1]), 3) for t in allr],
                "contigs": [int(t[2]) for t in allr], "first_cpu": [int(t[6]) for t in allr]}

    # ---- secondary (N=1): BASELINE.json configs[1] (C2) device-timed, and the BAM-ingest leg on it
    secondary, bam_leg = None, None
    if second:
        m2 = device_leg(eng, engine, torch, stream, dev, second, max(args.steps, 10), args.warmup, barrier, flush)
        a2 = algorithmic_bytes(second[0])["total"]
        peak, _src = measured_peak()
        secondary = {"c2": {"workload": CONFIG_TEXT["c2"] + ", 20 contigs x 250000 bp, 40x", "value": m2["aligned"] / (m2["ms_per_step"] / 1e3),
                            "unit": UNIT, "ms_per_step": m2["ms_per_step"], "gpu_launches_per_step": m2["launches"] // max(args.steps, 10),
                            "roofline_frac": a2 / (m2["kern_ms_per_step"] / 1e3) / 1e9 / peak, "kernel_ms": m2["kern_ms_per_step"],
                            "rows": m2["rows"]}}
        try:
            secondary["c2"]["disk_to_disk"] = disk_to_disk_leg(second[0], second_samples, m2["aligned"], local_rank)
        except Exception as e:                               # noqa: BLE001 -- reported, never fatal for the scored line
            secondary["c2"]["disk_to_disk"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
        if args.bam:
            bam_leg = bam_ingest_leg(eng, second[0], m2["aligned"], torch)
    elif args.bam and world == 1 and len(batches) == 1:
        bam_leg = bam_ingest_leg(eng, batches[0], aligned, torch)
    barrier()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and samples:
        items = [(s["records"], s["rec_off"]) for _ci, s in sorted(samples.items())]
        bases, dt, n = cpu_run(items, 1)
        cpu = {"value": bases / dt, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": "the %d sample contigs of rank 0 (%d aligned bases, %.1f s) through oracle/phasing_oracle.c, single thread; "
                         "the reference's own CPython loop measured ~1.25e6 bases/s/core (SURVEY.md section 6)" % (n, bases, dt)}
        try:
            ref = reference_python_leg(cfg, budget_s=30.0)
            if ref:
                cpu["reference_python"] = ref
        except Exception as e:                               # noqa: BLE001 -- optional leg
            cpu["reference_python"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        peak, peak_src = measured_peak()
        alg_total = sum(a["total"] for a in alg)
        achieved = alg_total / (m["kern_ms_per_step"] / 1e3) / 1e9
        traffic, traffic_src = ncu_traffic(aligned)
        line = {"metric": METRIC, "value": total_units / (ms_max / 1e3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                "config": workload_config(cfg, args, world, per_rank["contigs"], [int(t[3]) for t in allr],
                                          [int(t[4]) for t in allr], total_units, [int(t[5]) for t in allr]),
                "parity_checked": int(total_checked),
                "parity": "the six output files of %d sample contig(s) (the first contigs of every rank's share) byte-identical to "
                          "oracle/phasing_oracle.c" % int(total_checked) if total_checked else "skipped (--no-parity)",
                "e2e": {"value": total_units / (e2e_ms_max / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_all),
                        "d2h_bytes_per_step": int(d2h_all), "ms_per_step": e2e_ms_max, "steps": e2e_steps,
                        "api": "fuz_phase_batch_host per device batch (pinned host BAM records in, host row arrays out; header/name/CIGAR/SEQ of "
                               "each record cross PCIe, QUAL and tags do not), then the gather of every rank's row counts"},
                "gpu_launches": int(launches_all),
                "roofline": {"bound": "hbm", "kernel": "pileup group of rank 0 (CIGAR/SEQ decode + per-position counts + het test), timed with CUDA events around the group in every batch",
                             "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "frac_of_nominal_8TBps": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": alg_total, "kernel_ms": m["kern_ms_per_step"], "peak_source": peak_src,
                             "whole_step_frac": (alg_total + 16 * (m["rows"]["variant_map"] + 2 * m["rows"]["atable"] + m["rows"]["sites"] + m["rows"]["phased_reads"]))
                                                / (m["ms_per_step"] / 1e3) / 1e9 / peak,
                             "note": "algorithmic bytes per SURVEY.md 8(d): records once (36+4*n_cigar+ceil(l_seq/2)) + 32 B/position "
                                     "of pileup counts, summed over the batches of rank 0; whole_step_frac adds 16 B per output row and divides by the whole step"},
                "cpu_baseline": cpu,
                "clocks": clocks,
                "per_rank": per_rank,
                **({"secondary": secondary} if secondary else {}),
                **({"bam_ingest": bam_leg} if bam_leg else {}),
                **({"host_binding": binding or {"bound": False, "why": "--no-bind"}} if world > 1 else {}),
                "rows_rank0": m["rows"],
                "setup_s": {"generate": round(t_gen, 1)},
                "wall_ms_per_step_incl_l2_flush": m["wall_ms"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.config == "c4":
        from scripts import bench_rr
        return bench_rr.main_from_bench(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
