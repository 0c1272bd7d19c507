#!/usr/bin/env python
"""Regenerates tests/golden/ from the REFERENCE's own source (oracle/ref_exec.py: the source
text of /root/reference/falcon_unzip/phasing.py patched for Python 3 at run time).  Run in the
build container only; the fixtures travel to the GPU box, the reference does not.

    python scripts/make_golden.py
"""
import gzip
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from falcon_unzip_b200 import bam, synth  # noqa: E402
from oracle import ref_exec  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
FILES = ("variant_pos", "variant_map", "q_id_map", "atable", "phased_variants", "phased_reads")


def main():
    assert ref_exec.available(), "reference tree not mounted"
    mod = ref_exec.load_phasing()
    if os.path.isdir(GOLD):
        shutil.rmtree(GOLD)
    os.makedirs(GOLD)
    tmp = os.path.join(GOLD, "_tmp")
    todo = {k: v for k, v in cases.all_cases().items()}
    sset = synth.generate(synth.CONFIGS["tiny"])
    for c, (name, _l) in enumerate(sset.refs):
        todo["synth_tiny_%s" % name] = (sset.contig_records(c), sset.ref_seqs[c], name, sset.refs)
    for case, val in sorted(todo.items()):
        if len(val) == 2:
            recs, ref = val
            records, refs = cases.build(recs, len(ref))
            ctg = cases.CTG
        else:
            records, ref, ctg, refs = val
            # single-contig BAM: refID 0
            refs = [(ctg, len(ref))]
            arr = np.frombuffer(records, dtype=np.uint8).copy()
            off = [0]
            while off[-1] < len(arr):
                off.append(off[-1] + 4 + int(arr[off[-1]:off[-1] + 4].view("<i4")[0]))
            for o in off[:-1]:
                arr[o + 4:o + 8] = 0
            records = arr.tobytes()
        d = os.path.join(GOLD, case)
        os.makedirs(d)
        bam.write_bam(os.path.join(d, "in.bam"), refs, records, level=9)
        with open(os.path.join(d, "ref.fa"), "w") as f:
            f.write(">%s\n%s\n" % (ctg, ref))
        sam = os.path.join(tmp, case + ".sam")
        os.makedirs(tmp, exist_ok=True)
        with open(sam, "w") as f:
            f.write("\n".join(bam.sam_lines_from_records(records, refs)) + "\n")
        paths = ref_exec.run_phasing_stages(sam, ctg, ref, os.path.join(tmp, case), mod=mod)
        for k in FILES:
            shutil.copy(paths[k], os.path.join(d, k))
        print("%-28s sites %4d vmap %5d atable %5d V+P %4d reads %4d" % (
            case, *[len(open(paths[k]).read().splitlines()) for k in ("variant_pos", "variant_map", "atable",
                                                                       "phased_variants", "phased_reads")]))
    # stage-level fuzz fixtures (inputs + reference outputs of stages 2-4)
    from test_oracle_vs_reference import reference_stages_2_to_4, write_stage_inputs
    for seed in range(6):
        rng = np.random.default_rng(1000 + seed)
        n_sites = int(rng.integers(5, 120))
        _p, _r, rows = cases.random_vmap(rng, n_sites, int(rng.integers(8, 30)), int(rng.integers(30, 200)),
                                         dup_rate=float(rng.choice([0.0, 0.1, 0.3])))
        d = os.path.join(GOLD, "stage_fuzz_%d" % seed)
        write_stage_inputs(d, rows)
        reference_stages_2_to_4(d)
        print("%-28s atable %5d V+P %4d reads %4d" % ("stage_fuzz_%d" % seed, *[
            len(open(os.path.join(d, k)).read().splitlines()) for k in ("atable", "phased_variants", "phased_reads")]))
    shutil.rmtree(tmp)
    with open(os.path.join(GOLD, "README.md"), "w") as f:
        f.write("Golden vectors produced by scripts/make_golden.py from the reference's own source text\n"
                "(falcon_unzip/phasing.py executed under Python 3 with the patch list of SURVEY.md Appendix C).\n"
                "<case>/in.bam + ref.fa are the inputs, the six other files the reference's outputs;\n"
                "stage_fuzz_*/ hold variant_map + q_id_map inputs and the outputs of stages 2-4.\n")


if __name__ == "__main__":
    main()
