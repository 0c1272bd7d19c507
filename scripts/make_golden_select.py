"""Golden vector of the per-contig read selection, made from the REFERENCE's own select_reads_from_bam.py (executed by
oracle/ref_exec.py with a pysam stand-in; build container only): tests/golden/select_seed7.json.  The inputs are not
stored: tests/select_cases.make_case(seed=7) regenerates them; their digests are, so a drifted generator is noticed.

    python scripts/make_golden_select.py
"""
import hashlib
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import select_cases  # noqa: E402
from falcon_unzip_b200 import bam  # noqa: E402
from oracle import ref_exec  # noqa: E402


def input_digests(fofn, r2c, ids):
    base = os.path.dirname(fofn)
    fns = [r.strip() if os.path.isabs(r.strip()) else os.path.join(base, r.strip()) for r in open(fofn)]
    return {"rawread_to_contigs": hashlib.sha256(open(r2c, "rb").read()).hexdigest(),
            "rawread_ids": hashlib.sha256(open(ids, "rb").read()).hexdigest(),
            "bam_records": [hashlib.sha256(bytes(bam.read_bam(fn)[2])).hexdigest() for fn in fns],
            "bam_headers": [bam.read_bam(fn)[0] for fn in fns]}


def main():
    assert ref_exec.available(), "reference tree not mounted"
    seed = 7
    with tempfile.TemporaryDirectory() as d:
        fofn, r2c, ids = select_cases.make_case(d, seed=seed)
        sam_dir = os.path.join(d, "out")
        ref = ref_exec.run_select_reads(fofn, r2c, ids, sam_dir)
        gold = {"made_by": "scripts/make_golden_select.py (reference select_reads_from_bam.py:8-89, pysam stand-in)", "seed": seed,
                "inputs": input_digests(fofn, r2c, ids), "contigs": {}}
        header = None
        for path, (hdr, recs) in sorted(ref.items()):
            ctg = os.path.basename(path)[:-4]
            header = hdr
            gold["contigs"][ctg] = {"records": len(recs), "bytes": sum(len(r) for r in recs),
                                    "sha256": hashlib.sha256(b"".join(recs)).hexdigest(),
                                    "first_names": [r[36:36 + r[12] - 1].decode() for r in recs[:4]]}
        gold["header"] = header
    out = os.path.join(ROOT, "tests", "golden", "select_seed7.json")
    with open(out, "w") as f:
        json.dump(gold, f, indent=1, sort_keys=True)
        f.write("\n")
    print(out, {c: v["records"] for c, v in gold["contigs"].items()})


if __name__ == "__main__":
    main()
