"""TEST INFRASTRUCTURE ONLY (oracle/): run the reference's *own source text* under Python 3.

The reference (falcon_unzip/phasing.py, falcon_unzip/rr_hctg_track.py) is Python 2 and
its dependencies (pypeflow, falcon_kit, samtools, LA4Falcon) are absent, so it cannot
be imported.  This module reads the source text from ``$FALCON_UNZIP_REF`` (default
/root/reference) AT RUN TIME, applies the enumerated patch list of SURVEY.md Appendix C,
stubs the missing modules and ``exec``s the result.  Nothing of the reference is copied
into this repository.  It exists only in the build container (the GPU box has no
/root/reference): it validates oracle/phasing_oracle.c, oracle/rr_oracle.py, oracle/ovlp_oracle.py, oracle/select_oracle.py and
generates the committed fixtures under tests/golden/ (scripts/make_golden.py).

``available()`` is False when the reference tree is missing; callers skip.
"""
from __future__ import annotations

import os
import re
import stat
import sys
import types
from typing import Dict, Iterable, List

from . import py2emu

REF_ROOT = os.environ.get("FALCON_UNZIP_REF", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "falcon_unzip", "phasing.py"))


def _sub(pattern: str, repl, text: str, expect: int, flags: int = 0) -> str:
    out, n = re.subn(pattern, repl, text, flags=flags)
    if n != expect:
        raise RuntimeError("oracle patch %r matched %d sites, expected %d" % (pattern, n, expect))
    return out


def _stub_modules() -> Dict[str, types.ModuleType]:
    mods = {}
    pf = types.ModuleType("pypeflow")
    br = types.ModuleType("pypeflow.simple_pwatcher_bridge")
    br.fn = lambda p: p
    for name in ("PypeProcWatcherWorkflow", "MyFakePypeThreadTaskBase", "makePypeLocalFile",
                 "PypeTask"):
        setattr(br, name, object)
    pf.simple_pwatcher_bridge = br
    fk = types.ModuleType("falcon_kit")
    fr = types.ModuleType("falcon_kit.FastaReader")
    fr.FastaReader = object
    mp = types.ModuleType("falcon_kit.multiproc")

    class Pool:                                   # main-process pool (= --debug, n_core 0)
        def __init__(self, *_a, **_k):
            pass
        imap = staticmethod(map)

        def terminate(self):
            pass
    mp.Pool = Pool
    util = types.ModuleType("falcon_kit.util")
    io = types.ModuleType("falcon_kit.util.io")
    io.LOG = lambda *_a, **_k: None
    io.logstats = lambda *_a, **_k: None
    io.write_nothing = lambda *_a, **_k: None
    io.run_func = lambda args: args[0](*args[1:])
    io.CapturedProcessReaderContext = object
    io.StreamedProcessReaderContext = object
    util.io = io
    fk.FastaReader, fk.multiproc, fk.util = fr, mp, util
    mods.update({"pypeflow": pf, "pypeflow.simple_pwatcher_bridge": br, "falcon_kit": fk,
                 "falcon_kit.FastaReader": fr, "falcon_kit.multiproc": mp,
                 "falcon_kit.util": util, "falcon_kit.util.io": io})
    return mods


def _exec_patched(src: str, name: str, extra: dict) -> types.ModuleType:
    saved = {k: sys.modules.get(k) for k in _stub_modules()}
    stubs = _stub_modules()
    sys.modules.update(stubs)
    try:
        mod = types.ModuleType(name)
        mod.__dict__.update(extra)
        exec(compile(src, name + ".py(patched)", "exec"), mod.__dict__)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod


def load_phasing() -> types.ModuleType:
    """falcon_unzip/phasing.py with Appendix C patches 1-7."""
    with open(os.path.join(REF_ROOT, "falcon_unzip", "phasing.py")) as f:
        src = f.read()
    src = _sub(r"print >>\s*(\w+),\s*(.*)", r"print(\2, file=\1)", src, 10)
    src = _sub(r"\bxrange\b", "range", src, 3)
    src = _sub(r"pos_k = pileup\.keys\(\)\n(\s*)pos_k\.sort\(\)",
               r"pos_k = sorted(pileup.keys())", src, 1)
    src = _sub(r"stdout=subprocess\.PIPE\)", "stdout=subprocess.PIPE, universal_newlines=True)",
               src, 1)
    src = _sub(r"(list[12]) = (vmap\[ \(pos[12], rb[12]\) \])\.items\(\)",
               r'\1 = sorted(\2.items(), key=lambda kv: "ACTG".index(kv[0]))', src, 2)
    src = _sub(r"for r in read_to_variants:", "for r in py27_int_dict_order(read_to_variants):",
               src, 1)
    src = _sub(r"1\.0 \* \(max_-min_\)/len\(phase_blocks\[pid\]\)",
               "py27_float_str(1.0 * (max_-min_)/len(phase_blocks[pid]))", src, 1)
    return _exec_patched(src, "ref_phasing", dict(
        py27_int_dict_order=py2emu.py27_int_dict_order, py27_float_str=py2emu.py27_float_str))


def load_rr_hctg_track() -> types.ModuleType:
    """falcon_unzip/rr_hctg_track.py with the Appendix C patch + B.4 order emulators."""
    with open(os.path.join(REF_ROOT, "falcon_unzip", "rr_hctg_track.py")) as f:
        src = f.read()
    src = _sub(r"ctg_score = ctg_score\.items\(\)\n(\s*)ctg_score\.sort\(",
               r"ctg_score = [(k, ctg_score[k]) for k in py27_str_dict_order(ctg_score)]\n"
               r"\1ctg_score.sort(", src, 1)
    src = _sub(r"for bread in bread_to_areads:",
               "for bread in py27_str_dict_order(bread_to_areads):", src, 1)
    src = _sub(r"for k in res:", "for k in py27_str_dict_order(res):", src, 1)
    # py2 set iteration order for rid_to_ctg[rid] (rr_hctg_track.py:21-22,120)
    return _exec_patched(src, "ref_rr_hctg_track", dict(
        py27_str_dict_order=py2emu.py27_str_dict_order, set=py2emu.Py27StrSet))


class TaskSelf:
    """The ``self`` a PypeTask function sees: file attributes + parameters dict."""

    def __init__(self, parameters=None, **files):
        self.parameters = parameters or {}
        self.__dict__.update(files)


def fake_samtools(directory: str) -> str:
    """A stand-in for ``samtools``: ``<it> view <sam text file> <ctg>`` cats the file."""
    path = os.path.join(directory, "fake_samtools.sh")
    with open(path, "w") as f:
        f.write('#!/bin/sh\ncat "$2"\n')
    os.chmod(path, os.stat(path).st_mode | stat.S_IXUSR)
    return path


def run_phasing_stages(sam_path: str, ctg_id: str, ref_seq: str, out_dir: str,
                       stages: Iterable[str] = ("het", "atable", "blocks", "reads"),
                       mod: types.ModuleType | None = None) -> Dict[str, str]:
    """Run the four stage functions of the (patched) reference with the file layout of
    reference phasing.py:501-503,520,534,543.  Returns {logical name: path}."""
    mod = mod or load_phasing()
    base = os.path.join(out_dir, ctg_id)
    paths = dict(variant_map=os.path.join(base, "het_call", "variant_map"),
                 variant_pos=os.path.join(base, "het_call", "variant_pos"),
                 q_id_map=os.path.join(base, "het_call", "q_id_map"),
                 atable=os.path.join(base, "g_atable", "atable"),
                 phased_variants=os.path.join(base, "get_phased_blocks", "phased_variants"),
                 phased_reads=os.path.join(base, "phased_reads"))
    for p in paths.values():
        os.makedirs(os.path.dirname(p), exist_ok=True)
    st = fake_samtools(out_dir)
    if "het" in stages:
        mod.make_het_call(TaskSelf(
            dict(ctg_id=ctg_id, ref_seq=ref_seq, base_dir=out_dir, samtools=st),
            bam_file=sam_path, vmap_file=paths["variant_map"], vpos_file=paths["variant_pos"],
            q_id_map_file=paths["q_id_map"]))
        import gc
        gc.collect()                               # the reference never closes vmap/vpos
    if "atable" in stages:
        mod.generate_association_table(TaskSelf(
            dict(ctg_id=ctg_id, base_dir=out_dir), vmap_file=paths["variant_map"],
            atable_file=paths["atable"]))
    if "blocks" in stages:
        mod.get_phased_blocks(TaskSelf(
            {}, vmap_file=paths["variant_map"], atable_file=paths["atable"],
            phased_variant_file=paths["phased_variants"]))
    if "reads" in stages:
        mod.get_phased_reads(TaskSelf(
            dict(ctg_id=ctg_id), vmap_file=paths["variant_map"], q_id_map_file=paths["q_id_map"],
            phased_variant_file=paths["phased_variants"], phased_read_file=paths["phased_reads"]))
    return paths


class FakePool:
    imap = staticmethod(map)

    def terminate(self):
        pass


def run_rr_track(las_lines: Dict[str, List[str]], phased_read_file: str, read_to_contig_map: str,
                 rawread_ids: str, out_path: str, min_len: int = 2500, bestn: int = 40,
                 mod: types.ModuleType | None = None) -> str:
    """run_track_reads of the (patched) reference over in-memory LA4Falcon lines.
    ``las_lines`` maps a LAS file name to its ``LA4Falcon -m`` text lines; the file list
    is sorted (glob order is filesystem dependent upstream)."""
    mod = mod or load_rr_hctg_track()

    def run_tr_stage1(db_fn, fn, min_len, bestn, rid_to_ctg, rid_to_phase):
        return fn, mod.tr_stage1(lambda: iter(las_lines[fn]), min_len, bestn, rid_to_ctg,
                                 rid_to_phase)
    mod.run_tr_stage1 = run_tr_stage1
    mod.run_track_reads(FakePool(), phased_read_file, read_to_contig_map, rawread_ids,
                        sorted(las_lines), min_len, bestn, "raw_reads.db", out_path)
    return out_path


def load_ovlp_filter(las_lines: Dict[str, List[str]]) -> types.ModuleType:
    """falcon_unzip/ovlp_filter_with_phase.py: the print statement of main() (:352) and xrange are
    patched, the LA4Falcon pipe (`sp.check_output`, :60,:150,:195) is replaced by in-memory text.
    ``las_lines`` maps a LAS file name to its ``LA4Falcon -mo`` lines."""
    with open(os.path.join(REF_ROOT, "falcon_unzip", "ovlp_filter_with_phase.py")) as f:
        src = f.read()
    src = _sub(r'print " "\.join\(l\)', 'print(" ".join(l))', src, 1)
    src = _sub(r"\bxrange\b", "range", src, 4)
    mod = _exec_patched(src, "ref_ovlp_filter", {})

    def check_output(cmd):
        return "".join(x if x.endswith("\n") else x + "\n" for x in las_lines[cmd[-1]])
    mod.sp = types.SimpleNamespace(check_output=check_output)
    return mod


def run_ovlp_filter(las_lines: Dict[str, List[str]], rid_phase_rows: Iterable[str], max_diff: int, max_cov: int,
                    min_cov: int, min_len: int, bestn: int) -> str:
    """main() of the (patched) reference (:309-352) without the process pool -> the text it prints."""
    mod = load_ovlp_filter(las_lines)
    mod.arid2phase.clear()
    for row in rid_phase_rows:
        row = row.strip().split()
        mod.arid2phase[row[0]] = (row[1], row[2], row[3])
    files = list(las_lines)
    ignore_all: List = []
    for fn in files:
        ignore_all.extend(mod.filter_stage1(("db", fn, max_diff, max_cov, min_cov, min_len))[1])
    ignore_all = set(ignore_all)
    contained = set()
    for fn in files:
        contained.update(mod.filter_stage2(("db", fn, max_diff, max_cov, min_cov, min_len, ignore_all))[1])
    out = []
    for fn in files:
        for l in mod.filter_stage3(("db", fn, max_diff, max_cov, min_cov, min_len, ignore_all, contained, bestn))[1]:
            out.append(" ".join(l) + "\n")
    return "".join(out)


def load_phasing_readmap() -> types.ModuleType:
    """falcon_unzip/phasing_readmap.py: print chevron (:50), py2 int division (:23) and the dict order
    of the output loop (:48) patched."""
    with open(os.path.join(REF_ROOT, "falcon_unzip", "phasing_readmap.py")) as f:
        src = f.read()
    src = _sub(r"print >>\s*(\w+),\s*(.*)", r"print(\2, file=\1)", src, 1)
    src = _sub(r"rid = int\(fid\.split\('/'\)\[1\]\)/10", "rid = int(fid.split('/')[1])//10", src, 1)
    src = _sub(r"for arid, phase in arid_to_phase\.items\(\):",
               "for arid, phase in [(k, arid_to_phase[k]) for k in py27_str_dict_order(arid_to_phase)]:", src, 1)
    return _exec_patched(src, "ref_phasing_readmap", dict(py27_str_dict_order=py2emu.py27_str_dict_order))


def load_get_read_hctg_map() -> types.ModuleType:
    """falcon_unzip/get_read_hctg_map.py: print chevron (:59) and the dict / set orders (:55-57) patched."""
    with open(os.path.join(REF_ROOT, "falcon_unzip", "get_read_hctg_map.py")) as f:
        src = f.read()
    src = _sub(r"print >>\s*(\w+),\s*(.*)", r"print(\2, file=\1)", src, 1)
    src = _sub(r"for k in pread_to_contigs:", "for k in py27_tuple_dict_order(pread_to_contigs):", src, 1)
    src = _sub(r"pread_to_contigs\.setdefault\( (k[12]), set\(\) \)", r"pread_to_contigs.setdefault( \1, Py27StrSet() )", src, 2)
    return _exec_patched(src, "ref_get_read_hctg_map", dict(py27_tuple_dict_order=py2emu.py27_tuple_dict_order,
                                                           Py27StrSet=py2emu.Py27StrSet))


def get_rid_to_phase_all_source() -> types.ModuleType:
    """The module-level task get_rid_to_phase_all of falcon_unzip/unzip.py:303-314, cut out of the file
    (the rest of unzip.py needs pypeflow / ConfigParser and is not on the path)."""
    with open(os.path.join(REF_ROOT, "falcon_unzip", "unzip.py")) as f:
        src = f.read()
    m = re.search(r"^def get_rid_to_phase_all\(self\):\n(?:(?:[ \t]+.*)?\n)+", src, flags=re.M)
    if not m:
        raise RuntimeError("get_rid_to_phase_all not found in unzip.py")
    return _exec_patched(m.group(0), "ref_get_rid_to_phase_all", dict(fn=lambda p: p))


# ---------------------------------------------------------------- select_reads_from_bam.py
class _FakeRead:
    def __init__(self, raw: bytes):
        self.raw = raw
        self.query_name = raw[36:36 + raw[12] - 1].decode("latin-1")


def _fake_pysam(outputs: dict) -> types.ModuleType:
    """Stand-in for the three things the reference uses of pysam (select_reads_from_bam.py:44-88): AlignmentFile(fn, 'rb',
    check_sq=False) with .header (a dict of the parsed header) and .fetch(until_eof=True); AlignmentFile(fn, 'wb',
    header=...) with .write(read); .close().  Written files are not BAMs: outputs[fn] = (header dict at open time,
    [record bytes...])."""
    import copy
    from falcon_unzip_b200 import bam
    from . import select_oracle
    mod = types.ModuleType("pysam")

    class AlignmentFile:
        def __init__(self, fn, mode, check_sq=True, header=None):
            self.mode = mode
            if mode == "rb":
                text, _refs, recs = bam.read_bam(fn)
                self.header = select_oracle.parse_header(text)
                self._buf = bytes(recs)
            else:
                outputs[fn] = (copy.deepcopy(header), [])
                self._out = outputs[fn][1]

        def fetch(self, until_eof=False):
            off = bam.index_records(self._buf)
            for i in range(len(off) - 1):
                yield _FakeRead(self._buf[off[i]:off[i + 1]])

        def write(self, r):
            self._out.append(r.raw)

        def close(self):
            pass
    mod.AlignmentFile = AlignmentFile
    return mod


def run_select_reads(input_bam_fofn_fn: str, rawread_to_contigs_fn: str, rawread_ids_fn: str, sam_dir: str) -> dict:
    """The reference's select_reads_from_bam() on real files, pysam replaced by the stand-in above.
    -> {output path: (header dict, [record bytes, ...])}."""
    import contextlib
    import io as _io
    with open(os.path.join(REF_ROOT, "falcon_unzip", "select_reads_from_bam.py")) as f:
        src = f.read()
    src = _sub(r'^(\s*)print "([^\n#]*?)\s*(#[^\n]*)?$', r'\1print("\2)', src, 5, re.M)
    src = _sub(r"print >>\s*([\w.]+),\s*(.*)", r"print(\2, file=\1)", src, 1)
    src = _sub(r"ctgs = read_partition\.keys\(\)\n(\s*)ctgs\.sort\(\)", r"ctgs = sorted(read_partition.keys())", src, 1)
    outputs: dict = {}
    saved = sys.modules.get("pysam")
    sys.modules["pysam"] = _fake_pysam(outputs)
    try:
        mod = _exec_patched(src, "ref_select_reads_from_bam", {})
        with contextlib.redirect_stdout(_io.StringIO()), contextlib.redirect_stderr(_io.StringIO()):
            mod.select_reads_from_bam(input_bam_fofn_fn, rawread_to_contigs_fn, rawread_ids_fn, sam_dir)
    finally:
        if saved is None:
            sys.modules.pop("pysam", None)
        else:
            sys.modules["pysam"] = saved
    return outputs
