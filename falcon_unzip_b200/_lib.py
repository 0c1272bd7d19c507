"""ctypes binding of libfuz.so (include/fuz.h).  There is no fallback: if the shared
library is missing the import fails loudly, and creating a context without a B200 fails
inside the library."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfuz.so")

FUZ_OK, FUZ_E_CUDA, FUZ_E_ARG, FUZ_E_CAPACITY, FUZ_E_BADRECORD, FUZ_E_UNSORTED, FUZ_E_DEPTH, \
    FUZ_E_INTERNAL, FUZ_E_FORMAT = range(9)

_u8p, _i32p, _i64p, _u32p = (C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                             C.POINTER(C.c_uint32))


class FuzError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__("libfuz error %d: %s" % (code, msg))
        self.code = code


class Batch(C.Structure):
    _fields_ = [("n_ctg", C.c_int32), ("n_rec", C.c_int32), ("rec_bytes", C.c_int64),
                ("d_rec_buf", C.c_void_p), ("d_rec_off", C.c_void_p), ("d_rec_qid", C.c_void_p),
                ("d_ctg_rec_off", C.c_void_p), ("d_ctg_len", C.c_void_p), ("d_ctg_goff", C.c_void_p),
                ("d_ctg_nq", C.c_void_p), ("total_glen", C.c_int64), ("total_nq", C.c_int64)]


OUTPUT_ARRAYS = [  # (field, numpy dtype, capacity key, elements per row)
    ("site_ctg", "i4", "sites", 1), ("site_pos", "i4", "sites", 1), ("site_cnt", "i4", "sites", 4),
    ("site_al", "u1", "sites", 2), ("site_top", "u1", "sites", 2),
    ("vm_site", "i4", "vmap", 1), ("vm_qid", "i4", "vmap", 1), ("vm_base", "u1", "vmap", 1),
    ("at_s1", "i4", "atable", 1), ("at_s2", "i4", "atable", 1), ("at_ct", "i4", "atable", 4),
    ("ph_state", "u1", "sites", 1), ("ph_lext", "i4", "sites", 1), ("ph_rext", "i4", "sites", 1),
    ("ph_lscore", "i4", "sites", 1), ("ph_rscore", "i4", "sites", 1), ("ph_block", "i4", "sites", 1),
    ("pr_ctg", "i4", "reads", 1), ("pr_qid", "i4", "reads", 1), ("pr_block", "i4", "reads", 1),
    ("pr_phase", "i4", "reads", 1), ("pr_n0", "i4", "reads", 1), ("pr_n1", "i4", "reads", 1),
]
# field order of the C structs (must match include/fuz.h)
_OUT_ORDER = ["site_ctg", "site_pos", "site_cnt", "site_al", "site_top", "vm_site", "vm_qid", "vm_base",
              "at_s1", "at_s2", "at_ct", "ph_state", "ph_lext", "ph_rext", "ph_lscore", "ph_rscore",
              "ph_block", "pr_ctg", "pr_qid", "pr_block", "pr_phase", "pr_n0", "pr_n1"]


class Outputs(C.Structure):
    _fields_ = ([("cap_sites", C.c_int64), ("cap_vmap", C.c_int64), ("cap_atable", C.c_int64),
                 ("cap_reads", C.c_int64)] + [("d_" + n, C.c_void_p) for n in _OUT_ORDER]
                + [("d_counts", C.c_void_p), ("d_ctg_nq", C.c_void_p), ("d_name_first", C.c_void_p)])


class HostBatch(C.Structure):
    _fields_ = [("n_ctg", C.c_int32), ("n_rec", C.c_int32), ("rec_bytes", C.c_int64),
                ("h_rec_buf", C.c_void_p), ("h_rec_off", C.c_void_p), ("h_rec_qid", C.c_void_p),
                ("h_ctg_rec_off", C.c_void_p), ("h_ctg_len", C.c_void_p), ("h_ctg_nq", C.c_void_p)]


class HostOutputs(C.Structure):
    _fields_ = ([("cap_sites", C.c_int64), ("cap_vmap", C.c_int64), ("cap_atable", C.c_int64),
                 ("cap_reads", C.c_int64)] + [(n, C.c_void_p) for n in _OUT_ORDER] +
                [("ctg_nq", C.c_void_p), ("name_first", C.c_void_p)])


class RRInput(C.Structure):
    _fields_ = [("n_ovl", C.c_int64), ("d_q", C.c_void_p), ("d_t", C.c_void_p), ("d_len", C.c_void_p),
                ("d_tlen", C.c_void_p), ("d_file", C.c_void_p), ("n_reads", C.c_int32), ("d_in_map", C.c_void_p),
                ("d_ph_ctg", C.c_void_p), ("d_ph_block", C.c_void_p), ("d_ph_phase", C.c_void_p),
                ("d_rc_off", C.c_void_p), ("d_rc_ctg", C.c_void_p), ("min_len", C.c_int32), ("bestn", C.c_int32),
                ("n_ctg", C.c_int32)]


class RROutputs(C.Structure):
    _fields_ = [("d_keep", C.c_void_p), ("d_hp_n", C.c_void_p), ("d_hp_len", C.c_void_p), ("d_hp_q", C.c_void_p),
                ("cap_votes", C.c_int64), ("d_vt_off", C.c_void_p), ("d_vt_ctg", C.c_void_p),
                ("d_vt_count", C.c_void_p), ("d_vt_score", C.c_void_p)]


class OvlpInput(C.Structure):
    _fields_ = ([("n_ovl", C.c_int64)] + [("d_" + k, C.c_void_p) for k in ("q", "t", "len", "qs", "qe", "ql", "ts", "te", "tl",
                                                                          "flags", "file")] +
                [("n_reads", C.c_int32), ("d_in_map", C.c_void_p), ("d_ph_ctg", C.c_void_p), ("d_ph_block", C.c_void_p),
                 ("d_ph_phase", C.c_void_p), ("max_diff", C.c_int32), ("max_ovlp", C.c_int32), ("min_ovlp", C.c_int32),
                 ("min_len", C.c_int32), ("bestn", C.c_int32), ("stage", C.c_int32), ("d_ignore_in", C.c_void_p),
                 ("d_contained_in", C.c_void_p)])


class OvlpOutputs(C.Structure):
    _fields_ = [("d_ignore", C.c_void_p), ("d_contained", C.c_void_p), ("cap_groups", C.c_int64), ("d_grp_q", C.c_void_p),
                ("d_grp_line", C.c_void_p), ("d_grp_ignore", C.c_void_p), ("d_grp_tie", C.c_void_p), ("d_grp_off", C.c_void_p), ("cap_out", C.c_int64),
                ("d_out_line", C.c_void_p), ("d_cand", C.c_void_p)]


class Status(C.Structure):
    _fields_ = [("error", C.c_int32), ("error_index", C.c_int32), ("n_sites", C.c_int64),
                ("n_vmap", C.c_int64), ("n_atable", C.c_int64), ("n_reads", C.c_int64),
                ("need_sites", C.c_int64), ("need_vmap", C.c_int64), ("need_atable", C.c_int64),
                ("need_reads", C.c_int64), ("need_pairs", C.c_int64), ("n_accepted", C.c_int64),
                ("aligned_bases", C.c_int64), ("n_segments", C.c_int64), ("reserved", C.c_int64 * 4)]


# every symbol include/fuz.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("fuz_version", C.c_int, []),
    ("fuz_tile_size", C.c_int, []),
    ("fuz_ctx_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("fuz_ctx_destroy", C.c_int, [C.c_void_p]),
    ("fuz_last_error", C.c_char_p, [C.c_void_p]),
    ("fuz_set_stream", C.c_int, [C.c_void_p, C.c_void_p]),
    ("fuz_set_option", C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    ("fuz_sync", C.c_int, [C.c_void_p]),
    ("fuz_get_status", C.c_int, [C.c_void_p, C.POINTER(Status)]),
    ("fuz_launch_count", C.c_int64, [C.c_void_p]),
    ("fuz_kernel_timing", C.c_int, [C.c_void_p, C.c_int]),
    ("fuz_get_kernel_timing", C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    ("fuz_profile", C.c_int, [C.c_void_p, C.c_int]),
    ("fuz_profile_report", C.c_int64, [C.c_void_p, C.c_char_p, C.c_int64]),
    ("fuz_het_call", C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Outputs)]),
    ("fuz_association_table", C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(Outputs)]),
    ("fuz_phased_blocks", C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.POINTER(Outputs)]),
    ("fuz_phased_reads", C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                   C.POINTER(Outputs)]),
    ("fuz_phase_batch", C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(Outputs)]),
    ("fuz_phase_batch_host", C.c_int, [C.c_void_p, C.POINTER(HostBatch), C.POINTER(HostOutputs),
                                       C.POINTER(Status), _i64p, _i64p]),
    ("fuz_assign_qids", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    ("fuz_host_bgzf_index", C.c_int64, [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fuz_bgzf_inflate", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_void_p, C.c_int64]),
    ("fuz_bam_index_records", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                        _i64p, _i64p]),
    ("fuz_bam_index_window", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p,
                                        _i64p, _i64p, _i64p]),
    ("fuz_bam_index_files", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                      C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, _i64p, _i64p, _i64p]),
    ("fuz_gather_records", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_int64]),
    ("fuz_rr_track", C.c_int, [C.c_void_p, C.POINTER(RRInput), C.POINTER(RROutputs)]),
    ("fuz_ovlp_filter", C.c_int, [C.c_void_p, C.POINTER(OvlpInput), C.POINTER(OvlpOutputs)]),
    ("fuz_parse_la4falcon", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32] + [C.c_void_p] * 12),
    ("fuz_host_parse_la4falcon_mo", C.c_int64, [C.c_char_p, C.c_int64, C.c_int64] + [C.c_void_p] * 12),
    ("fuz_host_format_ovlp", C.c_int64, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("fuz_host_parse_la4falcon", C.c_int64, [C.c_char_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_void_p]),
    ("fuz_host_index_records", C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, _i64p]),
    ("fuz_host_assign_qids", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32,
                                       C.c_void_p, C.c_void_p, C.c_void_p]),
    ("fuz_host_py27_int_dict_order", C.c_int, [C.c_void_p, C.c_int64, C.c_void_p]),
    ("fuz_host_format_variant_map", C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_char_p,
                                                C.c_int64, C.c_char_p, C.c_int64]),
    ("fuz_host_format_atable", C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                           C.c_char_p, C.c_int64]),
    ("fuz_host_format_phased_reads", C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_int64, C.c_char_p, C.c_char_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int64]),
    ("fuz_host_format_phased_reads_rows", C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                      C.c_int64, C.c_char_p, C.c_void_p, C.c_int64, C.c_int64, C.c_char_p, C.c_int64]),
    ("fuz_host_format_q_id_map_rows", C.c_int64, [C.c_void_p, C.c_int64, C.c_int64, C.c_char_p, C.c_int64]),
    ("fuz_host_format_variant_pos", C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]),
    ("fuz_host_format_phased_variants", C.c_int64, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_int64, C.c_int64, C.c_char_p, C.c_int64, C.c_char_p, C.c_int64]),
    ("fuz_host_py27_str_dict_order", C.c_int64, [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("fuz_host_rr_bread_order", C.c_int64, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("fuz_host_rr_format_rows", C.c_int64, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
]

_lib = None


def lib():
    """The loaded library; raises if libfuz.so has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `make -C falcon_unzip_b200/csrc` (or "
                "`python -c 'import __graft_entry__ as g; g.build()'`).  falcon_unzip_b200 has no "
                "CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            f = getattr(l, name)
            f.restype = res
            f.argtypes = args
        _lib = l
    return _lib


def check(ctx, rc: int) -> None:
    if rc != FUZ_OK:
        msg = lib().fuz_last_error(ctx)
        raise FuzError(rc, msg.decode("utf-8", "replace") if msg else "?")
