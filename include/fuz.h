/*
 * fuz.h -- C ABI of libfuz.so, the B200 (sm_100a) implementation of FALCON-Unzip's
 * read-phasing hot path.  Plain pointers and sizes only; every function returns an int
 * status (0 = FUZ_OK) and records a message retrievable with fuz_last_error().
 *
 * Reference interfaces replaced (PacificBiosciences/FALCON_unzip):
 *   fuz_het_call            falcon_unzip/phasing.py:14-134   make_het_call
 *   fuz_association_table   falcon_unzip/phasing.py:137-206  generate_association_table
 *   fuz_phased_blocks       falcon_unzip/phasing.py:208-421  get_score + get_phased_blocks
 *   fuz_phased_reads        falcon_unzip/phasing.py:423-480  get_phased_reads
 *   fuz_phase_batch         falcon_unzip/phasing.py:482-553  phasing() (the four stages
 *                           chained on the device for a batch of contigs)
 *   fuz_rr_track            falcon_unzip/rr_hctg_track.py:31-65,97-105,113-123
 *                           tr_stage1 + heap merge + contig vote of run_track_reads
 *   fuz_bgzf_inflate, fuz_bam_index_records, fuz_bam_index_window, fuz_bam_index_files
 *                           falcon_unzip/phasing.py:27        the `samtools view` pipe: BGZF
 *                           inflate + record split, on the device
 *   fuz_ovlp_filter         falcon_unzip/ovlp_filter_with_phase.py:49-290  filter_stage1-3
 *   fuz_gather_records      falcon_unzip/select_reads_from_bam.py:70-86    per-contig partition of the
 *                           raw-read BAM records (with the BAM ingest entries above)
 *   fuz_host_*              host-side helpers of the same path (record index, QNAME ->
 *                           q_id of phasing.py:47-54)
 *
 * Conventions
 *   - "d_" pointers are DEVICE pointers owned by the caller (PyTorch tensors in the
 *     Python host layer); "h_" pointers are host pointers.  The library never frees
 *     caller memory.  Scratch memory is owned by the context and grows on demand.
 *   - Outputs go into caller-provided capacity; if a capacity is too small the call
 *     fails with FUZ_E_CAPACITY and fuz_status.need_* says how much is needed.
 *   - A context is bound to one device and one stream and is not thread safe
 *     (the reference runs one OS process per contig, unzip.py:255).
 *   - No CPU fallback exists: without a CUDA device fuz_ctx_create fails.
 *
 * Device data model (struct-of-arrays, all little endian)
 *   records   verbatim uncompressed BAM alignment records (block_size + body, SAM spec
 *             4.2) concatenated in `rec_buf`; rec_off[i] is the byte offset of record i
 *             (n_rec + 1 entries).  rec_buf must be followed by >= 64 readable bytes.
 *             Records are grouped by contig (ctg_rec_off) and coordinate sorted inside a
 *             contig, i.e. the order `samtools view <bam> <ctg>` prints.
 *   sites     het-SNP sites ordered by (contig, position): site_ctg, site_pos (1-based
 *             position inside the contig, the value the reference writes to files),
 *             site_cnt[4] depth of A,C,G,T, site_al[2] the two called alleles as base
 *             indices (0..3 = A,C,G,T) ordered by the string "ACTG" (CPython-2 dict order,
 *             SURVEY.md B.1), site_top[2] the same two alleles ordered major, minor.
 *   vmap      variant_map rows in file order: vm_site (index into sites), vm_base
 *             (0..3), vm_qid (q_id inside the contig).
 *   atable    association rows ordered by (site1, site2): at_s1, at_s2, at_ct[4] =
 *             c11 c12 c21 c22 with alleles in site_al order.
 *   phase     per site: ph_state (0 = (al0,al1), 1 = (al1,al0), 255 = site not in any
 *             accepted atable row), ph_lext/ph_rext (1-based positions), ph_lscore,
 *             ph_rscore, ph_block (block id, dense from 1 per contig, 0 = none).
 *   reads     phased_reads rows ordered by (contig, q_id, block): pr_ctg, pr_qid,
 *             pr_block, pr_phase, pr_n0, pr_n1.
 */
#ifndef FUZ_H_
#define FUZ_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FUZ_VERSION 1

enum {
    FUZ_OK = 0,
    FUZ_E_CUDA = 1,        /* CUDA runtime error (no device, launch failure, ...)      */
    FUZ_E_ARG = 2,         /* invalid argument                                          */
    FUZ_E_CAPACITY = 3,    /* an output / scratch capacity was too small (see status)   */
    FUZ_E_BADRECORD = 4,   /* a record the reference would crash on (CIGAR '*', SEQ '*',
                              M/=/X running past SEQ, unknown CIGAR op; phasing.py:72,84) */
    FUZ_E_UNSORTED = 5,    /* records not coordinate sorted / not grouped by contig      */
    FUZ_E_DEPTH = 6,       /* more than 65535 reads over one pileup tile                 */
    FUZ_E_INTERNAL = 7,    /* kernels disagree with each other (a bug)                   */
    FUZ_E_FORMAT = 8       /* stage input not in the reference-produced format           */
};

/* thresholds are the reference's literals (phasing.py:72-75,100,112,169,192,205,245,
 * 394,398,477-479); they are compiled in, not configurable. */

typedef struct fuz_ctx fuz_ctx;

/* One batch of contigs.  Global pileup coordinates: contig c occupies
 * [ctg_goff[c], ctg_goff[c] + ctg_len[c]); ctg_goff must be multiples of
 * FUZ_TILE (fuz_tile_size()) so that no pileup tile straddles two contigs. */
typedef struct {
    int32_t n_ctg;
    int32_t n_rec;
    int64_t rec_bytes;
    const uint8_t *d_rec_buf;     /* [rec_bytes + 64]                                   */
    const int64_t *d_rec_off;     /* [n_rec + 1]                                        */
    const int32_t *d_rec_qid;     /* [n_rec] q_id of the record's QNAME in its contig, or NULL:
                                   * fuz_phase_batch assigns them (d_ctg_nq / total_nq ignored) */
    const int32_t *d_ctg_rec_off; /* [n_ctg + 1] record range of each contig            */
    const int32_t *d_ctg_len;     /* [n_ctg]                                            */
    const int64_t *d_ctg_goff;    /* [n_ctg + 1] tile-aligned global offsets            */
    const int32_t *d_ctg_nq;      /* [n_ctg] number of distinct QNAMEs (q_ids) per contig */
    int64_t total_glen;           /* ctg_goff[n_ctg]                                    */
    int64_t total_nq;             /* sum of ctg_nq                                      */
} fuz_batch;

typedef struct {
    int64_t cap_sites, cap_vmap, cap_atable, cap_reads;
    /* sites */
    int32_t *d_site_ctg, *d_site_pos, *d_site_cnt /* [cap_sites*4] */;
    uint8_t *d_site_al /* [cap_sites*2] */, *d_site_top /* [cap_sites*2] */;
    /* vmap */
    int32_t *d_vm_site, *d_vm_qid; uint8_t *d_vm_base;
    /* atable */
    int32_t *d_at_s1, *d_at_s2, *d_at_ct /* [cap_atable*4] */;
    /* phase (per site, capacity cap_sites) */
    uint8_t *d_ph_state; int32_t *d_ph_lext, *d_ph_rext, *d_ph_lscore, *d_ph_rscore, *d_ph_block;
    /* reads */
    int32_t *d_pr_ctg, *d_pr_qid, *d_pr_block, *d_pr_phase, *d_pr_n0, *d_pr_n1;
    /* optional full pileup (debug / parity): [total_glen*4] depth of A,C,G,T, or NULL   */
    uint32_t *d_counts;
    /* optional, filled when fuz_phase_batch assigns the q_ids (fuz_batch.d_rec_qid == NULL):
     * distinct QNAMEs per contig [n_ctg]; first record of every q_id [n_rec]            */
    int32_t *d_ctg_nq;
    int64_t *d_name_first;
} fuz_outputs;

/* Filled on the device, copied to the host by fuz_get_status(). */
typedef struct {
    int32_t error;                /* FUZ_OK or the first FUZ_E_* raised by a kernel      */
    int32_t error_index;          /* record / site / tile index the error refers to      */
    int64_t n_sites, n_vmap, n_atable, n_reads;
    int64_t need_sites, need_vmap, need_atable, need_reads, need_pairs;
    int64_t n_accepted;           /* records passing the filter of phasing.py:72-75      */
    int64_t aligned_bases;        /* sum of M/=/X lengths over accepted records          */
    int64_t n_segments;
    int64_t reserved[4];
} fuz_status;

/* ---- context ------------------------------------------------------------------ */
int fuz_version(void);
int fuz_tile_size(void);
int fuz_ctx_create(int device, fuz_ctx **out);
int fuz_ctx_destroy(fuz_ctx *ctx);
const char *fuz_last_error(fuz_ctx *ctx);      /* ctx may be NULL: create-time error     */
int fuz_set_stream(fuz_ctx *ctx, void *cuda_stream);
/* options: "pileup_impl" 0 = reference-aligned 4-bit projection of every read + tiled register pileup fused with the het
 *                          test (default),
 *          1 = global-atomic pileup + separate het test (cross-check path),
 *          2 = segment-list pileup: CIGARs become match segments once, SEQ slices are staged into shared memory by
 *              bulk async copies (cp.async.bulk + mbarrier producer / consumer pipeline) and cut straight into
 *              bit-sliced counters without a projection (measured slower than 0; kept as an
 *              independent implementation that the tests compare bit for bit);
 *          3 = the projection of 0 produced from segment lists (k_segments + k_project_seg: one lane per segment, no owner
 *              search; as fast as 0, kept as a cross-check);
 *          "seg_cap" / "ent_cap" minimum reservation of segment slots / tile entries of pileup_impl 2 (a batch denser
 *                          than the built-in heuristics fails with FUZ_E_CAPACITY, error_index 6 / 9, and
 *                          fuz_status.n_segments / reserved[0] say how much it needs);
 *          "host_fetch" 1 = fuz_phase_batch_host reads page-locked records through the host mapping and
 *                       moves only header/name/CIGAR/SEQ (default), 0 = always copy the whole buffer,
 *          "pdl" 1 = kernels of a call are launched with programmatic stream serialisation, i.e. the
 *                launch of a kernel overlaps the tail of its predecessor (default), 0 = plain launches,
 *          "rr_filter_only" 1 = fuz_rr_track runs the overlap filter only (d_keep; the map step of the
 *                           multi-GPU run, the merge then runs on the gathered kept lines), 0 = default,
 *          "phase_staging" 0 = stage as much of a contig as fits in shared memory (default),
 *                          1 = at most the sweep tier, 2 = global memory only (both for tests),
 *          "fetch_ctas" CTAs of the selective record fetch of fuz_phase_batch_host (default 1184; two contexts that
 *                       alternate over the batches of a list use fewer, so that one computes while the other fetches),
 *          "gather_tma" 1 = the register pileup of pileup_impl 0 / 3 is fed by bulk async copies of the projection rows into a
 *                       shared-memory ring (persistent producer / consumer kernel, default), 0 = one CTA per tile with plain loads,
 *          "project_ctas" grid of the projection kernel (default 148 * 5, the CTAs per SM its registers allow),
 *          "grid_sig" / "grid_assoc" / "grid_reads" CTAs per SM (1 .. 8) of the grid-stride kernels of the signature, the
 *                       association and the read stage (defaults 8 / 6 / 8: they are latency bound),
 *          "grid_rr" CTAs per SM of k_rr_replay (default 4, the measured optimum; k_rr_fill / k_rr_vote then run on 8 / 16), 0 = all on 4,
 *          "sweep_passes" passes of the parallel fixed-point form of the pass-2 sweep (phasing.py:311-344) before
 *                         the sequential walk takes over (default 64; 0 = sequential only; same result either way),
 *          "max_pairs_per_site" capacity factor of the association scratch (default 96) */
int fuz_set_option(fuz_ctx *ctx, const char *key, int64_t value);
int fuz_sync(fuz_ctx *ctx);
/* device -> host copy of the status block of the last call (synchronises the stream) */
int fuz_get_status(fuz_ctx *ctx, fuz_status *h_status);
/* number of kernel launches issued by this context since creation */
int64_t fuz_launch_count(fuz_ctx *ctx);
/* CUDA-event timing of the dominant (pileup) kernel: enable, then read the accumulated
 * milliseconds and launch count (synchronises).  Used by bench.py's roofline block.    */
int fuz_kernel_timing(fuz_ctx *ctx, int enable);
int fuz_get_kernel_timing(fuz_ctx *ctx, double *h_ms_total, int64_t *h_launches);
/* Diagnostics: per-launch device time.  fuz_profile(ctx, 1) starts a capture (an event is
 * recorded after every kernel launch); fuz_profile_report writes "kernel\tms" lines. */
int fuz_profile(fuz_ctx *ctx, int enable);
int64_t fuz_profile_report(fuz_ctx *ctx, char *buf, int64_t cap);

/* ---- stages (asynchronous on the context stream; check with fuz_get_status) ---- */
/* phasing.py:14-134: record scan + filter, pileup, het call, variant_map rows.
 * Fills sites + vmap of `out`. */
int fuz_het_call(fuz_ctx *ctx, const fuz_batch *in, fuz_outputs *out);
/* phasing.py:137-206.  Inputs: sites (ctg,pos,al) + vmap of `out` with the given counts;
 * fills atable. */
int fuz_association_table(fuz_ctx *ctx, int32_t n_ctg, int64_t n_sites, int64_t n_vmap,
                          fuz_outputs *out);
/* phasing.py:216-421.  Inputs: sites + atable (n_atable rows, ordered by (s1,s2));
 * fills phase.  */
int fuz_phased_blocks(fuz_ctx *ctx, int32_t n_ctg, int64_t n_sites, int64_t n_atable,
                      fuz_outputs *out);
/* phasing.py:423-480.  Inputs: sites, vmap, phase; d_ctg_nq [n_ctg] with sum total_nq;
 * fills reads. */
int fuz_phased_reads(fuz_ctx *ctx, int32_t n_ctg, const int32_t *d_ctg_nq, int64_t total_nq,
                     int64_t n_sites, int64_t n_vmap, fuz_outputs *out);
/* phasing.py:482-553: all four stages for the batch, no host synchronisation between
 * them (counts stay on the device).  This is the timed hot path of bench.py. */
int fuz_phase_batch(fuz_ctx *ctx, const fuz_batch *in, fuz_outputs *out);

/* ---- host-buffer entry (the reference-facing call; includes H2D / D2H) ---------- */
typedef struct {
    int32_t n_ctg, n_rec;
    int64_t rec_bytes;
    const uint8_t *h_rec_buf;     /* [rec_bytes] verbatim BAM records                    */
    const int64_t *h_rec_off;     /* [n_rec + 1]                                         */
    const int32_t *h_rec_qid;     /* [n_rec], or NULL: q_ids are assigned on the device  */
    const int32_t *h_ctg_rec_off; /* [n_ctg + 1]                                         */
    const int32_t *h_ctg_len;     /* [n_ctg]                                             */
    const int32_t *h_ctg_nq;      /* [n_ctg] (ignored when h_rec_qid is NULL)            */
} fuz_host_batch;

typedef struct {
    int64_t cap_sites, cap_vmap, cap_atable, cap_reads;
    int32_t *site_ctg, *site_pos, *site_cnt; uint8_t *site_al, *site_top;
    int32_t *vm_site, *vm_qid; uint8_t *vm_base;
    int32_t *at_s1, *at_s2, *at_ct;
    uint8_t *ph_state; int32_t *ph_lext, *ph_rext, *ph_lscore, *ph_rscore, *ph_block;
    int32_t *pr_ctg, *pr_qid, *pr_block, *pr_phase, *pr_n0, *pr_n1;
    /* filled when the q_ids are assigned on the device (h_rec_qid == NULL); may be NULL   */
    int32_t *ctg_nq;              /* [n_ctg] distinct QNAMEs per contig                   */
    int64_t *name_first;          /* [n_rec] first record of every q_id, contig after contig */
} fuz_host_outputs;

/* H2D of the batch, fuz_phase_batch, D2H of every row array (only the filled prefix).
 * Device staging buffers are owned by the context and reused across calls.
 * h_status receives the final status; bytes moved are reported for bench.py. */
int fuz_phase_batch_host(fuz_ctx *ctx, const fuz_host_batch *in, fuz_host_outputs *out,
                         fuz_status *h_status, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- q_id assignment on the device (falcon_unzip/phasing.py:47-54) ------------------ */
/* First-seen QNAME -> q_id per contig, before any filtering: the device form of
 * fuz_host_assign_qids.  d_rec_qid [n_rec], d_ctg_nq [n_ctg], d_name_first [n_rec] (record
 * index of the first record of every q_id, contig after contig; may be NULL). */
int fuz_assign_qids(fuz_ctx *ctx, const uint8_t *d_rec_buf, const int64_t *d_rec_off, int32_t n_rec,
                    int64_t rec_bytes, const int32_t *d_ctg_rec_off, int32_t n_ctg, int32_t *d_rec_qid,
                    int32_t *d_ctg_nq, int64_t *d_name_first);

/* ---- BAM ingest on the device (replaces the `samtools view <bam> <ctg>` pipe of
 *      falcon_unzip/phasing.py:27) ------------------------------- */
/* Host: walk the BGZF block chain of a BAM file image (SAM spec 4.1).  Per block: offset and
 * size of the raw deflate stream, offset of its payload in the inflated stream (h_uoff has
 * cap + 1 slots, h_uoff[n] = inflated size), CRC32 of the payload (h_crc may be NULL).
 * Returns the number of blocks (call with h_coff = NULL to count), -1 on a malformed file. */
int64_t fuz_host_bgzf_index(const uint8_t *h_file, int64_t n_bytes, int64_t cap, int64_t *h_coff,
                            int32_t *h_csize, int64_t *h_uoff, uint32_t *h_crc);
/* Device: inflate n_blk raw deflate streams (RFC 1951), block i = d_comp[d_coff[i], +d_csize[i])
 * -> d_out[d_uoff[i], d_uoff[i+1]), one warp per block; when d_crc is not NULL the CRC32 of every
 * payload is checked.  d_comp must be 4-byte aligned.  Asynchronous; a corrupt block raises
 * FUZ_E_FORMAT (error_index = block, status.reserved[3] = reason) in the status block. */
int fuz_bgzf_inflate(fuz_ctx *ctx, const uint8_t *d_comp, int64_t comp_bytes, const int64_t *d_coff,
                     const int32_t *d_csize, const int64_t *d_uoff, const uint32_t *d_crc, int64_t n_blk,
                     uint8_t *d_out, int64_t out_bytes);
/* Device form of fuz_host_index_records plus the grouping by reference id: d_rec = the alignment
 * records of a coordinate-sorted BAM (everything after the header), followed by >= 64 readable
 * bytes.  Fills d_rec_off [cap_rec + 1] and d_ctg_rec_off [n_ref + 1] (records of reference c =
 * [d_ctg_rec_off[c], d_ctg_rec_off[c+1]); unmapped records, refID -1, lie behind d_ctg_rec_off[n_ref]).
 * Synchronises; *h_n_rec = record count.  FUZ_E_CAPACITY: *h_need_rec says how many slots are needed;
 * FUZ_E_BADRECORD: broken block_size chain; FUZ_E_UNSORTED: refIDs not ascending. */
int fuz_bam_index_records(fuz_ctx *ctx, const uint8_t *d_rec, int64_t rec_bytes, int32_t n_ref, int64_t cap_rec,
                          int64_t *d_rec_off, int32_t *d_ctg_rec_off, int64_t *h_n_rec, int64_t *h_need_rec);
/* The same for one WINDOW of a record stream that is decoded piece by piece (raw-read BAMs of hundreds of GB, which
 * select_reads_from_bam.py:55-76 streams record by record): d_rec starts at a record boundary; a record cut by the end of
 * the buffer ends the index instead of breaking the chain, and *h_tail receives its offset (= rec_bytes when the window
 * ends on a record boundary).  The caller carries d_rec[*h_tail, rec_bytes) over to the front of the next window. */
int fuz_bam_index_window(fuz_ctx *ctx, const uint8_t *d_rec, int64_t rec_bytes, int32_t n_ref, int64_t cap_rec,
                         int64_t *d_rec_off, int32_t *d_ctg_rec_off, int64_t *h_n_rec, int64_t *h_need_rec, int64_t *h_tail);
/* The same for SEVERAL BAM files inflated into one device buffer (one fuz_bgzf_inflate over the blocks of all
 * files) -- the reference leaves one sorted BAM per contig (falcon_unzip/unzip.py:90-91) and runs fc_phasing.py
 * once per file (unzip.py:124).  File s has its alignment records at d_raw[h_seg_start[s], h_seg_end[s])
 * (ascending, not overlapping; the headers lie between them) and h_seg_nref[s] references; d_raw is followed by
 * >= 64 readable bytes.  All files are indexed in one pass; the MAPPED records of all files are written back to
 * back into d_rec_out [cap_bytes + 64] (unmapped tails dropped), d_rec_off [cap_rec + 1] are their offsets in
 * d_rec_out, d_ctg_rec_off [sum(n_ref) + 1] the record range of every reference, file after file.  cap_rec
 * counts ALL records of the files.  Synchronises; *h_n_rec = mapped records, *h_rec_bytes = their bytes.
 * Errors as fuz_bam_index_records (FUZ_E_BADRECORD: error_index = file or record). */
int fuz_bam_index_files(fuz_ctx *ctx, const uint8_t *d_raw, int64_t raw_bytes, int32_t n_seg, const int64_t *h_seg_start,
                        const int64_t *h_seg_end, const int32_t *h_seg_nref, int64_t cap_rec, uint8_t *d_rec_out,
                        int64_t cap_bytes, int64_t *d_rec_off, int32_t *d_ctg_rec_off, int64_t *h_n_rec,
                        int64_t *h_need_rec, int64_t *h_rec_bytes);

/* Device: whole records picked out of a record buffer, in the given order: output record i = source record
 * d_sel[i] (source record r = d_src[d_src_off[r], d_src_off[r+1])), written at d_dst[d_dst_off[i], d_dst_off[i+1])
 * (the caller scans the sizes; a size mismatch or an index out of range raises FUZ_E_ARG, error_index = i).
 * One warp per record.  This is the per-contig partition of the raw-read BAMs in
 * falcon_unzip/select_reads_from_bam.py:70-86 (d_sel = the records of one contig after another, file order
 * kept inside a contig).  d_src is followed by >= 8 readable bytes.  Asynchronous. */
int fuz_gather_records(fuz_ctx *ctx, const uint8_t *d_src, const int64_t *d_src_off, int64_t n_src, int64_t src_bytes,
                       const int64_t *d_sel, const int64_t *d_dst_off, int64_t m, uint8_t *d_dst, int64_t dst_bytes);

/* ---- raw-read -> haplotig tracking (falcon_unzip/rr_hctg_track.py) --------------- */
/* fuz_rr_track replaces tr_stage1 (:31-65), the heap merge (:97-105) and the contig vote
 * (:113-123) of run_track_reads for ALL LAS files at once.  The host parses the LA4Falcon
 * text (fuz_host_parse_la4falcon), builds the tables (:15-23, :72-85) and formats the rows
 * (:126-138, with the CPython-2 orders of SURVEY.md B.4). */
typedef struct {
    int64_t n_ovl;                /* overlap lines, concatenated in (sorted file, line) order */
    const int32_t *d_q, *d_t;     /* read ids (LA4Falcon columns 0, 1)                    */
    const int32_t *d_len;         /* overlap length = -int(col 2)                         */
    const int32_t *d_tlen;        /* col 11                                               */
    const int32_t *d_file;        /* LAS file index of the line (non-decreasing)          */
    int32_t n_reads;              /* size of the read-id space (len of rawread_ids list)  */
    const uint8_t *d_in_map;      /* [n_reads] 1 if rid is a key of rid_to_ctg            */
    const int32_t *d_ph_ctg, *d_ph_block, *d_ph_phase; /* [n_reads] phase table; ctg -1 = None */
    const int32_t *d_rc_off;      /* [n_reads + 1] CSR rid -> contigs                     */
    const int32_t *d_rc_ctg;      /* contig indices, each list in CPython-2 set iteration order */
    int32_t min_len, bestn, n_ctg;
} fuz_rr_input;

typedef struct {
    uint8_t *d_keep;              /* [n_ovl] 1 = the line passed the filter (:45-57)       */
    int32_t *d_hp_n;              /* [n_reads] entries of the merged heap of every target  */
    int32_t *d_hp_len, *d_hp_q;   /* [n_reads * bestn] merged heaps in heapq ARRAY order   */
    int64_t cap_votes;
    int32_t *d_vt_off;            /* [n_reads + 1] vote rows of every target               */
    int32_t *d_vt_ctg, *d_vt_count; int64_t *d_vt_score;  /* [cap_votes] in dict insertion order */
} fuz_rr_outputs;

/* status after the call: reserved[1] = vote rows, reserved[2] = vote rows needed,
 * reserved[3] = kept overlap lines. */
int fuz_rr_track(fuz_ctx *ctx, const fuz_rr_input *in, fuz_rr_outputs *out);

/* ---- overlap filter with phase (falcon_unzip/ovlp_filter_with_phase.py) ---- */
/* fuz_ovlp_filter replaces filter_stage1 (:49-143), filter_stage2 (:145-186) and filter_stage3
 * (:188-290) for ALL LAS files at once.  The host parses the `LA4Falcon -mo` text
 * (fuz_host_parse_la4falcon_mo), interns the strings of the rid -> (ctg, block, phase) table
 * (:319-322) and prints the selected lines (fuz_host_format_ovlp, :352). */
typedef struct {
    int64_t n_ovl;                /* overlap lines, concatenated in (file, line) order            */
    const int32_t *d_q, *d_t;     /* read ids (columns 0, 1; %09d ids)                            */
    const int32_t *d_len;         /* overlap length = -int(col 2)                                 */
    const int32_t *d_qs, *d_qe, *d_ql, *d_ts, *d_te, *d_tl;   /* columns 5, 6, 7, 9, 10, 11       */
    const uint8_t *d_flags;       /* bit 0: float(col 3) >= 90; bits 1-2: last column, 1 "overlap",
                                   * 2 "contains", 3 "contained", 0 anything else                  */
    const int32_t *d_file;        /* LAS file index of the line (non-decreasing)                  */
    int32_t n_reads;              /* size of the read-id space                                    */
    const uint8_t *d_in_map;      /* [n_reads] 1 if the id is a key of arid2phase                 */
    const int32_t *d_ph_ctg, *d_ph_block, *d_ph_phase;   /* [n_reads] interned strings of the row  */
    int32_t max_diff, max_ovlp, min_ovlp, min_len, bestn;
    int32_t stage;                /* 1, 2, 3: stop after that stage (3 = the whole filter)        */
    const uint8_t *d_ignore_in;   /* [n_reads] or NULL: the ignore set is given (per-file stage 2/3 calls) */
    const uint8_t *d_contained_in;/* [n_reads] or NULL: the contained set is given                */
} fuz_ovlp_input;

typedef struct {
    uint8_t *d_ignore;            /* [n_reads] stage 1: reads to ignore                            */
    uint8_t *d_contained;         /* [n_reads] stage 2: contained reads                            */
    int64_t cap_groups;           /* a group = a run of one q among the lines passing the phase filter */
    int32_t *d_grp_q;             /* [cap_groups] q of the group                                   */
    int32_t *d_grp_line;          /* [cap_groups] input line that opens the group                  */
    uint8_t *d_grp_ignore;        /* [cap_groups] stage-1 verdict of the run (:80-87)              */
    uint8_t *d_grp_tie;           /* [cap_groups] 1: two candidates tie on (inphase, len, range, t):
                                   * the host re-sorts the group with the full string comparison   */
    int32_t *d_grp_off;           /* [cap_groups + 1] slice of d_out_line of every group           */
    int64_t cap_out;
    int32_t *d_out_line;          /* [cap_out] stage 3: selected lines (index into the input) in output order */
    uint8_t *d_cand;              /* [n_ovl] or NULL: stage-3 candidates, 1 = 5' list, 2 = 3' list (:265-273);
                                   * what the host needs to re-sort a group flagged in d_grp_tie      */
} fuz_ovlp_outputs;

/* status after the call: reserved[0] = groups, reserved[1] = selected lines, reserved[2] = selected
 * lines needed.  FUZ_E_CAPACITY: error_index 10 = cap_groups, 11 = cap_out, 9 = more than 512
 * candidates on one side of a read (max_cov beyond what the kernel holds in shared memory). */
int fuz_ovlp_filter(fuz_ctx *ctx, const fuz_ovlp_input *in, fuz_ovlp_outputs *out);

/* ---- LA4Falcon text -> columns on the device (rr_hctg_track.py:38-44, ovlp_filter_with_phase.py:60-62,95-99) ---- */
/* Device form of fuz_host_parse_la4falcon(_mo): d_text = the concatenated output of `LA4Falcon -m / -mo`.
 * Fills the columns of fuz_rr_input / fuz_ovlp_input (same meaning as the host parsers), d_line_off / d_line_len =
 * place of every line in the text; blank lines are skipped.  d_flags bit 7: the identity column of that line is
 * not a plain decimal of at most 15 significant digits and has to be evaluated by the host (float(col 3) < 90).
 * Asynchronous; status.reserved[0] = lines, reserved[1] = lines flagged for the host.  FUZ_E_CAPACITY (index 12):
 * more than `cap` lines; FUZ_E_FORMAT: a line the reference would raise on (error_index = line, reserved[3] =
 * 1 fewer than 12 columns, 2 not an integer, 3 id not a %09d id (require_id9), 4 identity column). */
int fuz_parse_la4falcon(fuz_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, int64_t cap, int32_t require_id9,
                        int32_t *d_q, int32_t *d_t, int32_t *d_len, int32_t *d_qs, int32_t *d_qe, int32_t *d_ql,
                        int32_t *d_ts, int32_t *d_te, int32_t *d_tl, uint8_t *d_flags, int64_t *d_line_off,
                        int32_t *d_line_len);

/* ---- host helpers (no CUDA) ------------------------------------------------------ */
/* Walk the block_size chain of a record buffer.  rec_off needs n_rec+1 slots; returns
 * the record count through n_rec (call with rec_off = NULL to count only). */
int fuz_host_index_records(const uint8_t *h_rec_buf, int64_t rec_bytes, int64_t *h_rec_off,
                           int64_t cap_rec, int64_t *n_rec);
/* first-seen QNAME -> q_id per contig (phasing.py:47-54).  h_name_first[q] receives the
 * index of the first record carrying q_id q of its contig, at ctg_q_off[c] + q. */
int fuz_host_assign_qids(const uint8_t *h_rec_buf, const int64_t *h_rec_off, int64_t n_rec,
                         const int32_t *h_ctg_rec_off, int32_t n_ctg, int32_t *h_rec_qid,
                         int32_t *h_ctg_nq, int64_t *h_name_first);
/* Parse LA4Falcon -m text (rr_hctg_track.py:38-44): columns 0, 1, 2, 11 of every line ->
 * q, t, len = -int(col 2), tlen.  Returns the number of lines parsed (stops at cap), or -1
 * on a malformed line (the reference would raise). */
int64_t fuz_host_parse_la4falcon(const char *text, int64_t n_bytes, int64_t cap,
                                 int32_t *q, int32_t *t, int32_t *len, int32_t *tlen);
/* Parse LA4Falcon -mo text (ovlp_filter_with_phase.py:60-62,95-99): every column the three stages
 * read, the flags of fuz_ovlp_input.d_flags and the place of the line in `text`.  Returns the number
 * of lines, -1 on a malformed line (the reference would raise), -2 when a read id is not a %09d id. */
int64_t fuz_host_parse_la4falcon_mo(const char *text, int64_t n_bytes, int64_t cap, int32_t *q, int32_t *t, int32_t *len,
                                    int32_t *qs, int32_t *qe, int32_t *ql, int32_t *ts, int32_t *te, int32_t *tl,
                                    uint8_t *flags, int64_t *line_off, int32_t *line_len);
/* Output text of the filter (:266-275, :352): tokens of every selected line joined by blanks plus the
 * phase strings of q and t (phase_text[phase_off[r] .. phase_off[r+1]) = "ctg.block.phase").  Returns
 * the size (call with out = NULL first), -1 if cap is too small. */
int64_t fuz_host_format_ovlp(const char *text, const int64_t *line_off, const int32_t *line_len, const int32_t *q,
                             const int32_t *t, const int64_t *sel, int64_t n_sel, const char *phase_text,
                             const int64_t *phase_off, char *out, int64_t cap);
/* Text of het_call/variant_map (phasing.py:126,128) for the rows [v0, v1) and of g_atable/atable (phasing.py:199) for
 * the rows [a0, a1) of the row arrays (fuz_outputs layout).  Return the size of the text; -1 if cap is too small
 * (32 bytes per variant_map row, 96 per atable row suffice), -2 if a position lies outside ref_seq (the reference
 * raises IndexError, phasing.py:123) or an allele is not one of A, C, G, T. */
int64_t fuz_host_format_variant_map(const int32_t *site_pos, const int32_t *vm_site, const uint8_t *vm_base,
                                    const int32_t *vm_qid, int64_t v0, int64_t v1, const char *ref_seq, int64_t ref_len,
                                    char *out, int64_t cap);
int64_t fuz_host_format_atable(const int32_t *site_pos, const uint8_t *site_al, const int32_t *at_s1, const int32_t *at_s2,
                               const int32_t *at_ct, int64_t a0, int64_t a1, char *out, int64_t cap);
/* Text of phased_reads of one contig (phasing.py:465-480): vm_qid = the contig's variant_map rows (they fix the row
 * order: CPython-2 dict of int keys inserted at first appearance), pr_* = its phased_reads rows sorted
 * by q_id, names = QNAME of every q_id.  With out = NULL returns an upper bound of the size; else the size, or -1. */
int64_t fuz_host_format_phased_reads(const int32_t *vm_qid, int64_t n_vm, const int32_t *pr_qid, const int32_t *pr_block,
                                     const int32_t *pr_phase, const int32_t *pr_n0, const int32_t *pr_n1, int64_t n_pr,
                                     const char *ctg_id, const char *name_blob, const int64_t *name_off, int64_t n_names,
                                     char *out, int64_t cap);
/* The same with the QNAMEs as fixed-width rows (NUL padded, `width` bytes each: the layout the device gathers them in). */
int64_t fuz_host_format_phased_reads_rows(const int32_t *vm_qid, int64_t n_vm, const int32_t *pr_qid, const int32_t *pr_block,
                                          const int32_t *pr_phase, const int32_t *pr_n0, const int32_t *pr_n1, int64_t n_pr,
                                          const char *ctg_id, const char *name_rows, int64_t width, int64_t n_names,
                                          char *out, int64_t cap);
/* Text of het_call/q_id_map (phasing.py:132-134) from fixed-width QNAME rows: "q_id qname" for q_id = 0 .. n-1.
 * cap >= n * (width + 13).  Returns the size, -1 on bad arguments. */
int64_t fuz_host_format_q_id_map_rows(const char *name_rows, int64_t width, int64_t n, char *out, int64_t cap);
/* Text of het_call/variant_pos (phasing.py:116-124) for the sites [s0, s1): "pos ref total b0 c0 b1 c1 b2 c2 b3 c3", bases
 * by descending (count, base); site_cnt holds 4 counts per site in A, C, G, T order.  cap >= 80 bytes per row.
 * Returns the size, -1 on bad arguments, -2 if a position lies outside ref_seq (IndexError in the reference, :123). */
int64_t fuz_host_format_variant_pos(const int32_t *site_pos, const int32_t *site_cnt, int64_t s0, int64_t s1,
                                    const char *ref_seq, int64_t ref_len, char *out, int64_t cap);
/* Text of get_phased_blocks/phased_variants (phasing.py:411-421) for the sites [s0, s1) of one contig: per block id in
 * ascending order a P row ("P pid min max span n span/n", the quotient as Python 2 prints a float) and the V rows of its
 * sites in position order.  cap >= 64 + 224 bytes per site.  Returns the size, -1 on bad arguments, -2 on a position
 * outside ref_seq or an invalid phase / allele. */
int64_t fuz_host_format_phased_variants(const int32_t *site_pos, const uint8_t *site_al, const int32_t *ph_block,
                                        const uint8_t *ph_state, const int32_t *ph_lext, const int32_t *ph_rext,
                                        const int32_t *ph_lscore, const int32_t *ph_rscore, int64_t s0, int64_t s1,
                                        const char *ref_seq, int64_t ref_len, char *out, int64_t cap);
/* CPython-2.7 dict / set iteration order of str keys (Objects/stringobject.c string_hash + the insert-only table
 * of dictobject.c): keys[i] = blob[off[i], off[i+1]) inserted in order, duplicates ignored;
 * out = index of every distinct key in iteration order.  Returns the number of distinct keys. */
int64_t fuz_host_py27_str_dict_order(const char *blob, const int64_t *off, int64_t n, int64_t *out);
/* b-reads in the iteration order of the reference's bread_to_areads dict (rr_hctg_track.py:97-100,113) from the
 * target / LAS file of every KEPT overlap line in (file, line) order.  Returns the number of b-reads. */
int64_t fuz_host_rr_bread_order(const int32_t *t_kept, const int32_t *file_kept, int64_t n, int32_t *out);
/* Text of rawread_to_contigs (rr_hctg_track.py:126-138) for the given b-reads from the vote rows of fuz_rr_track,
 * the contig names (ctg_blob / ctg_off) and the rid -> contigs table of fuz_rr_input.  Returns the size (call with
 * out = NULL first), -1 if cap is too small. */
int64_t fuz_host_rr_format_rows(const int32_t *breads, int64_t n_breads, const int32_t *vt_off, const int32_t *vt_ctg,
                                const int32_t *vt_count, const int64_t *vt_score, const char *ctg_blob,
                                const int64_t *ctg_off, const uint8_t *in_map, const int32_t *rc_off,
                                const int32_t *rc_ctg, char *out, int64_t cap);
/* CPython-2.7 dict iteration order of int keys inserted in the given order (B.3). */
int fuz_host_py27_int_dict_order(const int64_t *keys, int64_t n, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* FUZ_H_ */
