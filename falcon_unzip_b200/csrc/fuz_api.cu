// Fused batch entry (device pointers) and the host-buffer entry that the reference-facing
// Python layer calls (H2D, the four stages, D2H of the filled row prefixes).
#include <string.h>

#include "fuz_internal.cuh"

int fuz_het_call_impl(fuz_ctx *ctx, const fuz_batch *in, fuz_outputs *out);
int fuz_association_impl(fuz_ctx *ctx, int32_t n_ctg, fuz_outputs *out, bool row_off_valid);
int fuz_blocks_impl(fuz_ctx *ctx, int32_t n_ctg, fuz_outputs *out, bool at_off_valid);
int fuz_reads_csr(fuz_ctx *ctx, int32_t n_ctg, const int32_t *d_ctg_nq, int64_t total_nq, fuz_outputs *out, bool dup_valid);
int fuz_reads_vote(fuz_ctx *ctx, int32_t n_ctg, int64_t total_nq, fuz_outputs *out);
int fuz_assign_qids_impl(fuz_ctx *ctx, const uint8_t *d_rec_buf, const int64_t *d_rec_off, int32_t n_rec, int64_t rec_bytes,
                         const int32_t *d_ctg_rec_off, int32_t n_ctg, int32_t *d_rec_qid, int32_t *d_ctg_nq,
                         int64_t *d_name_first, int32_t *d_ctg_slots, uint8_t *scratch);
size_t fuz_qid_scratch_bytes(int32_t n_rec);

extern "C" int fuz_phase_batch(fuz_ctx *ctx, const fuz_batch *in_, fuz_outputs *out) {
    if (!ctx || !in_ || !out) return FUZ_E_ARG;
    fuz_batch b = *in_;
    const fuz_batch *in = &b;
    int rc;
    if (!b.d_rec_qid) {
        // q_ids assigned here (phasing.py:47-54).  The read stage then indexes its per-read tables by
        // (first record of the contig + q_id): one slot per record, so no count has to come back to
        // the host before the later stages are launched.
        if (b.n_rec < 0 || b.n_ctg < 1 || !b.d_ctg_rec_off) return fuz_fail(ctx, FUZ_E_ARG, "fuz_phase_batch: empty batch");
        const size_t sz_q = ((size_t)(b.n_rec + 1) * 4 + 255) & ~(size_t)255, sz_c = ((size_t)b.n_ctg * 4 + 255) & ~(size_t)255;
        const size_t need = sz_q + 2 * sz_c + fuz_qid_scratch_bytes(b.n_rec);
        if (need > ctx->qid_cap) {
            FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            if (ctx->qid_buf) FUZ_CUDA(ctx, cudaFree(ctx->qid_buf));
            ctx->qid_buf = nullptr; ctx->qid_cap = 0;
            cudaError_t e = cudaMalloc(&ctx->qid_buf, need + (need >> 2));
            if (e != cudaSuccess) return fuz_fail(ctx, FUZ_E_CUDA, "q_id buffer of %zu bytes: %s", need, cudaGetErrorString(e));
            ctx->qid_cap = need + (need >> 2);
        }
        int32_t *qid = reinterpret_cast<int32_t *>(ctx->qid_buf), *slots = reinterpret_cast<int32_t *>(ctx->qid_buf + sz_q);
        int32_t *nq = out->d_ctg_nq ? out->d_ctg_nq : reinterpret_cast<int32_t *>(ctx->qid_buf + sz_q + sz_c);
        uint8_t *scratch = ctx->qid_buf + sz_q + 2 * sz_c;
        // independent of the projection and the pileup: run on the side stream, join before k_signature
        const bool side = !ctx->profile;
        cudaStream_t main_stream = ctx->stream;
        if (side) {
            if (!ctx->aux_stream) {
                FUZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
                FUZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
                FUZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
            }
            FUZ_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
            FUZ_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
            ctx->stream = ctx->aux_stream;
        }
        rc = fuz_assign_qids_impl(ctx, b.d_rec_buf, b.d_rec_off, b.n_rec, b.rec_bytes, b.d_ctg_rec_off, b.n_ctg, qid, nq,
                                  out->d_name_first, slots, scratch);
        ctx->stream = main_stream;
        if (rc) return rc;
        if (side) {
            FUZ_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->aux_stream));
            ctx->join_pending = true;
        }
        b.d_rec_qid = qid; b.d_ctg_nq = slots; b.total_nq = b.n_rec;
    }
    rc = fuz_het_call_impl(ctx, in, out);
    if (ctx->join_pending) {                  // het_call_impl joins before k_signature; error paths join here
        cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
        ctx->join_pending = false;
    }
    if (rc) return rc;
    // the row range of every site (het call) and the duplicate flags (association) stay in
    // the context's inter-stage buffer and are reused by the later stages
    if ((rc = fuz_association_impl(ctx, in->n_ctg, out, true))) return rc;
    // The per-read lists of the read stage need the variant_map rows and the duplicate flags only: they are
    // built on the side stream while the block stage (one CTA per contig, most SMs idle) runs on the main one.
    const bool side = !ctx->profile;
    cudaStream_t main_stream = ctx->stream;
    if (side) {
        if (!ctx->aux_stream) {
            FUZ_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
            FUZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
            FUZ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        }
        FUZ_CUDA(ctx, cudaEventRecord(ctx->ev_fork, main_stream));
        FUZ_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
        ctx->stream = ctx->aux_stream;
    }
    rc = fuz_reads_csr(ctx, in->n_ctg, in->d_ctg_nq, in->total_nq, out, true);
    ctx->stream = main_stream;
    if (side) cudaEventRecord(ctx->ev_join, ctx->aux_stream);
    const int rc_blocks = rc ? rc : fuz_blocks_impl(ctx, in->n_ctg, out, true);
    if (side) cudaStreamWaitEvent(main_stream, ctx->ev_join, 0);
    if (rc_blocks) return rc_blocks;
    return fuz_reads_vote(ctx, in->n_ctg, in->total_nq, out);
}

static int ensure_stage(fuz_ctx *ctx, size_t dev_bytes, size_t pin_bytes) {
    if (dev_bytes > ctx->stage_dev_cap) {
        FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->stage_dev) FUZ_CUDA(ctx, cudaFree(ctx->stage_dev));
        ctx->stage_dev = nullptr; ctx->stage_dev_cap = 0;
        size_t want = dev_bytes + (dev_bytes >> 3) + (1 << 20);
        cudaError_t e = cudaMalloc(&ctx->stage_dev, want);
        if (e != cudaSuccess) return fuz_fail(ctx, FUZ_E_CUDA, "device staging of %zu bytes: %s", want, cudaGetErrorString(e));
        ctx->stage_dev_cap = want;
    }
    if (pin_bytes > ctx->stage_pin_cap) {
        FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->stage_pin) FUZ_CUDA(ctx, cudaFreeHost(ctx->stage_pin));
        ctx->stage_pin = nullptr; ctx->stage_pin_cap = 0;
        size_t want = pin_bytes + 4096;
        cudaError_t e = cudaMallocHost(&ctx->stage_pin, want);
        if (e != cudaSuccess) return fuz_fail(ctx, FUZ_E_CUDA, "pinned staging of %zu bytes: %s", want, cudaGetErrorString(e));
        ctx->stage_pin_cap = want;
    }
    return FUZ_OK;
}

// ---------------------------------------------------------------- selective record fetch
// Host-buffer entry, records in page-locked host memory: the kernels need the fixed header,
// read name, CIGAR and SEQ of a BAM record -- not QUAL (one byte per base, ~2/3 of a PacBio
// record) or the tags.  Instead of a cudaMemcpy of the whole buffer, one warp per record
// reads exactly those bytes through the host mapping (128-bit loads, 16-byte lines) and writes
// them at the SAME offsets of the device buffer, so every later kernel sees an ordinary
// record buffer whose unread parts are simply not populated.
// The first 512 bytes from the 16-byte line holding the record start are always fetched (they
// carry the header, from which the needed length follows); the rest up to the end of SEQ.
#define FUZ_FETCH_SLAB 512
__global__ void __launch_bounds__(256) k_fetch_records(const uint8_t *__restrict__ h_src, uint8_t *__restrict__ d_dst,
                                                       const int64_t *__restrict__ rec_off, int n_rec, int64_t rec_bytes,
                                                       unsigned long long *fetched) {
    fuz_pdl_enter();
    __shared__ __align__(16) uint8_t head[8][64];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int64_t whole = rec_bytes & ~(int64_t)15;               // 16-byte lines fully inside the buffer
    unsigned long long bytes = 0;
    for (int r = warp_g; r < n_rec; r += n_warps) {
        const int64_t off_r = rec_off[r], off_n = rec_off[r + 1];
        if (off_r < 0 || off_n < off_r || off_n > rec_bytes) continue;      // the record kernels report it
        const int64_t a0 = off_r & ~(int64_t)15;
        // slab: 32 lines from a0
        const int64_t s_off = a0 + 16 * lane;
        uint4 v = make_uint4(0, 0, 0, 0);
        const bool s_in = s_off + 16 <= whole;
        if (s_in) {
            v = __ldcs(reinterpret_cast<const uint4 *>(h_src + s_off));
            *reinterpret_cast<uint4 *>(d_dst + s_off) = v;
        }
        if (lane < 4) *reinterpret_cast<uint4 *>(&head[wib][16 * lane]) = v;
        __syncwarp();
        // header fields (record-relative offsets 12, 16, 20; the record start is not aligned)
        const int sh = (int)(off_r - a0);
        auto u32_at = [&](int o) {
            const uint8_t *b = &head[wib][sh + o];
            return (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 24);
        };
        const int64_t l_name = u32_at(12) & 0xFF, n_cig = u32_at(16) & 0xFFFF;
        const int64_t l_seq = (int32_t)u32_at(20);
        __syncwarp();
        int64_t need = 36 + l_name + 4 * n_cig + (l_seq > 0 ? (l_seq + 1) / 2 : 0);
        need = min(max(need, (int64_t)36), off_n - off_r);         // a bad header: the validation sees the record as it is
        int64_t end = min((off_r + need + 15) & ~(int64_t)15, whole);
        const int64_t body = a0 + FUZ_FETCH_SLAB;
        // body: 4 lines per lane in flight
        for (int64_t o = body + 16 * lane; o < end; o += 4 * 512) {
            uint4 w[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (o + 512 * j < end) w[j] = __ldcs(reinterpret_cast<const uint4 *>(h_src + o + 512 * j));
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (o + 512 * j < end) *reinterpret_cast<uint4 *>(d_dst + o + 512 * j) = w[j];
        }
        if (lane == 0) bytes += (unsigned long long)(max(end, min(body, whole)) - a0);
        // tail of the buffer that is not a whole line (last record only)
        if (off_r + need > whole && lane == 0)
            for (int64_t o = max(whole, a0); o < min(off_r + need, rec_bytes); o++) { d_dst[o] = h_src[o]; bytes++; }
    }
    if (lane == 0 && bytes) atomicAdd(fetched, bytes);
}

extern "C" int fuz_phase_batch_host(fuz_ctx *ctx, const fuz_host_batch *in, fuz_host_outputs *out, fuz_status *h_status,
                                    int64_t *h2d_bytes, int64_t *d2h_bytes) {
    if (!ctx || !in || !out || !h_status) return FUZ_E_ARG;
    if (in->n_ctg < 1 || in->n_rec < 0) return fuz_fail(ctx, FUZ_E_ARG, "fuz_phase_batch_host: empty batch");
    const int n_ctg = in->n_ctg, n_rec = in->n_rec;
    cudaStream_t st = ctx->stream;
    // tile-aligned global offsets (host, tiny)
    int64_t total_nq = 0;
    FuzLayout P;   // pinned
    size_t p_goff = P.add(8 * (size_t)(n_ctg + 1)), p_fetched = P.add(8);
    const bool dev_qid = in->h_rec_qid == nullptr;
    FuzLayout D;   // device
    size_t d_rec = D.add((size_t)in->rec_bytes + 64), d_off = D.add(8 * (size_t)(n_rec + 1)), d_qid = D.add(4 * (size_t)(n_rec + 1));
    size_t d_cro = D.add(4 * (size_t)(n_ctg + 1)), d_clen = D.add(4 * (size_t)n_ctg), d_goff = D.add(8 * (size_t)(n_ctg + 1));
    size_t d_cnq = D.add(4 * (size_t)n_ctg), d_fetched = D.add(8);
    size_t d_nfirst = D.add(8 * (size_t)(n_rec + 1));
    const int64_t cs = out->cap_sites, cv = out->cap_vmap, ca = out->cap_atable, cr = out->cap_reads;
    size_t o_sctg = D.add(4 * (size_t)cs), o_spos = D.add(4 * (size_t)cs), o_scnt = D.add(16 * (size_t)cs);
    size_t o_sal = D.add(2 * (size_t)cs), o_stop = D.add(2 * (size_t)cs);
    size_t o_vs = D.add(4 * (size_t)cv), o_vq = D.add(4 * (size_t)cv), o_vb = D.add((size_t)cv);
    size_t o_a1 = D.add(4 * (size_t)ca), o_a2 = D.add(4 * (size_t)ca), o_act = D.add(16 * (size_t)ca);
    size_t o_pst = D.add((size_t)cs), o_ple = D.add(4 * (size_t)cs), o_pre = D.add(4 * (size_t)cs);
    size_t o_pls = D.add(4 * (size_t)cs), o_prs = D.add(4 * (size_t)cs), o_pb = D.add(4 * (size_t)cs);
    size_t o_rc = D.add(4 * (size_t)cr), o_rq = D.add(4 * (size_t)cr), o_rb = D.add(4 * (size_t)cr);
    size_t o_rp = D.add(4 * (size_t)cr), o_r0 = D.add(4 * (size_t)cr), o_r1 = D.add(4 * (size_t)cr);
    int rc = ensure_stage(ctx, D.off, P.off);
    if (rc) return rc;
    int64_t *goff = reinterpret_cast<int64_t *>(ctx->stage_pin + p_goff);
    goff[0] = 0;
    for (int c = 0; c < n_ctg; c++) {
        if (in->h_ctg_len[c] < 0) return fuz_fail(ctx, FUZ_E_ARG, "negative contig length");
        int64_t padded = ((int64_t)in->h_ctg_len[c] + FUZ_TILE - 1) / FUZ_TILE * FUZ_TILE;
        if (padded == 0) padded = FUZ_TILE;
        goff[c + 1] = goff[c] + padded;
        if (!dev_qid) total_nq += in->h_ctg_nq[c];
    }
    uint8_t *dv = ctx->stage_dev;
    int64_t up = 0;
    auto h2d = [&](size_t off, const void *src, size_t bytes) -> cudaError_t {
        up += (int64_t)bytes;
        return bytes ? cudaMemcpyAsync(dv + off, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
    };
    // records in page-locked memory that the device can address: fetch only what the kernels read
    const uint8_t *mapped = nullptr;
    if (ctx->host_fetch && in->rec_bytes > 0 && n_rec > 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, in->h_rec_buf) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
            mapped = static_cast<const uint8_t *>(pa.devicePointer);
        else
            (void)cudaGetLastError();
    }
    FUZ_CUDA(ctx, h2d(d_off, in->h_rec_off, 8 * (size_t)(n_rec + 1)));
    if (mapped) {
        FUZ_CUDA(ctx, cudaMemsetAsync(dv + d_fetched, 0, 8, st));
        fuz_launch(ctx, k_fetch_records, ctx->fetch_ctas, 256, 0, st, mapped, dv + d_rec, reinterpret_cast<const int64_t *>(dv + d_off),
                                                             n_rec, in->rec_bytes,
                                                             reinterpret_cast<unsigned long long *>(dv + d_fetched));
        FUZ_LAUNCH_CHECK(ctx, "k_fetch_records");
        FUZ_CUDA(ctx, cudaMemcpyAsync(ctx->stage_pin + p_fetched, dv + d_fetched, 8, cudaMemcpyDeviceToHost, st));
    } else {
        FUZ_CUDA(ctx, h2d(d_rec, in->h_rec_buf, (size_t)in->rec_bytes));
    }
    FUZ_CUDA(ctx, cudaMemsetAsync(dv + d_rec + in->rec_bytes, 0, 64, st));
    if (!dev_qid) FUZ_CUDA(ctx, h2d(d_qid, in->h_rec_qid, 4 * (size_t)n_rec));
    FUZ_CUDA(ctx, h2d(d_cro, in->h_ctg_rec_off, 4 * (size_t)(n_ctg + 1)));
    FUZ_CUDA(ctx, h2d(d_clen, in->h_ctg_len, 4 * (size_t)n_ctg));
    FUZ_CUDA(ctx, h2d(d_goff, goff, 8 * (size_t)(n_ctg + 1)));
    if (!dev_qid) FUZ_CUDA(ctx, h2d(d_cnq, in->h_ctg_nq, 4 * (size_t)n_ctg));
    fuz_batch b;
    memset(&b, 0, sizeof(b));
    b.n_ctg = n_ctg; b.n_rec = n_rec; b.rec_bytes = in->rec_bytes;
    b.d_rec_buf = dv + d_rec; b.d_rec_off = reinterpret_cast<int64_t *>(dv + d_off);
    b.d_rec_qid = reinterpret_cast<int32_t *>(dv + d_qid); b.d_ctg_rec_off = reinterpret_cast<int32_t *>(dv + d_cro);
    b.d_ctg_len = reinterpret_cast<int32_t *>(dv + d_clen); b.d_ctg_goff = reinterpret_cast<int64_t *>(dv + d_goff);
    b.d_ctg_nq = reinterpret_cast<int32_t *>(dv + d_cnq);
    b.total_glen = goff[n_ctg]; b.total_nq = total_nq;
    if (dev_qid) b.d_rec_qid = nullptr;       // fuz_phase_batch assigns the q_ids on the device
    fuz_outputs o;
    memset(&o, 0, sizeof(o));
    o.cap_sites = cs; o.cap_vmap = cv; o.cap_atable = ca; o.cap_reads = cr;
#define DP(T, off) reinterpret_cast<T *>(dv + (off))
    o.d_site_ctg = DP(int32_t, o_sctg); o.d_site_pos = DP(int32_t, o_spos); o.d_site_cnt = DP(int32_t, o_scnt);
    o.d_site_al = DP(uint8_t, o_sal); o.d_site_top = DP(uint8_t, o_stop);
    o.d_vm_site = DP(int32_t, o_vs); o.d_vm_qid = DP(int32_t, o_vq); o.d_vm_base = DP(uint8_t, o_vb);
    o.d_at_s1 = DP(int32_t, o_a1); o.d_at_s2 = DP(int32_t, o_a2); o.d_at_ct = DP(int32_t, o_act);
    o.d_ph_state = DP(uint8_t, o_pst); o.d_ph_lext = DP(int32_t, o_ple); o.d_ph_rext = DP(int32_t, o_pre);
    o.d_ph_lscore = DP(int32_t, o_pls); o.d_ph_rscore = DP(int32_t, o_prs); o.d_ph_block = DP(int32_t, o_pb);
    o.d_pr_ctg = DP(int32_t, o_rc); o.d_pr_qid = DP(int32_t, o_rq); o.d_pr_block = DP(int32_t, o_rb);
    o.d_pr_phase = DP(int32_t, o_rp); o.d_pr_n0 = DP(int32_t, o_r0); o.d_pr_n1 = DP(int32_t, o_r1);
#undef DP
    o.d_counts = nullptr;
    if (dev_qid) { o.d_ctg_nq = reinterpret_cast<int32_t *>(dv + d_cnq); o.d_name_first = reinterpret_cast<int64_t *>(dv + d_nfirst); }
    if ((rc = fuz_phase_batch(ctx, &b, &o))) return rc;
    rc = fuz_get_status(ctx, h_status);          // synchronises; row counts now known
    if (mapped) up += *reinterpret_cast<const int64_t *>(ctx->stage_pin + p_fetched);
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = (int64_t)sizeof(fuz_status);
    if (rc) return rc;
    int64_t down = (int64_t)sizeof(fuz_status);
    auto d2h = [&](void *dst, size_t off, size_t bytes) -> cudaError_t {
        down += (int64_t)bytes;
        return (bytes && dst) ? cudaMemcpyAsync(dst, dv + off, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess;
    };
    const size_t ns = (size_t)h_status->n_sites, nv = (size_t)h_status->n_vmap, na = (size_t)h_status->n_atable,
                 nr = (size_t)h_status->n_reads;
    FUZ_CUDA(ctx, d2h(out->site_ctg, o_sctg, 4 * ns)); FUZ_CUDA(ctx, d2h(out->site_pos, o_spos, 4 * ns));
    FUZ_CUDA(ctx, d2h(out->site_cnt, o_scnt, 16 * ns)); FUZ_CUDA(ctx, d2h(out->site_al, o_sal, 2 * ns));
    FUZ_CUDA(ctx, d2h(out->site_top, o_stop, 2 * ns));
    FUZ_CUDA(ctx, d2h(out->vm_site, o_vs, 4 * nv)); FUZ_CUDA(ctx, d2h(out->vm_qid, o_vq, 4 * nv));
    FUZ_CUDA(ctx, d2h(out->vm_base, o_vb, nv));
    FUZ_CUDA(ctx, d2h(out->at_s1, o_a1, 4 * na)); FUZ_CUDA(ctx, d2h(out->at_s2, o_a2, 4 * na));
    FUZ_CUDA(ctx, d2h(out->at_ct, o_act, 16 * na));
    FUZ_CUDA(ctx, d2h(out->ph_state, o_pst, ns)); FUZ_CUDA(ctx, d2h(out->ph_lext, o_ple, 4 * ns));
    FUZ_CUDA(ctx, d2h(out->ph_rext, o_pre, 4 * ns)); FUZ_CUDA(ctx, d2h(out->ph_lscore, o_pls, 4 * ns));
    FUZ_CUDA(ctx, d2h(out->ph_rscore, o_prs, 4 * ns)); FUZ_CUDA(ctx, d2h(out->ph_block, o_pb, 4 * ns));
    FUZ_CUDA(ctx, d2h(out->pr_ctg, o_rc, 4 * nr)); FUZ_CUDA(ctx, d2h(out->pr_qid, o_rq, 4 * nr));
    FUZ_CUDA(ctx, d2h(out->pr_block, o_rb, 4 * nr)); FUZ_CUDA(ctx, d2h(out->pr_phase, o_rp, 4 * nr));
    FUZ_CUDA(ctx, d2h(out->pr_n0, o_r0, 4 * nr)); FUZ_CUDA(ctx, d2h(out->pr_n1, o_r1, 4 * nr));
    if (dev_qid) {
        FUZ_CUDA(ctx, d2h(out->ctg_nq, d_cnq, 4 * (size_t)n_ctg));
        FUZ_CUDA(ctx, d2h(out->name_first, d_nfirst, 8 * (size_t)n_rec));
    }
    FUZ_CUDA(ctx, cudaStreamSynchronize(st));
    if (d2h_bytes) *d2h_bytes = down;
    return FUZ_OK;
}
