// Raw-read -> haplotig tracking on the device: reference falcon_unzip/rr_hctg_track.py
//   R1 overlap filter with table probes          tr_stage1            :38-57
//   R2 per-target top-bestn, exact heapq replay  tr_stage1 :59-63 + merge of run_track_reads :97-105
//   R3 contig vote per target read               run_track_reads      :113-123
// The reference's result depends on the *array order* of CPython's heapq (rank ties between a
// primary contig and its haplotig are broken by dict insertion order), so R2
// replays heappush / heappushpop exactly: the kept overlaps of a target are grouped with
// atomics, put back into file order, and pushed by one thread per target.
#include "fuz_internal.cuh"

namespace {

struct RRScratch {
    int32_t *t_cnt, *t_off, *t_cur, *grp, *a_len, *a_q, *vt_cnt;
};

__global__ void k_rr_init(fuz_status *st, RRScratch R, int n_reads) {
    fuz_pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { st->error = 0; st->error_index = 0; st->reserved[1] = st->reserved[2] = st->reserved[3] = 0; }
    for (; i <= n_reads; i += gridDim.x * blockDim.x) { R.t_cnt[i] = 0; if (i < n_reads) R.t_cur[i] = 0; }
}

// R1: rr_hctg_track.py:45-57
__global__ void k_rr_filter(fuz_rr_input in, fuz_rr_outputs out, RRScratch R, fuz_status *st) {
    fuz_pdl_enter();
    long long kept = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < in.n_ovl; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = in.d_q[i], t = in.d_t[i];
        bool keep = false;
        if (q < 0 || q >= in.n_reads || t < 0 || t >= in.n_reads) {
            fuz_raise(st, FUZ_E_FORMAT, (int)i);                    // rid_to_phase[int(t_id)] would raise IndexError
        } else if (in.d_tlen[i] >= in.min_len && in.d_in_map[q]) {
            keep = true;
            const int tc = in.d_ph_ctg[t];
            if (tc >= 0 && in.d_ph_block[t] != -1) {
                const int qc = in.d_ph_ctg[q];
                if (qc >= 0 && qc == tc && in.d_ph_block[q] == in.d_ph_block[t] && in.d_ph_phase[q] != in.d_ph_phase[t])
                    keep = false;
            }
        }
        out.d_keep[i] = keep ? 1 : 0;
        if (keep) { atomicAdd(&R.t_cnt[t], 1); kept++; }
    }
    if (kept) atomicAdd((unsigned long long *)&st->reserved[3], (unsigned long long)kept);
}

__global__ void k_rr_fill(fuz_rr_input in, fuz_rr_outputs out, RRScratch R, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < in.n_ovl; i += (int64_t)gridDim.x * blockDim.x) {
        if (!out.d_keep[i]) continue;
        const int t = in.d_t[i];
        R.grp[R.t_off[t] + atomicAdd(&R.t_cur[t], 1)] = (int)i;
    }
}

// ---- CPython heapq on (len, q) tuples (Lib/heapq.py: heappush, heappushpop, _siftdown, _siftup)
__device__ __forceinline__ bool tup_lt(int l1, int q1, int l2, int q2) { return l1 < l2 || (l1 == l2 && q1 < q2); }

__device__ void hq_siftdown(int *hl, int *hq, int startpos, int pos) {
    const int nl = hl[pos], nq = hq[pos];
    while (pos > startpos) {
        const int parent = (pos - 1) >> 1;
        if (tup_lt(nl, nq, hl[parent], hq[parent])) { hl[pos] = hl[parent]; hq[pos] = hq[parent]; pos = parent; continue; }
        break;
    }
    hl[pos] = nl; hq[pos] = nq;
}
__device__ void hq_siftup(int *hl, int *hq, int n, int pos) {
    const int startpos = pos;
    const int nl = hl[pos], nq = hq[pos];
    int child = 2 * pos + 1;
    while (child < n) {
        const int right = child + 1;
        if (right < n && !tup_lt(hl[child], hq[child], hl[right], hq[right])) child = right;
        hl[pos] = hl[child]; hq[pos] = hq[child];
        pos = child;
        child = 2 * pos + 1;
    }
    hl[pos] = nl; hq[pos] = nq;
    hq_siftdown(hl, hq, startpos, pos);
}
// rr_hctg_track.py:59-63 / :102-105: push while fewer than bestn entries, pushpop afterwards
__device__ __forceinline__ void hq_offer(int *hl, int *hq, int &n, int bestn, int l, int q) {
    if (n < bestn) {
        hl[n] = l; hq[n] = q; n++;
        hq_siftdown(hl, hq, 0, n - 1);
    } else if (n > 0 && tup_lt(hl[0], hq[0], l, q)) {
        hl[0] = l; hq[0] = q;
        hq_siftup(hl, hq, n, 0);
    }
}

// in-place heapsort of the line indices of one target: back to file order
__device__ void sort_lines(int *a, int n) {
    for (int start = n / 2 - 1; start >= 0; start--) {
        int root = start;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= n) break;
            if (child + 1 < n && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            int tmp = a[root]; a[root] = a[child]; a[child] = tmp;
            root = child;
        }
    }
    for (int end = n - 1; end > 0; end--) {
        int tmp = a[0]; a[0] = a[end]; a[end] = tmp;
        int root = 0;
        for (;;) {
            int child = 2 * root + 1;
            if (child >= end) break;
            if (child + 1 < end && a[child] < a[child + 1]) child++;
            if (a[root] >= a[child]) break;
            tmp = a[root]; a[root] = a[child]; a[child] = tmp;
            root = child;
        }
    }
}

// R2: one thread per target read
__global__ void k_rr_replay(fuz_rr_input in, fuz_rr_outputs out, RRScratch R, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < in.n_reads; t += gridDim.x * blockDim.x) {
        const int n = R.t_cnt[t];
        out.d_hp_n[t] = 0;
        if (n == 0 || in.bestn < 1) continue;
        int *lines = R.grp + R.t_off[t];
        sort_lines(lines, n);
        int *al = R.a_len + (int64_t)t * in.bestn, *aq = R.a_q + (int64_t)t * in.bestn;        // this file's heap
        int *gl = out.d_hp_len + (int64_t)t * in.bestn, *gq = out.d_hp_q + (int64_t)t * in.bestn;  // merged heap
        int na = 0, ng = 0, cur_file = -1;
        for (int k = 0; k <= n; k++) {
            const int line = k < n ? lines[k] : -1;
            const int f = k < n ? in.d_file[line] : -2;
            if (f != cur_file) {                       // end of a LAS file: merge its heap, array order (:99-105)
                for (int j = 0; j < na; j++) hq_offer(gl, gq, ng, in.bestn, al[j], aq[j]);
                na = 0;
                cur_file = f;
            }
            if (k < n) hq_offer(al, aq, na, in.bestn, in.d_len[line], in.d_q[line]);
        }
        out.d_hp_n[t] = ng;
    }
}

// R3: contigs voted by the kept a-reads of a target, in insertion order (:113-123).  Up to FUZ_RR_MAXC distinct contigs
// are accumulated in a per-thread table; a target that votes for more (a repeat read among primaries and haplotigs)
// takes the unbounded path: distinct contigs are counted by first occurrence in the vote sequence and accumulated straight
// in the output rows, which the count pass reserved.
#define FUZ_RR_MAXC 64
__global__ void k_rr_vote(fuz_rr_input in, fuz_rr_outputs out, RRScratch R, int fill, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < in.n_reads; t += gridDim.x * blockDim.x) {
        const int ng = out.d_hp_n[t];
        int ctg[FUZ_RR_MAXC], cnt[FUZ_RR_MAXC];
        long long score[FUZ_RR_MAXC];
        int nc = 0;
        bool overflow = false;
        const int *gl = out.d_hp_len + (int64_t)t * in.bestn, *gq = out.d_hp_q + (int64_t)t * in.bestn;
        for (int j = 0; j < ng && !overflow; j++) {
            const int rid = gq[j], s = gl[j];
            for (int e = in.d_rc_off[rid]; e < in.d_rc_off[rid + 1]; e++) {      // CPython-2 set order (host)
                const int c = in.d_rc_ctg[e];
                int k = 0;
                while (k < nc && ctg[k] != c) k++;
                if (k == nc) {
                    if (nc == FUZ_RR_MAXC) { overflow = true; break; }
                    ctg[nc] = c; cnt[nc] = 0; score[nc] = 0; nc++;
                }
                score[k] += -(long long)s;                                        // ctg_score[ctg][0] += -s (:122)
                cnt[k] += 1;
            }
        }
        if (overflow) {
            const int64_t o = fill ? out.d_vt_off[t] : 0;
            nc = 0;
            for (int j = 0; j < ng; j++) {
                const int rid = gq[j], s = gl[j];
                for (int e = in.d_rc_off[rid]; e < in.d_rc_off[rid + 1]; e++) {
                    const int c = in.d_rc_ctg[e];
                    int k = -1;
                    if (fill) {                                                   // rows written so far, insertion order
                        for (int x = 0; x < nc && k < 0; x++)
                            if (o + x < out.cap_votes && out.d_vt_ctg[o + x] == c) k = x;
                    } else {                                                      // seen before in the vote sequence?
                        for (int j2 = 0; j2 <= j && k < 0; j2++) {
                            const int r2 = gq[j2], e_end = j2 < j ? in.d_rc_off[r2 + 1] : e;
                            for (int e2 = in.d_rc_off[r2]; e2 < e_end; e2++)
                                if (in.d_rc_ctg[e2] == c) { k = 0; break; }
                        }
                    }
                    if (k < 0) {
                        k = nc++;
                        if (fill && o + k < out.cap_votes) { out.d_vt_ctg[o + k] = c; out.d_vt_count[o + k] = 0; out.d_vt_score[o + k] = 0; }
                    }
                    if (fill && o + k < out.cap_votes) { out.d_vt_count[o + k] += 1; out.d_vt_score[o + k] += -(long long)s; }
                }
            }
            if (!fill) R.vt_cnt[t] = nc;
            continue;
        }
        if (!fill) { R.vt_cnt[t] = nc; continue; }
        const int64_t o = out.d_vt_off[t];
        for (int k = 0; k < nc; k++) {
            if (o + k < out.cap_votes) { out.d_vt_ctg[o + k] = ctg[k]; out.d_vt_count[o + k] = cnt[k]; out.d_vt_score[o + k] = score[k]; }
        }
    }
}

__global__ void k_rr_votes_total(fuz_rr_outputs out, int n_reads, fuz_status *st) {
    fuz_pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0 && !st->error) {
        int64_t total = out.d_vt_off[n_reads];
        st->reserved[2] = total;
        if (total > out.cap_votes) fuz_raise(st, FUZ_E_CAPACITY, 8); else st->reserved[1] = total;
    }
}

}  // namespace

extern "C" int fuz_rr_track(fuz_ctx *ctx, const fuz_rr_input *in, fuz_rr_outputs *out) {
    if (!ctx || !in || !out) return FUZ_E_ARG;
    if (in->n_ovl < 0 || in->n_ovl > 0x7fffffffLL || in->n_reads < 1 || in->bestn < 0)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_rr_track: bad sizes (n_ovl %lld, n_reads %d, bestn %d)", (long long)in->n_ovl,
                        in->n_reads, in->bestn);
    cudaStream_t st = ctx->stream;
    // plain launches here: with programmatic dependent launch the early-resident CTAs of the next
    // kernel slow the long, divergent replay kernel down (measured 0.65 -> 0.91 ms per call)
    struct PdlOff { fuz_ctx *c; int saved; PdlOff(fuz_ctx *c_) : c(c_), saved(c_->pdl) { c->pdl = 0; } ~PdlOff() { c->pdl = saved; } } pdl_off(ctx);
    const int64_t n_reads = in->n_reads, bestn = in->bestn > 0 ? in->bestn : 1;
    RRScratch R;
    FuzLayout L;
    size_t o_cnt = L.add(4 * (size_t)(n_reads + 2)), o_off = L.add(4 * (size_t)(n_reads + 2)), o_cur = L.add(4 * (size_t)(n_reads + 1));
    size_t o_grp = L.add(4 * (size_t)(in->n_ovl + 1)), o_al = L.add(4 * (size_t)(n_reads * bestn + 1));
    size_t o_aq = L.add(4 * (size_t)(n_reads * bestn + 1)), o_vc = L.add(4 * (size_t)(n_reads + 2));
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    R.t_cnt = fuz_at<int32_t>(ctx, o_cnt); R.t_off = fuz_at<int32_t>(ctx, o_off); R.t_cur = fuz_at<int32_t>(ctx, o_cur);
    R.grp = fuz_at<int32_t>(ctx, o_grp); R.a_len = fuz_at<int32_t>(ctx, o_al); R.a_q = fuz_at<int32_t>(ctx, o_aq);
    R.vt_cnt = fuz_at<int32_t>(ctx, o_vc);
    fuz_launch(ctx, k_rr_init, FUZ_GRID_BLOCKS, 256, 0, st, ctx->d_status, R, (int)n_reads);
    FUZ_LAUNCH_CHECK(ctx, "k_rr_init");
    fuz_launch(ctx, k_rr_filter, FUZ_GRID_BLOCKS, 256, 0, st, *in, *out, R, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_rr_filter");
    if (ctx->rr_filter_only) return FUZ_OK;      // map step of the multi-GPU run: d_keep (and reserved[3]) only
    if ((rc = fuz_scan_i32(ctx, R.t_cnt, R.t_off, n_reads, nullptr, FUZ_FIN_NONE, 0))) return rc;
    fuz_launch(ctx, k_rr_fill, ctx->grid_rr ? 148 * 8 : FUZ_GRID_BLOCKS, 256, 0, st, *in, *out, R, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_rr_fill");
    fuz_launch(ctx, k_rr_replay, ctx->grid_rr ? 148 * ctx->grid_rr : FUZ_GRID_BLOCKS, 256, 0, st, *in, *out, R, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_rr_replay");
    fuz_launch(ctx, k_rr_vote, ctx->grid_rr ? 148 * 16 : FUZ_GRID_BLOCKS, 128, 0, st, *in, *out, R, 0, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_rr_vote(count)");
    if ((rc = fuz_scan_i32(ctx, R.vt_cnt, out->d_vt_off, n_reads, nullptr, FUZ_FIN_NONE, 0))) return rc;
    fuz_launch(ctx, k_rr_votes_total, 1, 32, 0, st, *out, (int)n_reads, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_rr_votes_total");
    fuz_launch(ctx, k_rr_vote, ctx->grid_rr ? 148 * 16 : FUZ_GRID_BLOCKS, 128, 0, st, *in, *out, R, 1, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_rr_vote(fill)");
    return FUZ_OK;
}
