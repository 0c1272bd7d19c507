// Overlap filter with phase (reference falcon_unzip/ovlp_filter_with_phase.py:49-290; SURVEY.md
// section 8f-3) for ALL LAS files in one call.  Input: the parsed columns of `LA4Falcon -mo`
// (fuz_host_parse_la4falcon_mo) and the rid -> (ctg, block, phase) table with the strings
// interned to ints; output: the read sets of stage 1 / 2 and the selected lines of stage 3 in
// the reference's output order.
//
//   phase filter (:64-73, identical in the three stages)  k_of_pass + scan + k_of_compact
//   groups = runs of equal q among the passing lines of a file (:75,:228)  k_of_heads + scan
//   stage 1 (:75-142)  one warp per group counts its 5' / 3' overlaps -> ignore set
//   stage 2 (:145-186) one thread per passing line -> contained set
//   stage 3 (:188-290) one warp per group: candidates of a side into shared memory, rank of
//                      every candidate under the reference's sort key (-inphase, -len,
//                      t_l - (t_e - t_s), then the token list: t id), cut at the first rank >= bestn
//                      with a range > 1000; scan of the counts; k_of_emit writes the line indices.
// Candidates whose key ties completely (same pair, same length, same range) are ordered by line
// and the group is flagged: the host re-sorts such a group with the reference's full string
// comparison (identical lines need nothing).
#include "fuz_internal.cuh"

namespace {

#define OF_CAP 512                 // candidates of one side of one group held in shared memory

struct OvlpScratch {
    int32_t *flag, *pos;           // [n + 1] phase-filter flag and its exclusive scan
    uint8_t *bits;                 // [n] bit0 pass, bit1 idt/len ok, bit2 q_s == 0, bit3 q_e == q_l
    int32_t *P;                    // [n_pass] passing lines
    int32_t *head, *gscan;         // [n_pass + 1]
    int32_t *gstart;               // [n_groups + 1] first passing line of every group
    int32_t *rank;                 // [n_pass] output rank inside (group, side) or -1
    int32_t *ocnt, *ooff;          // [2 n_groups + 1] selected lines per (group, side)
    int64_t *n_pass, *n_groups;    // device counters (from the scans)
};

__device__ __forceinline__ bool of_in_map(const fuz_ovlp_input &in, int r) {
    return r >= 0 && r < in.n_reads && in.d_in_map[r];
}

__global__ void __launch_bounds__(256) k_of_pass(fuz_ovlp_input in, OvlpScratch S) {
    fuz_pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < in.n_ovl; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = in.d_q[i], t = in.d_t[i];
        bool pass = of_in_map(in, q) && of_in_map(in, t);
        if (pass) {
            pass = in.d_ph_ctg[q] == in.d_ph_ctg[t] &&
                   !(in.d_ph_block[q] == in.d_ph_block[t] && in.d_ph_phase[q] != in.d_ph_phase[t]);
        }
        const uint8_t f = in.d_flags[i];
        const bool ok2 = (f & 1) && in.d_ql[i] >= in.min_len && in.d_tl[i] >= in.min_len;
        S.flag[i] = pass ? 1 : 0;
        S.bits[i] = (uint8_t)((pass ? 1 : 0) | (ok2 ? 2 : 0) | (in.d_qs[i] == 0 ? 4 : 0) | (in.d_qe[i] == in.d_ql[i] ? 8 : 0));
    }
}

__global__ void __launch_bounds__(256) k_of_compact(int64_t n, OvlpScratch S) {
    fuz_pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (S.flag[i]) S.P[S.pos[i]] = (int32_t)i;
    if (blockIdx.x == 0 && threadIdx.x == 0) *S.n_pass = S.pos[n];
}

__global__ void __launch_bounds__(256) k_of_heads(fuz_ovlp_input in, OvlpScratch S) {
    fuz_pdl_enter();
    const int64_t np = *S.n_pass;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < np; j += (int64_t)gridDim.x * blockDim.x) {
        bool head = j == 0;
        if (!head) {
            const int a = S.P[j], b = S.P[j - 1];
            head = in.d_file[a] != in.d_file[b] || in.d_q[a] != in.d_q[b];
        }
        S.head[j] = head ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) k_of_gstart(OvlpScratch S, fuz_ovlp_outputs out, fuz_status *st) {
    fuz_pdl_enter();
    const int64_t np = *S.n_pass;
    const int64_t ng = S.gscan[np];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *S.n_groups = ng;
        st->reserved[0] = ng;
        if (ng > out.cap_groups) fuz_raise(st, FUZ_E_CAPACITY, 10);
        else S.gstart[ng] = (int32_t)np;
    }
    if (ng > out.cap_groups) return;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < np; j += (int64_t)gridDim.x * blockDim.x)
        if (S.head[j]) S.gstart[S.gscan[j]] = (int32_t)j;
}

// verdict of ovlp_filter_with_phase.py:80-87 for one run of a q
__device__ __forceinline__ bool of_ignored(int left, int right, const fuz_ovlp_input &in) {
    if (abs(left - right) > in.max_diff) return true;
    if (left > in.max_ovlp || right > in.max_ovlp) return true;
    return left < in.min_ovlp || right < in.min_ovlp;
}

__global__ void __launch_bounds__(256) k_of_stage1(fuz_ovlp_input in, OvlpScratch S, fuz_ovlp_outputs out, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int64_t ng = *S.n_groups;
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ng; g += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int j0 = S.gstart[g], j1 = S.gstart[g + 1];
        int left = 0, right = 0;
        for (int j = j0 + lane; j < j1; j += 32) {
            const uint8_t b = S.bits[S.P[j]];
            if (b & 2) { left += (b >> 2) & 1; right += (b >> 3) & 1; }
        }
        left = fuz_warp_sum(left); right = fuz_warp_sum(right);
        if (lane == 0) {
            const int q = in.d_q[S.P[j0]];
            const bool ig = of_ignored(left, right, in);
            out.d_grp_q[g] = q;
            out.d_grp_line[g] = S.P[j0];
            out.d_grp_ignore[g] = ig ? 1 : 0;
            if (ig && !in.d_ignore_in) out.d_ignore[q] = 1;
        }
    }
}

__global__ void __launch_bounds__(256) k_of_stage2(fuz_ovlp_input in, OvlpScratch S, fuz_ovlp_outputs out, const uint8_t *__restrict__ ignore,
                                                   const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t np = *S.n_pass;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < np; j += (int64_t)gridDim.x * blockDim.x) {
        const int i = S.P[j];
        if (!(S.bits[i] & 2)) continue;
        const int q = in.d_q[i], t = in.d_t[i];
        if (ignore[q] || ignore[t]) continue;
        const int tag = (in.d_flags[i] >> 1) & 3;
        if (tag == 3) out.d_contained[q] = 1;
        if (tag == 2) out.d_contained[t] = 1;
    }
}

__global__ void __launch_bounds__(128) k_of_stage3(fuz_ovlp_input in, OvlpScratch S, fuz_ovlp_outputs out, const uint8_t *__restrict__ ignore,
                                                   const uint8_t *__restrict__ contained, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    __shared__ unsigned long long s_hi[4][OF_CAP], s_lo[4][OF_CAP];
    __shared__ int s_j[4][OF_CAP];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long *hi = s_hi[w], *lo = s_lo[w];
    int *sj = s_j[w];
    const int64_t ng = *S.n_groups;
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ng; g += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int j0 = S.gstart[g], j1 = S.gstart[g + 1];
        const int q = in.d_q[S.P[j0]];
        const bool q_out = ignore[q] || contained[q];                    // :246-253, the q side
        bool tie = false;
        for (int side = 0; side < 2; side++) {
            // ---- candidates of this side, in line order
            int n = 0;
            for (int jb = j0; jb < j1 && !q_out; jb += 32) {
                const int j = jb + lane;
                bool cand = false;
                int i = 0, t = 0;
                if (j < j1) {
                    i = S.P[j];
                    const uint8_t b = S.bits[i];
                    t = in.d_t[i];
                    const bool s5 = (b & 4) != 0, s3 = !s5 && (b & 8);       // `if q_s == 0 ... elif q_e == q_l` (:265,:270)
                    cand = (b & 2) && (side == 0 ? s5 : s3) && !ignore[t] && !contained[t];
                }
                const uint32_t m = __ballot_sync(0xffffffffu, cand);
                if (cand && out.d_cand) out.d_cand[i] = (uint8_t)(side + 1);
                if (cand) {
                    const int k = n + __popc(m & ((1u << lane) - 1u));
                    if (k < OF_CAP) {
                        const bool inphase = in.d_ph_ctg[q] == in.d_ph_ctg[t] && in.d_ph_block[q] == in.d_ph_block[t] &&
                                             in.d_ph_phase[q] == in.d_ph_phase[t];
                        const int32_t neg_len = -in.d_len[i], range = in.d_tl[i] - (in.d_te[i] - in.d_ts[i]);
                        hi[k] = ((unsigned long long)(inphase ? 0u : 1u) << 32) | (uint32_t)(neg_len ^ 0x80000000);
                        lo[k] = ((unsigned long long)(uint32_t)(range ^ 0x80000000) << 32) | (uint32_t)t;
                        sj[k] = j;
                    }
                }
                n += __popc(m);
            }
            if (n > OF_CAP) {
                if (lane == 0) fuz_raise(st, FUZ_E_CAPACITY, 9);
                n = 0;
            }
            __syncwarp();
            // ---- rank of every candidate; cut = first rank >= bestn whose range is > 1000 (:237-240)
            int cut = n - 1;
            int my_rank[OF_CAP / 32];
#pragma unroll
            for (int r = 0; r < OF_CAP / 32; r++) {
                my_rank[r] = -1;
                const int k = r * 32 + lane;
                if (k < n) {
                    const unsigned long long h = hi[k], l = lo[k];
                    int rk = 0;
                    for (int o = 0; o < n; o++) {
                        const unsigned long long h2 = hi[o], l2 = lo[o];
                        const bool less = h2 < h || (h2 == h && (l2 < l || (l2 == l && o < k)));
                        if (h2 == h && l2 == l && o != k) tie = true;
                        rk += less ? 1 : 0;
                    }
                    my_rank[r] = rk;
                    const int32_t range = (int32_t)((uint32_t)(l >> 32) ^ 0x80000000u);
                    if (rk >= in.bestn && range > 1000) cut = min(cut, rk);
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cut = min(cut, __shfl_xor_sync(0xffffffffu, cut, d));
#pragma unroll
            for (int r = 0; r < OF_CAP / 32; r++) {
                const int k = r * 32 + lane;
                if (k < n) S.rank[sj[k]] = my_rank[r] <= cut ? my_rank[r] : -1;
            }
            if (lane == 0) S.ocnt[2 * g + side] = n ? cut + 1 : 0;
            __syncwarp();
        }
        tie = __any_sync(0xffffffffu, tie);
        if (lane == 0) out.d_grp_tie[g] = tie ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) k_of_emit(fuz_ovlp_input in, OvlpScratch S, fuz_ovlp_outputs out, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t ng = *S.n_groups;
    const int64_t total = S.ooff[2 * ng];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        st->reserved[2] = total;
        if (total > out.cap_out) fuz_raise(st, FUZ_E_CAPACITY, 11); else st->reserved[1] = total;
    }
    if (total > out.cap_out) return;
    // one warp per group keeps the loads of a group together
    const int lane = threadIdx.x & 31;
    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ng; g += ((int64_t)gridDim.x * blockDim.x) >> 5) {
        const int j0 = S.gstart[g], j1 = S.gstart[g + 1];
        if (lane == 0) out.d_grp_off[g] = S.ooff[2 * g];
        if (lane == 0 && g == ng - 1) out.d_grp_off[ng] = total;
        for (int j = j0 + lane; j < j1; j += 32) {
            const int rk = S.rank[j];
            if (rk < 0) continue;
            const int i = S.P[j];
            const int side = (S.bits[i] & 4) ? 0 : 1;
            out.d_out_line[S.ooff[2 * g + side] + rk] = i;
        }
    }
}

__global__ void k_of_fill_i32(int32_t *p, int64_t n, int32_t v) {
    fuz_pdl_enter();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

extern "C" int fuz_ovlp_filter(fuz_ctx *ctx, const fuz_ovlp_input *in, fuz_ovlp_outputs *out) {
    if (!ctx || !in || !out) return FUZ_E_ARG;
    if (in->n_ovl < 0 || in->n_ovl > 0x7ffffff0LL || in->n_reads < 0) return fuz_fail(ctx, FUZ_E_ARG, "fuz_ovlp_filter: bad sizes");
    if (in->stage < 1 || in->stage > 3) return fuz_fail(ctx, FUZ_E_ARG, "fuz_ovlp_filter: stage must be 1, 2 or 3");
    if (!out->d_ignore || !out->d_contained || !out->d_grp_q || !out->d_grp_line || !out->d_grp_ignore || !out->d_grp_tie || !out->d_grp_off ||
        (!out->d_out_line && out->cap_out > 0))
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_ovlp_filter: missing output buffer");
    cudaStream_t st = ctx->stream;
    const int64_t n = in->n_ovl;
    FuzLayout L;
    const size_t o_flag = L.add(4 * (size_t)(n + 1)), o_pos = L.add(4 * (size_t)(n + 2)), o_bits = L.add((size_t)n + 1);
    const size_t o_P = L.add(4 * (size_t)(n + 1)), o_head = L.add(4 * (size_t)(n + 1)), o_gscan = L.add(4 * (size_t)(n + 2));
    const size_t o_gstart = L.add(4 * (size_t)(out->cap_groups + 2)), o_rank = L.add(4 * (size_t)(n + 1));
    const size_t o_ocnt = L.add(4 * (size_t)(2 * out->cap_groups + 2)), o_ooff = L.add(4 * (size_t)(2 * out->cap_groups + 3));
    const size_t o_cnt = L.add(16);
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    OvlpScratch S;
    S.flag = fuz_at<int32_t>(ctx, o_flag); S.pos = fuz_at<int32_t>(ctx, o_pos); S.bits = fuz_at<uint8_t>(ctx, o_bits);
    S.P = fuz_at<int32_t>(ctx, o_P); S.head = fuz_at<int32_t>(ctx, o_head); S.gscan = fuz_at<int32_t>(ctx, o_gscan);
    S.gstart = fuz_at<int32_t>(ctx, o_gstart); S.rank = fuz_at<int32_t>(ctx, o_rank);
    S.ocnt = fuz_at<int32_t>(ctx, o_ocnt); S.ooff = fuz_at<int32_t>(ctx, o_ooff);
    S.n_pass = fuz_at<int64_t>(ctx, o_cnt); S.n_groups = S.n_pass + 1;
    FUZ_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(fuz_status), st));
    FUZ_CUDA(ctx, cudaMemsetAsync(S.n_pass, 0, 16, st));
    FUZ_CUDA(ctx, cudaMemsetAsync(S.head, 0, 4 * (size_t)(n + 1), st));
    if (!in->d_ignore_in) FUZ_CUDA(ctx, cudaMemsetAsync(out->d_ignore, 0, (size_t)in->n_reads, st));
    if (!in->d_contained_in) FUZ_CUDA(ctx, cudaMemsetAsync(out->d_contained, 0, (size_t)in->n_reads, st));
    fuz_launch(ctx, k_of_pass, FUZ_GRID_BLOCKS, 256, 0, st, *in, S);
    FUZ_LAUNCH_CHECK(ctx, "k_of_pass");
    if ((rc = fuz_scan_i32_wide(ctx, S.flag, S.pos, n))) return rc;
    fuz_launch(ctx, k_of_compact, FUZ_GRID_BLOCKS, 256, 0, st, n, S);
    FUZ_LAUNCH_CHECK(ctx, "k_of_compact");
    fuz_launch(ctx, k_of_heads, FUZ_GRID_BLOCKS, 256, 0, st, *in, S);
    FUZ_LAUNCH_CHECK(ctx, "k_of_heads");
    // heads beyond n_pass stay 0 (memset), so scanning all n entries gives gscan[n_pass] = groups
    if ((rc = fuz_scan_i32_wide(ctx, S.head, S.gscan, n))) return rc;
    fuz_launch(ctx, k_of_gstart, FUZ_GRID_BLOCKS, 256, 0, st, S, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_of_gstart");
    fuz_launch(ctx, k_of_stage1, FUZ_GRID_BLOCKS, 256, 0, st, *in, S, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_of_stage1");
    if (in->stage == 1) return FUZ_OK;
    const uint8_t *ignore = in->d_ignore_in ? in->d_ignore_in : out->d_ignore;
    if (!in->d_contained_in) {
        fuz_launch(ctx, k_of_stage2, FUZ_GRID_BLOCKS, 256, 0, st, *in, S, *out, ignore, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_of_stage2");
    }
    if (in->stage == 2) return FUZ_OK;
    const uint8_t *contained = in->d_contained_in ? in->d_contained_in : out->d_contained;
    fuz_launch(ctx, k_of_fill_i32, FUZ_GRID_BLOCKS, 256, 0, st, S.rank, n, -1);
    FUZ_LAUNCH_CHECK(ctx, "k_of_fill_i32");
    FUZ_CUDA(ctx, cudaMemsetAsync(S.ocnt, 0, 4 * (size_t)(2 * out->cap_groups + 2), st));
    if (out->d_cand && n) FUZ_CUDA(ctx, cudaMemsetAsync(out->d_cand, 0, (size_t)n, st));
    fuz_launch(ctx, k_of_stage3, 148 * 8, 128, 0, st, *in, S, *out, ignore, contained, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_of_stage3");
    // counts beyond 2 * n_groups are 0 (memset): scanning 2 * cap_groups entries leaves ooff[2 n_groups] = total
    if ((rc = fuz_scan_i32_wide(ctx, S.ocnt, S.ooff, 2 * out->cap_groups))) return rc;
    fuz_launch(ctx, k_of_emit, FUZ_GRID_BLOCKS, 256, 0, st, *in, S, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_of_emit");
    return FUZ_OK;
}
