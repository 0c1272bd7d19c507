// Stages 2-4 of the phasing path on device arrays:
//   association table   reference falcon_unzip/phasing.py:137-206
//   phased blocks       reference falcon_unzip/phasing.py:208-421
//   phased reads        reference falcon_unzip/phasing.py:423-480
// All row counts live in the device status block; no host synchronisation in between.
#include "fuz_internal.cuh"

namespace {

// ------------------------------------------------------------------ shared helpers
__global__ void k_set_counts(fuz_status *st, int64_t n_sites, int64_t n_vmap, int64_t n_atable, int reset_error) {
    fuz_pdl_enter();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (reset_error) { st->error = 0; st->error_index = 0; }
        if (n_sites >= 0) st->n_sites = n_sites;
        if (n_vmap >= 0) st->n_vmap = n_vmap;
        if (n_atable >= 0) st->n_atable = n_atable;
    }
}

// row range of every site inside the (site-grouped) vmap rows
__global__ void k_site_rowoff(const int32_t *__restrict__ vm_site, int32_t *__restrict__ row_off, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int n_sites = (int)st->n_sites, n_vmap = (int)st->n_vmap;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s <= n_sites; s += gridDim.x * blockDim.x)
        row_off[s] = s == n_sites ? n_vmap : fuz_lower_bound(vm_site, 0, n_vmap, s);
}

// One warp per site.  dup[i] = an earlier row of the same (site, allele) carries the same
// q_id (the reference builds set(qids): phasing.py:189, :448-449).
__global__ void __launch_bounds__(256) k_dup_flags(const int32_t *__restrict__ row_off, const uint8_t *__restrict__ vm_base,
                                                   const int32_t *__restrict__ vm_qid, uint8_t *__restrict__ dup,
                                                   const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_sites = (int)st->n_sites;
    for (int s = warp_g; s < n_sites; s += n_warps) {
        const int off = row_off[s], n = row_off[s + 1] - off;
        for (int i = lane; i < n; i += 32) {
            const int q = vm_qid[off + i];
            const uint8_t b = vm_base[off + i];
            uint8_t d = 0;
            for (int j = 0; j < i; j++)
                if (vm_qid[off + j] == q && vm_base[off + j] == b) { d = 1; break; }
            dup[off + i] = d;
        }
    }
}

// ================================================================== association table
struct AssocScratch {
    int32_t *row_off, *uq, *uq_n, *na0, *qmin, *qmax, *cand_cnt, *cand_off, *at_cnt, *at_off;
    const int32_t *site_ctg, *site_pos;
    int4 *pair_ct;
    uint8_t *dup;
    uint32_t *qmask;                 // per site 2 x FUZ_QM_WORDS words: the q_id set of each allele as bits over [qmin & ~31, +544)
    uint8_t *has_mask;               // 1: qmask[s] is valid (q_id range of the site <= FUZ_UQ_RANGE)
    int64_t max_pairs;
};
#define FUZ_QM_WORDS 17              // 512 q_ids from an anchor rounded down to 32

// sorted unique q_id lists per (site, allele): uq[row_off[s] ..) for allele al0 and
// uq[row_off[s] + na0[s] ..) for allele al1, lengths uq_n[2s], uq_n[2s+1]; duplicate flags
// of the rows (an earlier row of the same (site, allele) carries the same q_id: the
// reference builds set(qids), phasing.py:189, :448-449).  One warp per site.  The q_ids of a
// site span a small range (q_ids are handed out in coordinate order), so the common path
// is a direct-address table in shared memory: first row per (q_id, allele) by atomicMin,
// then ballot/popc compaction in q_id order.  Wider sites use all-pairs ranking.
#define FUZ_UQ_RANGE 512
__global__ void __launch_bounds__(256) k_uniq_lists(const uint8_t *__restrict__ site_al, const uint8_t *__restrict__ vm_base,
                                                    const int32_t *__restrict__ vm_qid, AssocScratch A, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    __shared__ int s_first[8][2][FUZ_UQ_RANGE];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_sites = (int)st->n_sites;
    const uint32_t lt = (1u << lane) - 1u;
    for (int s = warp_g; s < n_sites; s += n_warps) {
        const int off = A.row_off[s], n = A.row_off[s + 1] - off;
        const uint8_t al0 = site_al[2 * s], al1 = site_al[2 * s + 1];
        int c0 = 0, mn = 0x7fffffff, mx = -0x7fffffff - 1;
        bool bad = al0 == al1 || al0 > 3 || al1 > 3;
        for (int i = lane; i < n; i += 32) {
            const uint8_t b = vm_base[off + i];
            const int q = vm_qid[off + i];
            if (b != al0 && b != al1) bad = true;
            c0 += b == al0;
            mn = min(mn, q); mx = max(mx, q);
        }
        c0 = __reduce_add_sync(0xffffffffu, c0);
        mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx);
        bad = __any_sync(0xffffffffu, bad);
        if (bad || c0 == 0 || c0 == n) {          // a site must carry exactly two alleles
            if (lane == 0) { fuz_raise(st, FUZ_E_FORMAT, s); A.cand_cnt[s] = 0; A.has_mask[s] = 0; }
            continue;
        }
        int u0 = 0, u1 = 0;
        const long long range = (long long)mx - mn + 1;
        if (range <= FUZ_UQ_RANGE) {
            int *f0 = s_first[wib][0], *f1 = s_first[wib][1];
            __syncwarp();
            for (int x = lane; x < range; x += 32) { f0[x] = 0x7fffffff; f1[x] = 0x7fffffff; }
            __syncwarp();
            for (int i = lane; i < n; i += 32)
                atomicMin((vm_base[off + i] == al0 ? f0 : f1) + (vm_qid[off + i] - mn), i);
            __syncwarp();
            for (int i = lane; i < n; i += 32)
                A.dup[off + i] = (vm_base[off + i] == al0 ? f0 : f1)[vm_qid[off + i] - mn] != i;
            // in q_id order, 32 q_ids per round from the anchor qmin & ~31: the ballots are the words of the site's q_id masks
            // (k_pair_count intersects two sites by AND + POPC of such words: the anchors differ by whole words)
            const int sh = mn & 31;
            uint32_t w0 = 0, w1 = 0;                      // lane w keeps word w
#pragma unroll 1
            for (int w = 0; 32 * w < sh + (int)range; w++) {                   // (the words beyond stay 0)
                const int x = 32 * w + lane - sh;
                const bool in = x >= 0 && x < range;
                const bool p0 = in && f0[x] != 0x7fffffff, p1 = in && f1[x] != 0x7fffffff;
                const uint32_t m0 = __ballot_sync(0xffffffffu, p0), m1 = __ballot_sync(0xffffffffu, p1);
                if (p0) A.uq[off + u0 + __popc(m0 & lt)] = mn + x;
                if (p1) A.uq[off + c0 + u1 + __popc(m1 & lt)] = mn + x;
                u0 += __popc(m0); u1 += __popc(m1);
                if (lane == w) { w0 = m0; w1 = m1; }
            }
            if (lane < FUZ_QM_WORDS) {
                A.qmask[(size_t)s * 2 * FUZ_QM_WORDS + lane] = w0;
                A.qmask[(size_t)s * 2 * FUZ_QM_WORDS + FUZ_QM_WORDS + lane] = w1;
            }
        } else {
            // all-pairs: duplicate flag, then the slot of every first occurrence in its sorted unique list
            for (int i = lane; i < n; i += 32) {
                const int q = vm_qid[off + i];
                const uint8_t b = vm_base[off + i];
                bool dup = false;
                for (int j = 0; j < i; j++)
                    if (vm_qid[off + j] == q && vm_base[off + j] == b) { dup = true; break; }
                A.dup[off + i] = dup ? 1 : 0;
            }
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                if (A.dup[off + i]) continue;
                const int q = vm_qid[off + i];
                const uint8_t b = vm_base[off + i];
                int rank = 0;
                for (int j = 0; j < n; j++) rank += (!A.dup[off + j] && vm_base[off + j] == b && vm_qid[off + j] < q);
                A.uq[off + (b == al0 ? 0 : c0) + rank] = q;
                if (b == al0) u0++; else u1++;
            }
            u0 = __reduce_add_sync(0xffffffffu, u0); u1 = __reduce_add_sync(0xffffffffu, u1);
        }
        if (lane == 0) {
            A.na0[s] = c0; A.uq_n[2 * s] = u0; A.uq_n[2 * s + 1] = u1; A.qmin[s] = mn; A.qmax[s] = mx;
            A.has_mask[s] = range <= FUZ_UQ_RANGE ? 1 : 0;
            // number of later sites of the same contig within 65536 bp (phasing.py:166-170)
            const int c = A.site_ctg[s];
            const long long lim = (long long)A.site_pos[s] + (1 << 16);
            int lo = s + 1, hi = n_sites;                // first index with (ctg,pos) > (c, lim)
            while (lo < hi) {
                int m = (lo + hi) >> 1;
                bool le = A.site_ctg[m] < c || (A.site_ctg[m] == c && (long long)A.site_pos[m] <= lim);
                if (le) lo = m + 1; else hi = m;
            }
            A.cand_cnt[s] = lo - s - 1;
        }
    }
}

__device__ __forceinline__ int sorted_intersect(const int32_t *__restrict__ a, int na, const int32_t *__restrict__ b, int nb) {
    int i = 0, j = 0, s = 0;
    while (i < na && j < nb) {
        int x = a[i], y = b[j];
        s += x == y; i += x <= y; j += y <= x;
    }
    return s;
}

// One warp per left site, one lane per candidate partner: 2x2 set-intersection sizes
// (phasing.py:187-191), kept for the fill pass; emitted rows are capped at 501 per left
// site AFTER the total >= 6 filter (phasing.py:192-206).  The two q_id sets of every site are bit masks
// over its (small) q_id range (k_uniq_lists); the left site's sit in shared memory and a pair is 4 x (AND, POPC)
// per word both sites cover.  A partner without masks is tested q_id by q_id, a left site without merges sorted lists.
__global__ void __launch_bounds__(256) k_pair_count(AssocScratch A, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    __shared__ uint32_t s_mask[8][2][FUZ_QM_WORDS + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_sites = (int)st->n_sites;
    uint32_t *m0 = s_mask[wib][0], *m1 = s_mask[wib][1];
    for (int i1 = warp_g; i1 < n_sites; i1 += n_warps) {
        const int nc = A.cand_cnt[i1];
        const int64_t base = A.cand_off[i1];
        const int o1 = A.row_off[i1], n10 = A.uq_n[2 * i1], n11 = A.uq_n[2 * i1 + 1];
        const int32_t *a0 = A.uq + o1, *a1 = A.uq + o1 + A.na0[i1];
        const int mn1 = A.qmin[i1], mx1 = A.qmax[i1];
        const int anc1 = mn1 & ~31;
        const bool masked = A.has_mask[i1] != 0;
        __syncwarp();
        if (masked && nc > 0) {                             // the q_id masks of the left site: k_uniq_lists built them
            if (lane < FUZ_QM_WORDS) {
                m0[lane] = A.qmask[(size_t)i1 * 2 * FUZ_QM_WORDS + lane];
                m1[lane] = A.qmask[(size_t)i1 * 2 * FUZ_QM_WORDS + FUZ_QM_WORDS + lane];
            }
            __syncwarp();
        }
        int emitted = 0;
        for (int k0 = 0; k0 < nc && emitted <= 500; k0 += 32) {
            const int k = k0 + lane;
            const bool valid = k < nc;
            int4 ct = make_int4(0, 0, 0, 0);
            if (valid) {
                const int i2 = i1 + 1 + k;
                const int mn2 = A.qmin[i2];
                if (mn2 <= mx1 && mn1 <= A.qmax[i2]) {             // q_id ranges overlap
                    if (masked && A.has_mask[i2]) {
                        // both sites as masks: the anchors are multiples of 32, so the words line up; only the words both cover
                        const uint32_t *b0 = A.qmask + (size_t)i2 * 2 * FUZ_QM_WORDS, *b1 = b0 + FUZ_QM_WORDS;
                        const int dw = ((mn2 & ~31) - anc1) >> 5;            // word of the left mask under word 0 of the right one
                        const int w_lo = max(0, -dw), w_hi = min(FUZ_QM_WORDS, FUZ_QM_WORDS - dw);
                        for (int w = w_lo; w < w_hi; w++) {
                            const uint32_t x0 = b0[w], x1 = b1[w], l0 = m0[w + dw], l1 = m1[w + dw];
                            ct.x += __popc(l0 & x0); ct.y += __popc(l0 & x1); ct.z += __popc(l1 & x0); ct.w += __popc(l1 & x1);
                        }
                    } else {
                        const int o2 = A.row_off[i2], n20 = A.uq_n[2 * i2], n21 = A.uq_n[2 * i2 + 1];
                        const int32_t *b0 = A.uq + o2, *b1 = A.uq + o2 + A.na0[i2];
                        if (masked) {
                            for (int e = 0; e < n20; e++) {
                                const unsigned x = (unsigned)(b0[e] - anc1);
                                if (x < 32u * FUZ_QM_WORDS) { ct.x += (m0[x >> 5] >> (x & 31)) & 1u; ct.z += (m1[x >> 5] >> (x & 31)) & 1u; }
                            }
                            for (int e = 0; e < n21; e++) {
                                const unsigned x = (unsigned)(b1[e] - anc1);
                                if (x < 32u * FUZ_QM_WORDS) { ct.y += (m0[x >> 5] >> (x & 31)) & 1u; ct.w += (m1[x >> 5] >> (x & 31)) & 1u; }
                            }
                        } else {
                            ct.x = sorted_intersect(a0, n10, b0, n20);
                            ct.y = sorted_intersect(a0, n10, b1, n21);
                            ct.z = sorted_intersect(a1, n11, b0, n20);
                            ct.w = sorted_intersect(a1, n11, b1, n21);
                        }
                    }
                }
                A.pair_ct[base + k] = ct;
            }
            emitted += __popc(__ballot_sync(0xffffffffu, valid && ct.x + ct.y + ct.z + ct.w >= 6));
        }
        if (lane == 0) A.at_cnt[i1] = min(emitted, 501);
    }
}

__global__ void __launch_bounds__(256) k_pair_fill(AssocScratch A, fuz_outputs O, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_sites = (int)st->n_sites;
    const uint32_t lt = (1u << lane) - 1u;
    for (int i1 = warp_g; i1 < n_sites; i1 += n_warps) {
        const int nc = A.cand_cnt[i1];
        const int64_t base = A.cand_off[i1];
        const int64_t out0 = A.at_off[i1];
        int run = 0;
        for (int k0 = 0; k0 < nc && run <= 500; k0 += 32) {
            const int k = k0 + lane;
            const bool valid = k < nc;
            int4 ct = valid ? A.pair_ct[base + k] : make_int4(0, 0, 0, 0);
            const bool pass = valid && ct.x + ct.y + ct.z + ct.w >= 6;
            const uint32_t m = __ballot_sync(0xffffffffu, pass);
            const int rank = run + __popc(m & lt);
            if (pass && rank <= 500) {
                const int64_t o = out0 + rank;
                O.d_at_s1[o] = i1; O.d_at_s2[o] = i1 + 1 + k;
                reinterpret_cast<int4 *>(O.d_at_ct)[o] = ct;
            }
            run += __popc(m);
        }
    }
}

// ================================================================== phased blocks
struct BlockScratch {
    int32_t *ctg_site_off, *left_cnt, *left_off, *left_cur, *lq, *ld, *right_off, *min_k;
    int32_t *bidx, *bsize, *bnew;
    uint32_t *fp;   // forest pointer: parent * 2 + parity bit
    int32_t *g_pk;  // [edges] packed left edges (d << 6 | back) for contigs whose sweep runs from global memory
    uint8_t *g_slow;    // [sites] their per-site flags (bit 0: generic sweep step, bit 1: site is in positions)
    long long *dbg; // optional: per-contig phase timestamps (16 per contig), diagnostics only
    int staging;    // ctx option "phase_staging"
    int sweep_passes;   // ctx option "sweep_passes": parallel fixed-point passes before the sequential sweep (0 = none)
};

#define FUZ_PHASE_THREADS 1024
#define FUZ_PHASE_SMEM (160 * 1024)      // dynamic shared memory of k_ctg_phase

__device__ __forceinline__ int row_d(const int32_t *__restrict__ at_ct, int row) {       // cis - trans
    int4 c = reinterpret_cast<const int4 *>(at_ct)[row];
    return (c.x + c.w) - (c.y + c.z);
}

// zero the counters; first site of every contig; row range of every site as left site
__global__ void k_blk_init(BlockScratch B, fuz_outputs O, int n_ctg, int right_off_valid, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int n_sites = (int)st->n_sites, n_at = (int)st->n_atable;
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s <= n_sites; s += gridDim.x * blockDim.x) {
        B.left_cnt[s] = 0;
        B.bsize[s] = 0;
        if (s < n_sites) B.left_cur[s] = 0;
        if (!right_off_valid) B.right_off[s] = s == n_sites ? n_at : fuz_lower_bound(O.d_at_s1, 0, n_at, s);
        if (s <= n_ctg) B.ctg_site_off[s] = s == n_ctg ? n_sites : fuz_lower_bound(O.d_site_ctg, 0, n_sites, s);
    }
    // n_ctg may exceed n_sites + 1
    for (int c = n_sites + 1 + blockIdx.x * blockDim.x + threadIdx.x; c <= n_ctg; c += gridDim.x * blockDim.x)
        B.ctg_site_off[c] = c == n_ctg ? n_sites : fuz_lower_bound(O.d_site_ctg, 0, n_sites, c);
}

// accepted rows |cis - trans| >= 6 (phasing.py:245) -> in-degree of the right site
__global__ void k_edge_count(BlockScratch B, fuz_outputs O, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int n_at = (int)st->n_atable;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_at; r += gridDim.x * blockDim.x) {
        int d = row_d(O.d_at_ct, r);
        int s2 = O.d_at_s2[r];
        if (abs(d) >= 6 && s2 >= 0 && s2 < (int)st->n_sites) atomicAdd(&B.left_cnt[s2], 1);
    }
}

// left adjacency in CSR form: lq = left partner, ld = cis - trans (order inside a site's list
// is arbitrary: only sums / minima over it are used)
__global__ void k_edge_fill(BlockScratch B, fuz_outputs O, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int n_at = (int)st->n_atable, n_sites = (int)st->n_sites;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_at; r += gridDim.x * blockDim.x) {
        const int s1 = O.d_at_s1[r], s2 = O.d_at_s2[r];
        if (s1 < 0 || s2 <= s1 || s2 >= n_sites || O.d_site_ctg[s1] != O.d_site_ctg[s2] ||
            (r > 0 && (O.d_at_s1[r - 1] > s1 || (O.d_at_s1[r - 1] == s1 && O.d_at_s2[r - 1] >= s2)))) {
            fuz_raise(st, FUZ_E_FORMAT, r);     // rows must be strictly ordered by (site1, site2)
            continue;
        }
        const int d = row_d(O.d_at_ct, r);
        if (abs(d) >= 6) {
            int k = B.left_off[s2] + atomicAdd(&B.left_cur[s2], 1);
            B.lq[k] = s1; B.ld[k] = d;
        }
    }
}

// CTA-wide exclusive scans over one value per thread (sum / max); s_tmp has 33 ints.
__device__ __forceinline__ int cta_excl_sum(int v, int *s_tmp, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = fuz_warp_incl_scan(v, lane);
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = lane < nw ? s_tmp[lane] : 0;
        int ti = fuz_warp_incl_scan(t, lane);
        s_tmp[lane] = ti - t;
        if (lane == 31) s_tmp[32] = ti;
    }
    __syncthreads();
    int r = s_tmp[warp] + incl - v;
    *total = s_tmp[32];
    __syncthreads();
    return r;
}
__device__ __forceinline__ int cta_excl_max(int v, int *s_tmp, int *total) {   // values >= 0, identity 0
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl = max(incl, t);
    }
    int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0;
    if (lane == 31) s_tmp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = lane < nw ? s_tmp[lane] : 0, ti = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int u = __shfl_up_sync(0xffffffffu, ti, d);
            if (lane >= d) ti = max(ti, u);
        }
        int te = __shfl_up_sync(0xffffffffu, ti, 1);
        if (lane == 0) te = 0;
        s_tmp[lane] = te;
        if (lane == 31) s_tmp[32] = ti;
    }
    __syncthreads();
    int r = max(s_tmp[warp], excl);
    *total = s_tmp[32];
    __syncthreads();
    return r;
}

// The pass-2 sweep of one contig, executed by one warp (adjacency in shared memory, or -- for contigs too large for
// that -- in global memory with the same layout).  Per site the loop-carried
// chain is: window shift -> select +-d -> REDUX -> sign -> window update.  The adjacency is
// pre-packed per edge as (d << 6 | back), back = distance to the partner in sites; list offsets
// are read four sites ahead and edges two sites ahead into registers, so no load (and no
// address computation depending on a load) sits on the chain.  Sites whose partners do not
// fit the fast path (more than 32 of them, or one more than 32 sites back) are flagged in
// s_slow and take the generic code.
__device__ __noinline__ void sweep_sites_lean(int n, int lane, uint32_t *sbits, const int *__restrict__ s_loff,
                                                 const int *__restrict__ s_lq, const int *__restrict__ s_ld, int q_base,
                                                 const int *__restrict__ s_pk, const uint8_t *__restrict__ s_slow) {
    volatile uint32_t *vb = sbits;
    uint32_t recent = 0;                                    // bit j = state of site i - 1 - j
    uint32_t word = vb[0];
    auto off_at = [&](int site) { return s_loff[min(site, n)]; };
    auto edge_at = [&](int lo, int hi) { const int k = lo + lane; return k < hi ? s_pk[k] : 0; };
    // register pipeline: offsets of sites i+2 .. i+4, packed edges of sites i, i+1
    int o2 = off_at(2), o3 = off_at(3), o4 = off_at(4);
    int pk_cur = edge_at(off_at(0), off_at(1)), pk_nxt = edge_at(off_at(1), o2);
    uint32_t slow_cur = s_slow[0] & 1u, slow_nxt = s_slow[min(1, n - 1)] & 1u;
    int d = pk_cur >> 6, back = pk_cur & 31;
    for (int i = 0; i < n; i++) {
        // ---- chain of site i (everything it needs is already in registers)
        const uint32_t own = (word >> (i & 31)) & 1u;
        const uint32_t sq = (recent >> back) & 1u;
        int s0 = (d ^ -(int)sq) + (int)sq;                  // sq ? -d : d ; lanes without a partner carry d = 0
        if (slow_cur) {                                     // uniform, rare: far or > 32 partners
            s0 = 0;
#pragma unroll 1
            for (int k = s_loff[i] + lane; k < s_loff[i + 1]; k += 32) {
                const int q2 = s_lq[k] - q_base, d2 = s_ld[k], b2 = i - 1 - q2;
                const uint32_t sq2 = b2 < 32 ? (recent >> b2) & 1u : (vb[q2 >> 5] >> (q2 & 31)) & 1u;
                s0 += sq2 ? -d2 : d2;
            }
        }
        // ---- off the chain: edges of site i + 2, offsets of site i + 5, unpack site i + 1
        const int pk_new = edge_at(o2, o3);
        const uint32_t slow_new = s_slow[min(i + 2, n - 1)] & 1u;
        o2 = o3; o3 = o4; o4 = off_at(i + 5);
        const int d_n = pk_nxt >> 6, back_n = pk_nxt & 31;
        s0 = __reduce_add_sync(0xffffffffu, s0);
        const uint32_t nw = s0 < 0 ? 1u : (s0 > 0 ? 0u : own);
        recent = (recent << 1) | nw;
        word = (word & ~(1u << (i & 31))) | (nw << (i & 31));
        if ((i & 31) == 31 || i == n - 1) {                 // publish the finished word, fetch the next
            if (lane == 0) vb[i >> 5] = word;
            __syncwarp();
            if (i + 1 < n) word = vb[(i + 1) >> 5];
        }
        pk_nxt = pk_new; slow_cur = slow_nxt; slow_nxt = slow_new; d = d_n; back = back_n;
    }
}

// One CTA per contig: pass-1 forest + pointer jumping, pass-2 sweep, pass-3 extents and
// scores, pass-4 block chaining (phasing.py:240-408; parallel forms of SURVEY.md A.3).
// Everything the passes touch (left / right adjacency, positions, forest pointers, phase
// bits) is staged into shared memory with one parallel load phase when it fits, so the
// inherently sequential sweep and the pointer chasing run at shared-memory latency.
// Contigs too large for that (> ~10^4 sites) fall back to global memory for the adjacency.
__global__ void __launch_bounds__(FUZ_PHASE_THREADS) k_ctg_phase(BlockScratch B, fuz_outputs O, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    extern __shared__ uint32_t smem[];
    __shared__ int s_tmp[33];
    const int c = blockIdx.x;
    const int cs0 = B.ctg_site_off[c], cs1 = B.ctg_site_off[c + 1];
    const int n = cs1 - cs0;
    if (n <= 0) return;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = nt >> 5;
    const int e0 = B.left_off[cs0], n_e = B.left_off[cs1] - e0;         // accepted left edges of the contig
    const int r0 = B.right_off[cs0], n_r = B.right_off[cs1] - r0;       // atable rows of the contig
    const int nbw = (n + 31) >> 5;
    if ((size_t)nbw * 4 > FUZ_PHASE_SMEM) { if (tid == 0) fuz_raise(st, FUZ_E_CAPACITY, 5); return; }
    // shared-memory layout (words).  Tier "sweep" (what the sequential sweep and the pointer
    // jumping touch): bits | loff | pk | slow | fp.  Tier "full" adds what the parallel passes
    // read: lq | ld | roff | rq | rd | pos | sc.  Large contigs that do not fit "full" still run
    // the sweep from shared memory; the parallel passes then read global memory.
    const size_t need_sweep = (size_t)nbw + (n + 1) + n_e + ((n + 3) / 4 + 1) + n;
    const size_t need = need_sweep + 2 * (size_t)n_e + (n + 1) + 2 * (size_t)n_r + n + 4 * (size_t)n;
    const bool staged = need * 4 <= FUZ_PHASE_SMEM && B.staging == 0;
    const bool sweep_staged = need_sweep * 4 <= FUZ_PHASE_SMEM && B.staging <= 1;
    uint32_t *sbits = smem;                                   // [nbw] phase bit per site
    int *s_loff = reinterpret_cast<int *>(smem + nbw);        // [n + 1] left CSR offsets (relative to e0)
    int *s_pk = s_loff + n + 1;                               // [n_e] packed left edges for the sweep
    uint8_t *s_slow = reinterpret_cast<uint8_t *>(s_pk + n_e);    // [n] site needs the generic sweep step (bit 0), in positions (bit 1)
    volatile uint32_t *s_fp = reinterpret_cast<uint32_t *>(s_pk + n_e + (n + 3) / 4 + 1);   // [n] forest pointers (contig-local)
    int *s_lq = s_pk + n_e + (n + 3) / 4 + 1 + n, *s_ld = s_lq + n_e;   // [n_e] partner (contig-local), cis - trans
    int *s_roff = s_ld + n_e;                                 // [n + 1] right row offsets (relative to r0)
    int *s_rq = s_roff + n + 1, *s_rd = s_rq + n_r;           // [n_r] partner (contig-local), cis - trans (0: rejected row)
    // [n] 1-based positions: with the rest of tier "full", else right behind the sweep tier / the phase bits when that fits (pass 3
    // looks up the position of every cis partner: from global memory that is a second dependent load per edge)
    const size_t pos_at = staged ? 0 : (sweep_staged ? need_sweep : (size_t)nbw);
    const bool pos_staged = staged || (pos_at + (size_t)n) * 4 <= FUZ_PHASE_SMEM;
    int *s_pos = staged ? s_rd + n_r : reinterpret_cast<int *>(smem + pos_at);
    int *s_sc = s_pos + n;                                    // [4n] lscore, rscore, lext, rext of pass 3
    if (B.dbg && tid == 0) B.dbg[c * 16 + 0] = clock64();
    for (int w = tid; w < nbw; w += nt) sbits[w] = 0;
    if (sweep_staged)
        for (int i = tid; i <= n; i += nt) s_loff[i] = B.left_off[cs0 + i] - e0;
    if (staged) {
        for (int i = tid; i <= n; i += nt) s_roff[i] = B.right_off[cs0 + i] - r0;
        for (int k = tid; k < n_e; k += nt) { s_lq[k] = B.lq[e0 + k] - cs0; s_ld[k] = B.ld[e0 + k]; }
        for (int k = tid; k < n_r; k += nt) {
            int d = row_d(O.d_at_ct, r0 + k);
            s_rq[k] = O.d_at_s2[r0 + k] - cs0; s_rd[k] = abs(d) >= 6 ? d : 0;
        }
        for (int i = tid; i < n; i += nt) s_pos[i] = O.d_site_pos[cs0 + i];
    } else if (pos_staged) {
        for (int i = tid; i < n; i += nt) s_pos[i] = O.d_site_pos[cs0 + i];
    }
    __syncthreads();
    // accessors (contig-local indices)
    auto loff = [&](int i) { return sweep_staged ? s_loff[i] : B.left_off[cs0 + i] - e0; };
    auto lq = [&](int k) { return staged ? s_lq[k] : B.lq[e0 + k] - cs0; };
    auto ld = [&](int k) { return staged ? s_ld[k] : B.ld[e0 + k]; };
    auto roff = [&](int i) { return staged ? s_roff[i] : B.right_off[cs0 + i] - r0; };
    auto rq = [&](int k) { return staged ? s_rq[k] : O.d_at_s2[r0 + k] - cs0; };
    auto rd = [&](int k) { if (staged) return s_rd[k]; int d = row_d(O.d_at_ct, r0 + k); return abs(d) >= 6 ? d : 0; };
    auto pos_of = [&](int i) { return pos_staged ? s_pos[i] : O.d_site_pos[cs0 + i]; };
    volatile uint32_t *fp = sweep_staged ? s_fp : reinterpret_cast<volatile uint32_t *>(B.fp + cs0);
    // smallest left partner of site i (slot in the left list), -1 if none
    auto min_left = [&](int i, int &d_out) {
        int best = -1, bq = 0x7fffffff;
        for (int k = loff(i); k < loff(i + 1); k++) { int q = lq(k); if (q < bq) { bq = q; best = k; } }
        d_out = best >= 0 ? ld(best) : 0;
        return best >= 0 ? bq : -1;
    };
    if (B.dbg && tid == 0) B.dbg[c * 16 + 1] = clock64();
    // ---- pass 1 as a forest (rows are ordered by (site1, site2), ties impossible)
    for (int i = tid; i < n; i += nt) {
        int parent = i, bit = 0, d;
        bool in_pos = false;
        int ml = min_left(i, d);
        uint8_t slow_flag = (loff(i + 1) - loff(i) > 32) || (ml >= 0 && i - 1 - ml >= 32);
        if (sweep_staged)
            for (int k = loff(i); k < loff(i + 1); k++) s_pk[k] = (ld(k) << 6) | ((i - 1 - lq(k)) & 31);
        else        // global tier: the same packed edges, indexed like the global CSR
            for (int k = loff(i); k < loff(i + 1); k++) B.g_pk[e0 + k] = (ld(k) << 6) | ((i - 1 - lq(k)) & 31);
        if (ml >= 0) {
            in_pos = true;
            parent = ml;
            bit = d < 0;                                       // trans > cis flips the state
        } else {
            for (int k = roff(i); k < roff(i + 1); k++) {
                int dr = rd(k);
                if (dr == 0) continue;
                in_pos = true;                                 // first accepted row = smallest right partner
                int rm = rq(k), d2;
                int mlr = min_left(rm, d2);
                // partner already has a state when i is first touched iff it has a left partner < i
                if (mlr >= 0 && mlr < i) { parent = rm; bit = dr < 0; }
                break;
            }
        }
        fp[i] = ((uint32_t)parent << 1) | (uint32_t)bit;
        O.d_ph_state[cs0 + i] = in_pos ? 0 : 255;
        if (sweep_staged) s_slow[i] = slow_flag | (in_pos ? 2 : 0);
        else B.g_slow[cs0 + i] = slow_flag | (in_pos ? 2 : 0);
    }
    __syncthreads();
    if (B.dbg && tid == 0) B.dbg[c * 16 + 2] = clock64();
    int rounds = 1;
    while ((1 << rounds) < n) rounds++;
    rounds++;
    for (int it = 0; it < rounds; it++) {
        for (int i = tid; i < n; i += nt) {
            uint32_t me = fp[i];
            int p = (int)(me >> 1);
            if (p != i) {
                uint32_t pp = fp[p];
                fp[i] = (pp & ~1u) | ((me ^ pp) & 1u);
            }
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt)
        if ((sweep_staged ? (s_slow[i] & 2) != 0 : O.d_ph_state[cs0 + i] != 255) && (fp[i] & 1u)) atomicOr(&sbits[i >> 5], 1u << (i & 31));
    __syncthreads();
    // ---- pass 2: one left-to-right sweep (a second sweep never changes anything).
    // Sequential by construction: one warp walks the sites and nothing hides its latency, so
    // the loop is kept to a few dozen instructions per site: the states of the last 32 sites
    // live in a warp-uniform bit window (`recent`), the adjacency of the next site is
    // prefetched, the vote is one REDUX, new states are written back one 32-site word at a time.
    if (B.dbg && tid == 0) B.dbg[c * 16 + 3] = clock64();
    // The sweep is a triangular system: the new state of a site is a function of the FINAL states of its left partners and
    // of its own initial state (ties keep it), new = sum_q (state(q) ? -d : d) < 0 ? 1 : > 0 ? 0 : initial.  Its unique
    // solution is also the fixed point of evaluating all sites in parallel, over and over, in place: a pass in which no
    // bit changes has reached it.  Few sites flip in practice, so a handful of passes of the whole CTA replaces the
    // site-by-site walk of one warp; contigs whose flips cascade for more than FUZ_SWEEP_PASSES passes restart from the
    // initial states and take the sequential walk.
    bool swept = false;
    if (B.sweep_passes > 0) {
        uint32_t *sinit = const_cast<uint32_t *>(fp);                 // the forest pointers are not needed any more
        for (int w = tid; w < nbw; w += nt) sinit[w] = sbits[w];
        __syncthreads();
        for (int pass = 0; pass < B.sweep_passes && !swept; pass++) {
            int changed = 0;
            for (int i = tid; i < n; i += nt) {
                int s0 = 0;
                for (int k = loff(i); k < loff(i + 1); k++) {
                    const int q = lq(k), d = ld(k);
                    s0 += ((sbits[q >> 5] >> (q & 31)) & 1u) ? -d : d;
                }
                const uint32_t cur = (sbits[i >> 5] >> (i & 31)) & 1u;
                const uint32_t nw = s0 < 0 ? 1u : s0 > 0 ? 0u : (sinit[i >> 5] >> (i & 31)) & 1u;
                if (nw != cur) { atomicXor(&sbits[i >> 5], 1u << (i & 31)); changed = 1; }
            }
            swept = !__syncthreads_or(changed);
        }
        if (!swept) {
            for (int w = tid; w < nbw; w += nt) sbits[w] = sinit[w];
            __syncthreads();
        }
    }
    if (!swept && warp == 0) {
        if (staged) sweep_sites_lean(n, lane, sbits, s_loff, s_lq, s_ld, 0, s_pk, s_slow);
        else if (sweep_staged) sweep_sites_lean(n, lane, sbits, s_loff, B.lq + e0, B.ld + e0, cs0, s_pk, s_slow);
        else        // global tier: the same loop over the global arrays (absolute CSR indices); its register pipeline
                    // also hides part of the L2 latency (627 -> 488 cycles per site on a 20 Mb contig)
            sweep_sites_lean(n, lane, sbits, B.left_off + cs0, B.lq, B.ld, cs0, B.g_pk, B.g_slow + cs0);
    }
    __syncthreads();
    if (B.dbg && tid == 0) B.dbg[c * 16 + 4] = clock64();
    // ---- pass 3: scores and extents, eight lanes per site, each over every eighth edge (positions = 1-based file positions);
    //      sums, min and max are exact in any order.  One thread per site walked ~120 edges serially (half of the kernel's time
    //      at 2 Mb contigs); a whole warp per site leaves too few independent chains of loads in flight (measured: 2x slower).
    for (int i0 = 0; i0 < n; i0 += nt >> 3) {
        const int i = i0 + (tid >> 3), sub = tid & 7;
        const bool live = i < n;
        const int x = cs0 + i;
        const bool in_pos = live && (sweep_staged ? (s_slow[i] & 2) != 0 : O.d_ph_state[x] != 255);
        const int px = live ? pos_of(i) : 0;
        int lscore = 0, rscore = 0, lext = px, rext = px;
        uint32_t sx = 0;
        if (in_pos) {
            sx = (sbits[i >> 5] >> (i & 31)) & 1u;
            const int l_end = loff(i + 1), r_end = roff(i + 1);
#pragma unroll 4
            for (int k = loff(i) + sub; k < l_end; k += 8) {
                const int q = lq(k), d = ld(k);
                const uint32_t sq = (sbits[q >> 5] >> (q & 31)) & 1u;
                const int dd = sq == sx ? d : -d;
                lscore += dd;
                if (dd > 0) lext = min(lext, pos_of(q));
            }
#pragma unroll 4
            for (int k = roff(i) + sub; k < r_end; k += 8) {
                const int d = rd(k);
                if (d == 0) continue;
                const int q = rq(k);
                const uint32_t sq = (sbits[q >> 5] >> (q & 31)) & 1u;
                const int dd = sq == sx ? d : -d;
                rscore += dd;
                if (dd > 0) rext = max(rext, pos_of(q));
            }
        }
        __syncwarp();
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {                   // butterflies inside the group of eight lanes
            lscore += __shfl_xor_sync(0xffffffffu, lscore, d); rscore += __shfl_xor_sync(0xffffffffu, rscore, d);
            lext = min(lext, __shfl_xor_sync(0xffffffffu, lext, d)); rext = max(rext, __shfl_xor_sync(0xffffffffu, rext, d));
        }
        if (live && sub == 0) {
            if (in_pos) O.d_ph_state[x] = (uint8_t)sx;
            O.d_ph_lscore[x] = lscore; O.d_ph_rscore[x] = rscore; O.d_ph_lext[x] = lext; O.d_ph_rext[x] = rext;
            if (staged) { s_sc[4 * i] = lscore; s_sc[4 * i + 1] = rscore; s_sc[4 * i + 2] = lext; s_sc[4 * i + 3] = rext; }
        }
    }
    __syncthreads();
    // ---- pass 4: chain sites into blocks by the running maximum of right extents (never reset)
    //      = exclusive prefix max + boundary flags + segment sizes + dense ids of blocks with > 3 sites
    if (B.dbg && tid == 0) B.dbg[c * 16 + 5] = clock64();
    int carry_max = 0, carry_b = 0;
    for (int base = 0; base < n; base += nt) {
        const int x = cs0 + base + tid, il = base + tid;
        const bool valid = il < n;
        bool f;
        int my_rext = 0, my_lext = 0;
        if (staged) {
            f = valid && (s_slow[il] & 2) && s_sc[4 * il + 1] >= 10 && s_sc[4 * il] >= 10;
            if (f) { my_lext = s_sc[4 * il + 2]; my_rext = s_sc[4 * il + 3]; }
        } else {
            f = valid && O.d_ph_state[x] != 255 && O.d_ph_rscore[x] >= 10 && O.d_ph_lscore[x] >= 10;
            if (f) { my_lext = O.d_ph_lext[x]; my_rext = O.d_ph_rext[x]; }
        }
        int tot;
        int mb = max(carry_max, cta_excl_max(f ? my_rext : 0, s_tmp, &tot));
        carry_max = max(carry_max, tot);
        const int boundary = f && mb < my_lext;
        int bi = carry_b + cta_excl_sum(boundary, s_tmp, &tot) + boundary;   // inclusive: temp block number
        carry_b += tot;
        if (valid) B.bidx[x] = f ? bi : 0;
        if (f) atomicAdd(&B.bsize[cs0 + bi], 1);       // bi <= number of filtered sites <= n
    }
    __threadfence_block();
    __syncthreads();
    int carry_id = 0;
    for (int base = 0; base <= n; base += nt) {        // temp block numbers 0..n
        const int b = base + tid;
        const int keep = b >= 1 && b <= n && B.bsize[cs0 + b] > 3;
        int tot;
        int id = carry_id + cta_excl_sum(keep, s_tmp, &tot);
        carry_id += tot;
        if (b <= n) B.bnew[cs0 + c + b] = keep ? id + 1 : 0;
    }
    __threadfence_block();
    __syncthreads();
    for (int x = cs0 + tid; x < cs1; x += nt) {
        int bi = B.bidx[x];
        O.d_ph_block[x] = bi > 0 ? B.bnew[cs0 + c + bi] : 0;
    }
    if (B.dbg && tid == 0) B.dbg[c * 16 + 6] = clock64();
}

// ================================================================== phased reads
struct ReadScratch {
    int32_t *row_off, *ctg_q_off, *q_cnt, *q_off, *q_cur, *q_ent, *pr_cnt, *pr_off;
    uint8_t *dup;
    int64_t total_nq;
};

__global__ void k_rd_init(ReadScratch R) {
    fuz_pdl_enter();
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q <= R.total_nq; q += (int64_t)gridDim.x * blockDim.x) {
        R.q_cnt[q] = 0;
        if (q < R.total_nq) R.q_cur[q] = 0;
    }
}

// rows that can vote: first occurrence of (site, allele, q_id).  Whether the site lies inside a block is
// decided later (k_vote): the per-read lists do not depend on the phasing result, so they can be built
// while the block stage runs.
__device__ __forceinline__ int voting_gq(int i, const ReadScratch &R, const fuz_outputs &O, fuz_status *st) {
    if (R.dup[i]) return -1;
    const int s = O.d_vm_site[i];
    const int c = O.d_site_ctg[s];
    const int q = O.d_vm_qid[i];
    if (q < 0 || q >= R.ctg_q_off[c + 1] - R.ctg_q_off[c]) { fuz_raise(st, FUZ_E_FORMAT, i); return -1; }
    return R.ctg_q_off[c] + q;
}

__global__ void k_q_count(ReadScratch R, fuz_outputs O, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int n_vmap = (int)st->n_vmap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vmap; i += gridDim.x * blockDim.x) {
        int gq = voting_gq(i, R, O, st);
        if (gq >= 0) atomicAdd(&R.q_cnt[gq], 1);
    }
}

__global__ void k_q_fill(ReadScratch R, fuz_outputs O, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int n_vmap = (int)st->n_vmap;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vmap; i += gridDim.x * blockDim.x) {
        int gq = voting_gq(i, R, O, st);
        if (gq >= 0)       // entry = (site << 2) | base; k_vote turns it into (block << 1) | phase
            R.q_ent[R.q_off[gq] + atomicAdd(&R.q_cur[gq], 1)] = (O.d_vm_site[i] << 2) | (int)O.d_vm_base[i];
    }
}

// One thread per entry of the per-read lists: (site << 2) | base becomes (block << 1) | phase of the
// variant -- phase 0 if the row's base is the hap-0 allele of the site's state (phasing.py:462-463); 0 for a
// site outside every block.  Runs after the block stage (the lists themselves are built beside it).
__global__ void k_q_pack(ReadScratch R, fuz_outputs O, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int n_ent = R.q_off[R.total_nq];
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_ent; e += gridDim.x * blockDim.x) {
        const int v = R.q_ent[e], s = v >> 2;
        const int blk = O.d_ph_block[s];
        int packed = 0;
        if (blk > 0) {
            const uint8_t h0 = O.d_ph_state[s] == 0 ? O.d_site_al[2 * s] : O.d_site_al[2 * s + 1];
            packed = (blk << 1) | ((uint8_t)(v & 3) == h0 ? 0 : 1);
        }
        R.q_ent[e] = packed;
    }
}

// One thread per read: distinct phased variants -> (block, phase) counts; blocks in
// ascending id; a row when |n0 - n1| > 1 (phasing.py:465-480).  fill = 0 counts rows.
__global__ void k_vote(ReadScratch R, fuz_outputs O, int n_ctg, int fill, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    for (int64_t gq = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; gq < R.total_nq; gq += (int64_t)gridDim.x * blockDim.x) {
        const int e0 = R.q_off[gq], e1 = R.q_off[gq + 1];
        int rows = 0, last = 0;
        int64_t out = fill ? R.pr_off[gq] : 0;
        int c = -1;
        while (e0 < e1) {
            int cur = 0x7fffffff;
            for (int e = e0; e < e1; e++) {
                int b = R.q_ent[e] >> 1;
                if (b > last && b < cur) cur = b;
            }
            if (cur == 0x7fffffff) break;
            int n0 = 0, n1 = 0;
            for (int e = e0; e < e1; e++) {
                int v = R.q_ent[e];
                if ((v >> 1) == cur) { if (v & 1) n1++; else n0++; }
            }
            int phase = n0 - n1 > 1 ? 0 : (n1 - n0 > 1 ? 1 : -1);
            if (phase >= 0) {
                if (fill && out < O.cap_reads) {
                    if (c < 0) c = fuz_upper_bound(R.ctg_q_off, 0, n_ctg + 1, (int)gq) - 1;
                    O.d_pr_ctg[out] = c; O.d_pr_qid[out] = (int)gq - R.ctg_q_off[c];
                    O.d_pr_block[out] = cur; O.d_pr_phase[out] = phase; O.d_pr_n0[out] = n0; O.d_pr_n1[out] = n1;
                }
                out++; rows++;
            }
            last = cur;
        }
        if (!fill) R.pr_cnt[gq] = rows;
    }
}

}  // namespace

// ------------------------------------------------------------------ host side
int fuz_association_impl(fuz_ctx *ctx, int32_t n_ctg, fuz_outputs *out, bool row_off_valid) {
    (void)n_ctg;
    cudaStream_t st = ctx->stream;
    const int64_t cs = out->cap_sites, cv = out->cap_vmap;
    AssocScratch A;
    A.max_pairs = ctx->max_pairs_per_site * (cs > 0 ? cs : 1);
    FuzLayout L;
    size_t o_uq = L.add(4 * (size_t)(cv + 1)), o_uqn = L.add(8 * (size_t)(cs + 1));
    size_t o_na0 = L.add(4 * (size_t)(cs + 1)), o_qmin = L.add(4 * (size_t)(cs + 1)), o_qmax = L.add(4 * (size_t)(cs + 1));
    size_t o_cc = L.add(4 * (size_t)(cs + 2)), o_co = L.add(4 * (size_t)(cs + 2));
    size_t o_ac = L.add(4 * (size_t)(cs + 2)), o_ao = L.add(4 * (size_t)(cs + 2));
    size_t o_pc = L.add(16 * (size_t)(A.max_pairs + 1));
    size_t o_qm = L.add(4 * 2 * FUZ_QM_WORDS * (size_t)(cs + 1)), o_hm = L.add((size_t)cs + 1);
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    if ((rc = fuz_keep_commit(ctx, cs, cv, &A.row_off, &A.dup, &A.at_off))) return rc;   // at_off = row range per left site, reused by the block stage
    A.site_ctg = out->d_site_ctg; A.site_pos = out->d_site_pos;
    A.uq = fuz_at<int32_t>(ctx, o_uq); A.uq_n = fuz_at<int32_t>(ctx, o_uqn);
    A.na0 = fuz_at<int32_t>(ctx, o_na0); A.qmin = fuz_at<int32_t>(ctx, o_qmin); A.qmax = fuz_at<int32_t>(ctx, o_qmax);
    A.cand_cnt = fuz_at<int32_t>(ctx, o_cc); A.cand_off = fuz_at<int32_t>(ctx, o_co);
    A.at_cnt = fuz_at<int32_t>(ctx, o_ac); (void)o_ao;
    A.pair_ct = fuz_at<int4>(ctx, o_pc);
    A.qmask = fuz_at<uint32_t>(ctx, o_qm); A.has_mask = fuz_at<uint8_t>(ctx, o_hm);
    const int64_t *d_ns = &ctx->d_status->n_sites;

    if (!row_off_valid) {
        fuz_launch(ctx, k_site_rowoff, FUZ_GRID_BLOCKS, 256, 0, st, out->d_vm_site, A.row_off, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_site_rowoff");
    }
    fuz_launch(ctx, k_uniq_lists, 148 * ctx->grid_assoc, 256, 0, st, out->d_site_al, out->d_vm_base, out->d_vm_qid, A, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_uniq_lists");
    if ((rc = fuz_scan_i32_wide(ctx, A.cand_cnt, A.cand_off, cs, FUZ_FIN_PAIRS, A.max_pairs, d_ns))) return rc;
    fuz_launch(ctx, k_pair_count, 148 * ctx->grid_assoc, 256, 0, st, A, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_pair_count");
    if ((rc = fuz_scan_i32_wide(ctx, A.at_cnt, A.at_off, cs, FUZ_FIN_ATABLE, out->cap_atable, d_ns))) return rc;
    fuz_launch(ctx, k_pair_fill, 148 * ctx->grid_assoc, 256, 0, st, A, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_pair_fill");
    return FUZ_OK;
}

int fuz_blocks_impl(fuz_ctx *ctx, int32_t n_ctg, fuz_outputs *out, bool at_off_valid) {
    cudaStream_t st = ctx->stream;
    const int64_t cs = out->cap_sites, ca = out->cap_atable;
    BlockScratch B;
    FuzLayout L;
    size_t o_cso = L.add(4 * (size_t)(n_ctg + 2)), o_lc = L.add(4 * (size_t)(cs + 2)), o_lo = L.add(4 * (size_t)(cs + 2));
    size_t o_lcur = L.add(4 * (size_t)(cs + 1)), o_lq = L.add(4 * (size_t)(ca + 1)), o_ld = L.add(4 * (size_t)(ca + 1));
    size_t o_ro = L.add(4 * (size_t)(cs + 2)), o_mk = L.add(4 * (size_t)(cs + 1)), o_fp = L.add(4 * (size_t)(cs + 1));
    size_t o_bi = L.add(4 * (size_t)(cs + 1)), o_bs = L.add(4 * (size_t)(cs + 2)), o_bn = L.add(4 * (size_t)(cs + n_ctg + 2));
    size_t o_dbg = L.add(8 * 16 * (size_t)n_ctg);
    size_t o_gpk = L.add(4 * (size_t)(ca + 1)), o_gslow = L.add((size_t)cs + 1);
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    B.g_pk = fuz_at<int32_t>(ctx, o_gpk); B.g_slow = fuz_at<uint8_t>(ctx, o_gslow);
    B.ctg_site_off = fuz_at<int32_t>(ctx, o_cso); B.left_cnt = fuz_at<int32_t>(ctx, o_lc); B.left_off = fuz_at<int32_t>(ctx, o_lo);
    B.left_cur = fuz_at<int32_t>(ctx, o_lcur); B.lq = fuz_at<int32_t>(ctx, o_lq); B.ld = fuz_at<int32_t>(ctx, o_ld);
    B.right_off = fuz_at<int32_t>(ctx, o_ro);
    if (at_off_valid) {      // the association stage left the row range of every left site in the inter-stage buffer
        int32_t *ro; uint8_t *dp; int32_t *ao;
        if ((rc = fuz_keep_commit(ctx, cs, out->cap_vmap, &ro, &dp, &ao))) return rc;
        B.right_off = ao;
    } B.min_k = fuz_at<int32_t>(ctx, o_mk); B.fp = fuz_at<uint32_t>(ctx, o_fp);
    B.bidx = fuz_at<int32_t>(ctx, o_bi); B.bsize = fuz_at<int32_t>(ctx, o_bs); B.bnew = fuz_at<int32_t>(ctx, o_bn);
    B.dbg = ctx->profile ? fuz_at<long long>(ctx, o_dbg) : nullptr;
    B.staging = ctx->phase_staging;
    B.sweep_passes = ctx->sweep_passes;
    if (!ctx->phase_attr_set) {
        FUZ_CUDA(ctx, cudaFuncSetAttribute(k_ctg_phase, cudaFuncAttributeMaxDynamicSharedMemorySize, FUZ_PHASE_SMEM));
        ctx->phase_attr_set = true;
    }
    fuz_launch(ctx, k_blk_init, FUZ_GRID_BLOCKS, 256, 0, st, B, *out, n_ctg, at_off_valid ? 1 : 0, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_blk_init");
    fuz_launch(ctx, k_edge_count, FUZ_GRID_BLOCKS, 256, 0, st, B, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_edge_count");
    if ((rc = fuz_scan_i32_wide(ctx, B.left_cnt, B.left_off, cs, FUZ_FIN_NONE, 0, &ctx->d_status->n_sites))) return rc;
    fuz_launch(ctx, k_edge_fill, FUZ_GRID_BLOCKS, 256, 0, st, B, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_edge_fill");
    fuz_launch(ctx, k_ctg_phase, n_ctg, FUZ_PHASE_THREADS, FUZ_PHASE_SMEM, st, B, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_ctg_phase");
    if (B.dbg) {            // diagnostics: phase durations (cycles) of contig 0
        long long h[16];
        cudaMemcpyAsync(h, B.dbg, sizeof(h), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "k_ctg_phase contig 0 cycles: stage %lld pass1 %lld jump %lld sweep %lld pass3 %lld pass4 %lld\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4], h[6] - h[5]);
    }
    return FUZ_OK;
}

// The read stage in two halves.  fuz_reads_csr builds the per-read lists of voting rows; it needs the
// variant_map rows and the duplicate flags only, so fuz_phase_batch runs it on the side stream while the
// block stage runs.  fuz_reads_vote needs the blocks.  Scratch comes from a buffer of its own (the arena
// belongs to the stage running on the main stream).
static int reads_scratch(fuz_ctx *ctx, int32_t n_ctg, int64_t total_nq, fuz_outputs *out, ReadScratch &R) {
    const int64_t cs = out->cap_sites, cv = out->cap_vmap;
    R.total_nq = total_nq;
    FuzLayout L;
    const size_t o_cq = L.add(4 * (size_t)(n_ctg + 2));
    const size_t o_qc = L.add(4 * (size_t)(total_nq + 2)), o_qo = L.add(4 * (size_t)(total_nq + 2)), o_qcur = L.add(4 * (size_t)(total_nq + 1));
    const size_t o_qe = L.add(4 * (size_t)(cv + 1)), o_pc = L.add(4 * (size_t)(total_nq + 2)), o_po = L.add(4 * (size_t)(total_nq + 2));
    if (L.off > ctx->reads_cap) {
        FUZ_CUDA(ctx, cudaDeviceSynchronize());
        if (ctx->reads_buf) FUZ_CUDA(ctx, cudaFree(ctx->reads_buf));
        ctx->reads_buf = nullptr; ctx->reads_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->reads_buf, L.off + (L.off >> 2));
        if (e != cudaSuccess) return fuz_fail(ctx, FUZ_E_CUDA, "read-stage buffer of %zu bytes: %s", L.off, cudaGetErrorString(e));
        ctx->reads_cap = L.off + (L.off >> 2);
    }
    int rc = fuz_keep_commit(ctx, cs, cv, &R.row_off, &R.dup);
    if (rc) return rc;
    uint8_t *base = ctx->reads_buf;
    R.ctg_q_off = reinterpret_cast<int32_t *>(base + o_cq);
    R.q_cnt = reinterpret_cast<int32_t *>(base + o_qc); R.q_off = reinterpret_cast<int32_t *>(base + o_qo);
    R.q_cur = reinterpret_cast<int32_t *>(base + o_qcur); R.q_ent = reinterpret_cast<int32_t *>(base + o_qe);
    R.pr_cnt = reinterpret_cast<int32_t *>(base + o_pc); R.pr_off = reinterpret_cast<int32_t *>(base + o_po);
    return FUZ_OK;
}

int fuz_reads_csr(fuz_ctx *ctx, int32_t n_ctg, const int32_t *d_ctg_nq, int64_t total_nq, fuz_outputs *out, bool dup_valid) {
    cudaStream_t st = ctx->stream;
    ReadScratch R;
    int rc = reads_scratch(ctx, n_ctg, total_nq, out, R);
    if (rc) return rc;
    if ((rc = fuz_scan_i32(ctx, d_ctg_nq, R.ctg_q_off, n_ctg, nullptr, FUZ_FIN_NONE, 0))) return rc;
    if (!dup_valid) {
        fuz_launch(ctx, k_site_rowoff, FUZ_GRID_BLOCKS, 256, 0, st, out->d_vm_site, R.row_off, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_site_rowoff");
        fuz_launch(ctx, k_dup_flags, FUZ_GRID_BLOCKS, 256, 0, st, R.row_off, out->d_vm_base, out->d_vm_qid, R.dup, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_dup_flags");
    }
    fuz_launch(ctx, k_rd_init, FUZ_GRID_BLOCKS, 256, 0, st, R);
    FUZ_LAUNCH_CHECK(ctx, "k_rd_init");
    fuz_launch(ctx, k_q_count, 148 * ctx->grid_reads, 256, 0, st, R, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_q_count");
    if ((rc = fuz_scan_i32(ctx, R.q_cnt, R.q_off, total_nq, nullptr, FUZ_FIN_NONE, 0))) return rc;
    fuz_launch(ctx, k_q_fill, 148 * ctx->grid_reads, 256, 0, st, R, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_q_fill");
    return FUZ_OK;
}

int fuz_reads_vote(fuz_ctx *ctx, int32_t n_ctg, int64_t total_nq, fuz_outputs *out) {
    cudaStream_t st = ctx->stream;
    ReadScratch R;
    int rc = reads_scratch(ctx, n_ctg, total_nq, out, R);
    if (rc) return rc;
    fuz_launch(ctx, k_q_pack, 148 * ctx->grid_reads, 256, 0, st, R, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_q_pack");
    fuz_launch(ctx, k_vote, 148 * ctx->grid_reads, 256, 0, st, R, *out, n_ctg, 0, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_vote(count)");
    if ((rc = fuz_scan_i32_wide(ctx, R.pr_cnt, R.pr_off, total_nq, FUZ_FIN_READS, out->cap_reads))) return rc;
    fuz_launch(ctx, k_vote, 148 * ctx->grid_reads, 256, 0, st, R, *out, n_ctg, 1, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_vote(fill)");
    return FUZ_OK;
}

int fuz_reads_impl(fuz_ctx *ctx, int32_t n_ctg, const int32_t *d_ctg_nq, int64_t total_nq, fuz_outputs *out,
                   bool dup_valid) {
    int rc = fuz_reads_csr(ctx, n_ctg, d_ctg_nq, total_nq, out, dup_valid);
    return rc ? rc : fuz_reads_vote(ctx, n_ctg, total_nq, out);
}

static int set_counts(fuz_ctx *ctx, int64_t n_sites, int64_t n_vmap, int64_t n_atable) {
    fuz_launch(ctx, k_set_counts, 1, 32, 0, ctx->stream, ctx->d_status, n_sites, n_vmap, n_atable, 1);
    FUZ_LAUNCH_CHECK(ctx, "k_set_counts");
    return FUZ_OK;
}

extern "C" int fuz_association_table(fuz_ctx *ctx, int32_t n_ctg, int64_t n_sites, int64_t n_vmap, fuz_outputs *out) {
    if (!ctx || !out || n_sites < 0 || n_vmap < 0 || n_sites > out->cap_sites || n_vmap > out->cap_vmap)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_association_table: bad arguments");
    int rc = set_counts(ctx, n_sites, n_vmap, 0);
    return rc ? rc : fuz_association_impl(ctx, n_ctg, out, false);
}

extern "C" int fuz_phased_blocks(fuz_ctx *ctx, int32_t n_ctg, int64_t n_sites, int64_t n_atable, fuz_outputs *out) {
    if (!ctx || !out || n_ctg < 1 || n_sites < 0 || n_atable < 0 || n_sites > out->cap_sites || n_atable > out->cap_atable)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_phased_blocks: bad arguments");
    int rc = set_counts(ctx, n_sites, -1, n_atable);
    return rc ? rc : fuz_blocks_impl(ctx, n_ctg, out, false);
}

extern "C" int fuz_phased_reads(fuz_ctx *ctx, int32_t n_ctg, const int32_t *d_ctg_nq, int64_t total_nq, int64_t n_sites,
                                int64_t n_vmap, fuz_outputs *out) {
    if (!ctx || !out || n_ctg < 1 || !d_ctg_nq || total_nq < 0 || n_sites < 0 || n_vmap < 0 || n_sites > out->cap_sites ||
        n_vmap > out->cap_vmap)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_phased_reads: bad arguments");
    int rc = set_counts(ctx, n_sites, n_vmap, -1);
    return rc ? rc : fuz_reads_impl(ctx, n_ctg, d_ctg_nq, total_nq, out, false);
}
