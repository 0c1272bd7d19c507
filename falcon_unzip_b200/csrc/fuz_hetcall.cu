// Stage 1 of the phasing path: record scan + filter, pileup, het call, variant_map rows.
// Reproduces reference falcon_unzip/phasing.py:14-134 (make_het_call) for a batch of
// contigs; see DESIGN.md for the data layout and the equivalence argument
// (streaming flush == full pileup evaluated for pos < POS_last).
#include <algorithm>

#include "fuz_internal.cuh"

namespace {

// scratch of the het-call stage (device pointers into the context arena)
struct FuzTileEnt;
struct HetScratch {
    int32_t *r_gstart, *r_gend, *r_nwords, *r_woff;
    int64_t *r_seq;
    uint8_t *r_flags;
    uint32_t *proj;            // reference-aligned 4-bit projection of every accepted record
    int64_t proj_cap;          // capacity of proj in words
    int32_t *ctg_last_rec, *ctg_maxspan;
    int32_t *rec_cursor;       // next record to hand to a warp of k_project
    int32_t *tile_ctg, *tile_rlo, *tile_rhi, *tile_limit, *tile_site_base, *tile_site_cnt, *tile_site_off;
    int32_t *us_gpos, *us_tile;
    uint32_t *us_cnt;
    int32_t *s_gpos, *site_rows, *site_row_off;
    uint32_t *counts;
    int32_t n_tiles;
    // segment-list pileup (fuz_pileup_seg.cuh)
    int4 *segs;                // match segments of every record: {ref start, ref end, query - ref, 0}
    int64_t seg_cap;
    int32_t *r_seg_off, *r_nseg;
    struct FuzTileEnt *ents;   // reads of every tile
    int64_t ent_cap;
    int32_t *tile_ent_base, *tile_ent_cnt;
    int32_t *tile_cursor;      // next tile to hand to a CTA of k_pileup_tma
    uint32_t *spill;           // 16-bit overflow counters of k_pileup_tma, one block per CTA
};

__device__ __forceinline__ bool op_is_match(uint32_t op) { return op == 0 || op == 7 || op == 8; }

// ---------------------------------------------------------------- init
// first word (8 positions) of a record's projection row: the row starts on a 16-byte boundary of the GLOBAL word grid, so
// that the part of a row inside a pileup tile is a 16-byte aligned run at a 16-byte aligned tile offset (bulk copies)
__device__ __forceinline__ int fuz_row_w0(int gstart) { return (gstart >> 3) & ~3; }

__global__ void k_het_init(fuz_status *st, int32_t *ctg_last_rec, int32_t *ctg_maxspan, int32_t *rec_cursor, int n_ctg) {   // rec_cursor[1] = tile cursor
    fuz_pdl_enter();
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        rec_cursor[0] = 0; rec_cursor[1] = 0;
        st->error = 0; st->error_index = 0;
        st->n_sites = st->n_vmap = st->n_atable = st->n_reads = 0;
        st->need_sites = st->need_vmap = st->need_atable = st->need_reads = st->need_pairs = 0;
        st->n_accepted = 0; st->aligned_bases = 0; st->n_segments = 0; st->reserved[0] = 0;
    }
    for (; i < n_ctg; i += gridDim.x * blockDim.x) { ctg_last_rec[i] = -1; ctg_maxspan[i] = 0; }
}

// candidate record range, contig and evaluation limit of every pileup tile.  The limit is
// POS_last of the contig in global coordinates: the start of the last accepted record in
// file order; positions >= POS_last are never evaluated (no final flush, phasing.py:98-129).
// One WARP per tile: the searches are 32-ary (every lane probes one pivot, a ballot picks the interval), so a tile
// costs ~7 dependent memory round trips instead of ~25 with one thread and binary searches.
// first index in [lo, hi) with a[i] >= key (a ascending), all lanes of the warp cooperate
template <typename T>
__device__ __forceinline__ int fuz_warp_lower_bound(const T *__restrict__ a, int lo, int hi, T key, int lane) {
    while (hi - lo > 32) {
        const int n = hi - lo, step = (n + 31) >> 5;                 // 32 pivots: lo + (k + 1) * step - 1
        const int p = min(lo + (lane + 1) * step - 1, hi - 1);
        const uint32_t ge = __ballot_sync(0xffffffffu, a[p] >= key);   // monotone: a suffix of the lanes
        const int k = ge ? __ffs(ge) - 1 : 32;                         // first pivot >= key
        const int nlo = lo + k * step;                                 // everything before pivot k-1 (inclusive) is < key
        if (k < 32) hi = min(lo + (k + 1) * step - 1, hi - 1) + 1;
        lo = min(nlo, hi);
        if (k == 32) break;                                            // every pivot < key: the answer is hi
    }
    const int i = lo + lane;
    const uint32_t ge = __ballot_sync(0xffffffffu, i < hi && a[i] >= key);
    return ge ? lo + __ffs(ge) - 1 : hi;
}

__global__ void __launch_bounds__(256) k_tile_ranges(int n_tiles, int n_ctg, const int64_t *__restrict__ ctg_goff,
                                                     const int32_t *__restrict__ ctg_rec_off, HetScratch S, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles; t += (gridDim.x * blockDim.x) >> 5) {
        const int64_t t0 = (int64_t)t * FUZ_PTILE;
        // last c with ctg_goff[c] <= t0  =  (first c with ctg_goff[c] >= t0 + 1) - 1
        int c = fuz_warp_lower_bound<int64_t>(ctg_goff, 0, n_ctg + 1, t0 + 1, lane) - 1;
        if (c >= n_ctg) c = n_ctg - 1;
        const int r0 = ctg_rec_off[c], r1 = ctg_rec_off[c + 1];
        const int t1 = (int)(t0 + FUZ_PTILE);
        const int span = S.ctg_maxspan[c];
        const int lr = S.ctg_last_rec[c];
        const int rlo = fuz_warp_lower_bound<int32_t>(S.r_gstart, r0, r1, (int)t0 - span + 1, lane);
        const int rhi = fuz_warp_lower_bound<int32_t>(S.r_gstart, r0, r1, t1, lane);
        if (lane == 0) {
            S.tile_ctg[t] = c;
            S.tile_limit[t] = lr >= 0 ? S.r_gstart[lr] : (int32_t)ctg_goff[c];
            S.tile_rlo[t] = rlo;
            S.tile_rhi[t] = rhi;
        }
    }
}

// ---------------------------------------------------------------- het test + site emission
// key = count*4 + base index: descending key order == the reference's sort()+reverse() on
// (count, base) tuples (phasing.py:116-117; T > G > C > A on ties).
__device__ __forceinline__ bool het_test(uint32_t cA, uint32_t cC, uint32_t cG, uint32_t cT) {
    uint32_t total = cA + cC + cG + cT;
    if (total < 10) return false;                                    // phasing.py:112
    uint32_t k[4] = {cA << 2, (cC << 2) | 1, (cG << 2) | 2, (cT << 2) | 3};
    uint32_t m0 = max(max(k[0], k[1]), max(k[2], k[3]));
    uint32_t m1 = 0;
#pragma unroll
    for (int b = 0; b < 4; b++) if (k[b] != m0) m1 = max(m1, k[b]);
    uint32_t c0 = m0 >> 2, c1 = m1 >> 2;
    // c0/total < 0.75 and c1/total > 0.25 (phasing.py:118-120), exact in integers (B.2)
    return 4 * c0 < 3 * total && 4 * c1 > total;
}

// Ordered compaction of the het positions of one tile: every thread owns 8 consecutive
// positions (bit i of hetmask: position i is a het site with counts cnt[i][0..3]); sites of
// a tile land contiguously (in position order) at an atomically claimed base; tiles are put
// in order afterwards.
template <bool NAMED_BARRIER = false>      // true: the 256 consumer threads of k_pileup_gather_tma (bar.sync 1), else the whole CTA
__device__ __forceinline__ void emit_tile_sites(uint32_t hetmask, const uint32_t (&cnt)[8][4], int tile, int t0,
                                                HetScratch &S, int64_t cap_sites, fuz_status *st,
                                                int *s_warp_tot, int *s_base) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int nh = __popc(hetmask);
    int incl = fuz_warp_incl_scan(nh, lane);
    if (lane == 31) s_warp_tot[warp] = incl;
    if (NAMED_BARRIER) { __syncwarp(); asm volatile("bar.sync 1, 256;" ::: "memory"); } else __syncthreads();
    if (warp == 0) {
        int t = lane < FUZ_NW ? s_warp_tot[lane] : 0;
        int ti = fuz_warp_incl_scan(t, lane);
        if (lane < FUZ_NW) s_warp_tot[lane] = ti - t;
        if (lane == FUZ_NW - 1) {
            int total = ti;
            int base = 0;
            if (total > 0) base = (int)atomicAdd((unsigned long long *)&st->need_sites, (unsigned long long)total);
            S.tile_site_base[tile] = base;
            S.tile_site_cnt[tile] = total;
            *s_base = base;
        }
    }
    if (NAMED_BARRIER) { __syncwarp(); asm volatile("bar.sync 1, 256;" ::: "memory"); } else __syncthreads();
    if (nh) {
        int64_t o = (int64_t)*s_base + s_warp_tot[warp] + (incl - nh);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (hetmask & (1u << i)) {
                if (o < cap_sites) {
                    S.us_gpos[o] = t0 + tid * 8 + i;
                    S.us_tile[o] = tile;
                    reinterpret_cast<uint4 *>(S.us_cnt)[o] = make_uint4(cnt[i][0], cnt[i][1], cnt[i][2], cnt[i][3]);
                }
                o++;
            }
        }
    }
}

__device__ __forceinline__ uint32_t het_mask_of(const uint32_t (&cnt)[8][4], int t0, int pos_limit) {
    uint32_t hetmask = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int p = t0 + (int)threadIdx.x * 8 + i;
        if (p < pos_limit && het_test(cnt[i][0], cnt[i][1], cnt[i][2], cnt[i][3])) hetmask |= 1u << i;
    }
    return hetmask;
}

// ---------------------------------------------------------------- projection (read major)
// 8 nibbles (query bases q0..q0+7, BAM 4-bit codes) -> one word, base i at bits 4i.
__device__ __forceinline__ uint32_t swap_nibbles(uint32_t x) {      // BAM: first base of a byte in the HIGH nibble
    return ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
}
// zero every nibble that is not one of A=1 C=2 G=4 T=8 (ambiguity codes never count as a
// base in the pileup, phasing.py:108-111, and never match a called allele)
__device__ __forceinline__ uint32_t multi_bits(uint32_t x) {     // non-zero inside nibbles with 2+ bits set
    return (x & (x >> 1) & 0x77777777u) | (x & (x >> 2) & 0x33333333u) | (x & (x >> 3) & 0x11111111u);
}
// non-zero iff some nibble of x has two or more bits set: per nibble n & (n - 1), the decrement done on (n | 8) so that no
// borrow crosses a nibble ((n | 8) - 1 keeps n - 1 in the low three bits, and bit 3 of it is set exactly when n > 8)
__device__ __forceinline__ uint32_t ambiguous_nibbles(uint32_t x) { return x & ((x | 0x88888888u) - 0x11111111u); }
__device__ __forceinline__ uint32_t keep_acgt(uint32_t x) {
    uint32_t any = multi_bits(x);
    if (any) {
        uint32_t f = (any | (any >> 1) | (any >> 2)) & 0x11111111u;
        x &= ~(f * 15u);
    }
    return x;
}

#define FUZ_SEGCAP 256         // match segments buffered per warp before their quads are emitted

// The match segments of a record (maximal runs of M/=/X; reference [rs, re), query offset
// dq = query position - reference position) sit in a per-warp shared-memory list ordered by
// rs.  The projection is produced in quads (4 words = 32 reference positions, one 128-bit
// store).  Every quad has at most one OWNER, the first segment that intersects it; a segment
// is a non-owner only in the quad it starts in (when an earlier segment ends there).  So:
//   phase 1: one lane per quad, rows of 32 consecutive quads: look up the owner, cut its 32
//            query nibbles out of SEQ (five aligned 32-bit loads, nibble swap, funnel shift),
//            mask to the segment, store the quad (zeros where nothing lands: deletions, padding);
//   phase 2: one lane per segment that starts in a quad it does not own: same cut, OR-merged
//            into the stored quad (atomicOr; the words are still in L2).
struct ProjRec {
    const uint32_t *base4;     // SEQ address rounded down to 4 bytes
    int nphase;                // nibble index of SEQ[0] relative to base4 (0, 2, 4, 6)
    int W0;                    // first word of the record on the global 8-position grid
    uint32_t *out;             // the record's projection
    const uint32_t *lim;       // end of the record (prefetches stay below it)
};

// the piece of segment (a, b, dq) inside the quad starting at reference position Q0:
// positions [lo, hi) of the quad (hi <= lo: nothing), ambiguity codes removed
__device__ __forceinline__ void quad_piece(const ProjRec &R, int Q0, int a, int b, int dq, uint32_t (&v)[4]) {
    const int lo4 = 4 * max(a - Q0, 0), hi4 = 4 * min(b - Q0, 32);
    const int n = Q0 + dq + R.nphase;                     // nibble of position Q0 relative to base4 (>= -31)
    const uint32_t *src = R.base4 + (n >> 3);
    const uint32_t sh = (uint32_t)(n & 7) * 4;
    uint32_t m[5];
#pragma unroll
    for (int k = 0; k < 5; k++) m[k] = __ldg(src + k);
#pragma unroll
    for (int k = 0; k < 5; k++) m[k] = swap_nibbles(m[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t mh = __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(hi4 - 32 * k, 0));
        const uint32_t ml = __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(lo4 - 32 * k, 0));
        v[k] = __funnelshift_r(m[k], m[k + 1], sh) & mh & ~ml;
    }
    if (ambiguous_nibbles(v[0]) | ambiguous_nibbles(v[1]) | ambiguous_nibbles(v[2]) | ambiguous_nibbles(v[3])) {   // ambiguity codes: rare
        v[0] = keep_acgt(v[0]); v[1] = keep_acgt(v[1]); v[2] = keep_acgt(v[2]); v[3] = keep_acgt(v[3]);
    }
}
// The same with the position masks from a table in shared memory: s_low[n] = the n lowest nibbles of a quad (n = 0 .. 32).
// k_project is bound by the integer ALU pipe (ncu: 71 % of its cycles), the two LDS.128 replace 24 ALU operations.
template <bool PREFETCH>
__device__ __forceinline__ void quad_piece_t(const ProjRec &R, int Q0, int a, int b, int dq, const uint4 *__restrict__ s_low, uint32_t (&v)[4]) {
    const uint4 ml = s_low[max(a - Q0, 0)], mh = s_low[min(b - Q0, 32)];
    const int n = Q0 + dq + R.nphase;
    const uint32_t *src = R.base4 + (n >> 3);
    const uint32_t sh = (uint32_t)(n & 7) * 4;
    uint32_t m[5];
#pragma unroll
    for (int k = 0; k < 5; k++) m[k] = __ldg(src + k);
    // the same lane cuts the quad 1024 positions on in the next round of the row loop: its SEQ window sits 512 bytes on
    // (give or take the indels in between); pull that sector into L1 now -- waiting for these loads from L2 is the largest
    // single stall of the kernel (ncu source view)
    if (PREFETCH && src + 128 + 8 <= R.lim) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + 128 + 2));
#pragma unroll
    for (int k = 0; k < 5; k++) m[k] = swap_nibbles(m[k]);
    v[0] = __funnelshift_r(m[0], m[1], sh) & mh.x & ~ml.x;
    v[1] = __funnelshift_r(m[1], m[2], sh) & mh.y & ~ml.y;
    v[2] = __funnelshift_r(m[2], m[3], sh) & mh.z & ~ml.z;
    v[3] = __funnelshift_r(m[3], m[4], sh) & mh.w & ~ml.w;
    if (ambiguous_nibbles(v[0]) | ambiguous_nibbles(v[1]) | ambiguous_nibbles(v[2]) | ambiguous_nibbles(v[3])) {   // ambiguity codes: rare
        v[0] = keep_acgt(v[0]); v[1] = keep_acgt(v[1]); v[2] = keep_acgt(v[2]); v[3] = keep_acgt(v[3]);
    }
}
__device__ __forceinline__ void fill_low_nibble_masks(uint4 *s_low) {        // threads 0 .. 32 of the CTA; barrier by the caller
    const int n = threadIdx.x;
    if (n <= 32) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = __funnelshift_lc(0xFFFFFFFFu, 0u, (uint32_t)max(4 * n - 32 * k, 0));
        s_low[n] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// quads [q_from, q_to) of the record from the segment list (see above)
__device__ __forceinline__ void emit_quads(const int *__restrict__ s_rs, const int *__restrict__ s_re,
                                           const int *__restrict__ s_dq, int n_seg, int q_from, int q_to,
                                           const ProjRec &R, int lane, const uint4 *__restrict__ s_low) {
    if (q_from >= q_to) return;
    auto quad_pos = [&](int q) { return (R.W0 + 4 * q) << 3; };
    // ---- phase 1
    // The lanes hold the ends of segments row_sp .. row_sp + 31 (a window over the list; every segment before it ends at or
    // before the row).  The window is moved only when it may not reach the end of the row: at one segment per ~150
    // positions it serves about four rows of 1024 positions.
    int row_sp = 0;
    int re_l = lane < n_seg ? s_re[lane] : 0x7fffffff;
    for (int q_row = q_from; q_row < q_to; q_row += 32) {
        const int Qr = quad_pos(q_row);
        if (__shfl_sync(0xffffffffu, re_l, 31) <= Qr + 31 * 32) {
            for (;;) {                                     // to the first segment with re > first position of the row
                const int adv = __popc(__ballot_sync(0xffffffffu, re_l <= Qr));
                if (adv == 0) break;
                row_sp += adv;
                re_l = row_sp + lane < n_seg ? s_re[row_sp + lane] : 0x7fffffff;
            }
        }
        const int q = q_row + lane;
        const int Q0 = Qr + 32 * lane;
        // owner = first segment with re > Q0: binary search over the 32 ends held by the lanes
        int idx = 0;
#pragma unroll
        for (int step = 16; step > 0; step >>= 1) {
            const int t = __shfl_sync(0xffffffffu, re_l, idx + step - 1);
            if (t <= Q0) idx += step;
        }
        int s = row_sp + idx;
        const int re_31 = __shfl_sync(0xffffffffu, re_l, 31);
        if (idx == 31 && re_31 <= Q0) {      // more than 32 segments in the row: rare
            int lo = s + 1, hi = n_seg;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_re[mid] <= Q0) lo = mid + 1; else hi = mid;
            }
            s = lo;
        }
        if (q < q_to) {
            const bool has = s < n_seg;
            const int a = has ? s_rs[s] : 0x7fffffff, b = has ? s_re[s] : 0x7fffffff, dq = has ? s_dq[s] : 0;
            uint32_t v[4] = {0u, 0u, 0u, 0u};
            if (a < Q0 + 32) quad_piece_t<true>(R, Q0, a, b, dq, s_low, v);
            *reinterpret_cast<uint4 *>(R.out + 4 * q) = make_uint4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncwarp();
    // ---- phase 2
    for (int s = 1 + lane; s < n_seg; s += 32) {
        const int a = s_rs[s];
        const int q = ((a >> 3) - R.W0) >> 2;
        if (q < q_from || q >= q_to) continue;
        const int Q0 = quad_pos(q);
        if (s_re[s - 1] <= Q0) continue;                   // no earlier segment in this quad: s owns it
        uint32_t v[4];
        quad_piece_t<false>(R, Q0, a, s_re[s], s_dq[s], s_low, v);
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (v[k]) atomicOr(R.out + 4 * q + k, v[k]);
    }
}

// sum of a 64-bit quantity (< 2^48) over the warp with three 16-bit-limb REDUX adds
__device__ __forceinline__ long long warp_sum48(long long v) {
    unsigned long long u = (unsigned long long)v;
    unsigned a = __reduce_add_sync(0xffffffffu, (unsigned)(u & 0xFFFFu));
    unsigned b = __reduce_add_sync(0xffffffffu, (unsigned)((u >> 16) & 0xFFFFu));
    unsigned c = __reduce_add_sync(0xffffffffu, (unsigned)((u >> 32) & 0xFFFFu));
    return (long long)((unsigned long long)a + ((unsigned long long)b << 16) + ((unsigned long long)c << 32));
}

// One warp per record.
// Pass 1 (every record): one sweep over the CIGAR for the filter of phasing.py:63-75 (total
// and soft-clipped length, IEEE double like the reference) and the reference span (M/=/X/D
// advance the reference; N/H/P advance nothing, the quirk of phasing.py:77-96); record
// validation; space for the projection is claimed with one atomic add.
// Pass 2 (accepted records): walks the CIGAR 32 ops at a time (prefix positions by warp
// scans), turns runs of adjacent M/=/X ops into match segments and writes the read
// in REFERENCE coordinates, aligned to the global 8-position grid:
// proj[r_woff[r] + w] holds positions (fuz_row_w0(gstart) + w) * 8 .. +7 as 4-bit codes: A=1 C=2
// G=4 T=8, 0 where the read shows no A/C/G/T (outside the alignment, deletions, N, ...).
// Closed segments are appended to the warp's shared-memory list; the list is turned into
// quads of the projection (emit_quads) when it fills up and at the end of the CIGAR.
#ifndef FUZ_PROJ_MINB
#define FUZ_PROJ_MINB 5             // CTAs per SM: 48 registers; at 6 (40 registers) the spills cost 8 % (measured), 4 is no faster
#endif
__global__ void __launch_bounds__(256, FUZ_PROJ_MINB) k_project(
    const uint8_t *__restrict__ rec_buf, const int64_t *__restrict__ rec_off, int n_rec,
    const int32_t *__restrict__ ctg_rec_off, const int64_t *__restrict__ ctg_goff, int n_ctg, HetScratch S, fuz_status *st) {
    fuz_pdl_enter();
    __shared__ int segs[8][3][FUZ_SEGCAP];
    __shared__ uint4 s_low[33];
    fill_low_nibble_masks(s_low);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int *s_rs = segs[threadIdx.x >> 5][0], *s_re = segs[threadIdx.x >> 5][1], *s_dq = segs[threadIdx.x >> 5][2];
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    long long acc_aligned = 0, acc_accepted = 0;
    // records are handed out dynamically (their cost varies with length and CIGAR): the first
    // one per warp statically, the rest from a global cursor
    int r_next = 0;
    for (int r = warp_g; r < n_rec; r = r_next) {
        if (lane == 0) r_next = n_warps + atomicAdd(S.rec_cursor, 1);
        r_next = __shfl_sync(0xffffffffu, r_next, 0);
        // ---- pass 1: issue every independent load first: offsets, then this and the previous header
        const int64_t off_r = rec_off[r], off_n = rec_off[r + 1], off_p = r > 0 ? rec_off[r - 1] : 0;
        const int64_t off_x = rec_off[min(r_next, n_rec - 1)];      // the warp's next record: header + CIGAR into L2
        const uint8_t *rec = rec_buf + off_r;
        const int32_t block_size = (int32_t)fuz_ld_u32_un(rec);
        const int32_t pos = (int32_t)fuz_ld_u32_un(rec + 8);
        const uint32_t w12 = fuz_ld_u32_un(rec + 12);     // l_read_name, mapq, bin
        const uint32_t w16 = fuz_ld_u32_un(rec + 16);     // n_cigar_op, flag
        const int32_t l_seq = (int32_t)fuz_ld_u32_un(rec + 20);
        const int32_t prev_pos = r > 0 ? (int32_t)fuz_ld_u32_un(rec_buf + off_p + 8) : 0;
        // contig of the record = position of r in ctg_rec_off (records grouped by contig)
        const int c = fuz_upper_bound(ctg_rec_off, 0, n_ctg + 1, r) - 1;
        const int l_name = w12 & 0xFF;
        const int n_cig = w16 & 0xFFFF;
        const uint8_t *cig = rec + 36 + l_name;
        const int64_t seq_off = off_r + 36 + l_name + 4 * (int64_t)n_cig;
        if (lane == 0) {
            S.r_flags[r] = 0; S.r_nwords[r] = 0; S.r_gstart[r] = 0; S.r_gend[r] = 0; S.r_woff[r] = 0;
        }
        if (c < 0 || c >= n_ctg || pos < 0 || l_seq < 0 || off_n - off_r != (int64_t)block_size + 4 ||
            36 + l_name + 4 * (int64_t)n_cig + ((int64_t)l_seq + 1) / 2 + l_seq > (int64_t)block_size + 4) {
            if (lane == 0) fuz_raise(st, FUZ_E_BADRECORD, r);
            continue;
        }
        // SEQ is needed after the CIGAR has been walked: start moving it to L2 now
        for (int64_t o = seq_off + 128 * lane; o < seq_off + ((l_seq + 1) >> 1); o += 128 * 32)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rec_buf + o));
        if (lane < 8 && r_next < n_rec) asm volatile("prefetch.global.L2 [%0];" ::"l"(rec_buf + off_x + 128 * lane));
        const int64_t gstart64 = ctg_goff[c] + pos;
        // coordinate order inside the contig
        if (lane == 0 && r > ctg_rec_off[c] && prev_pos > pos) fuz_raise(st, FUZ_E_UNSORTED, r);
        long long total = 0, skip = 0, aligned = 0, span = 0;
        bool badop = false;
#pragma unroll 4
        for (int k = lane; k < n_cig; k += 32) {
            uint32_t cw = fuz_ld_u32_un(cig + 4 * k);
            uint32_t len = cw >> 4, op = cw & 15;
            if (op > 8) badop = true;
            total += len;
            if (op == 4) skip += len;
            if (op_is_match(op)) { aligned += len; span += len; }
            if (op == 2) span += len;
        }
        total = warp_sum48(total); skip = warp_sum48(skip); aligned = warp_sum48(aligned); span = warp_sum48(span);
        badop = __any_sync(0xffffffffu, badop);
        if (badop || total == 0 || gstart64 + span > 0x7fffffffLL) {   // unknown op / ZeroDivisionError :72
            if (lane == 0) fuz_raise(st, FUZ_E_BADRECORD, r);
            continue;
        }
        // phasing.py:72-75 in IEEE double, same operation order as the reference
        // (skip == 0: 1.0 - 0.0 < 0.1 is false, no division needed)
        const bool accept = (skip == 0 || !(1.0 - 1.0 * (double)skip / (double)total < 0.1)) && !(total < 2000);
        const int gstart = (int)gstart64;
        // words of the projection, padded to whole quads (128-bit flushes)
        const int n_words = (accept && span > 0) ? (int)((((gstart64 + span - 1) >> 3) - ((gstart64 >> 3) & ~3LL) + 4) & ~3LL) : 0;
        int woff = 0;
        if (lane == 0) {
            S.r_gstart[r] = gstart;
            S.r_gend[r] = accept ? gstart + (int32_t)span : gstart;
            S.r_flags[r] = accept ? 1 : 0;
            S.r_nwords[r] = n_words;
            if (accept) {
                atomicMax(&S.ctg_last_rec[c], r);
                atomicMax(&S.ctg_maxspan[c], (int32_t)span);
                acc_aligned += aligned; acc_accepted += 1;
            }
            if (n_words) {
                long long o = (long long)atomicAdd((unsigned long long *)&st->reserved[0], (unsigned long long)n_words);
                if (o + n_words > S.proj_cap) { fuz_raise(st, FUZ_E_CAPACITY, 6); o = -1; }
                woff = (int)o;
                S.r_woff[r] = woff < 0 ? 0 : woff;
            }
        }
        woff = __shfl_sync(0xffffffffu, woff, 0);
        if (n_words == 0 || woff < 0) continue;
        // ---- pass 2: projection
        ProjRec R;
        {
            const uintptr_t sa = reinterpret_cast<uintptr_t>(rec_buf + seq_off);
            R.base4 = reinterpret_cast<const uint32_t *>(sa & ~(uintptr_t)3);
            R.nphase = (int)(sa & 3) * 2;
            R.W0 = fuz_row_w0(gstart);
            R.out = S.proj + woff;
            R.lim = reinterpret_cast<const uint32_t *>((reinterpret_cast<uintptr_t>(rec_buf + off_n)) & ~(uintptr_t)3);
        }
        int n_seg = 0, q_next = 0;
        int carry_rp = gstart, carry_qp = 0;
        bool overrun = false;
        for (int k0 = 0; k0 < n_cig; k0 += 32) {
            const int k = k0 + lane;
            const bool valid = k < n_cig;
            const uint32_t cw = valid ? fuz_ld_u32_un(cig + 4 * k) : 0u;
            const int len = (int)(cw >> 4);
            const uint32_t op = cw & 15;
            const bool mt = valid && op_is_match(op);                                  // M = X
            const bool is_m = mt && len > 0;
            const int radv = (mt || (valid && op == 2)) ? len : 0;                      // M = X D
            const int qadv = (mt || (valid && (op == 1 || op == 4))) ? len : 0;         // M = X I S
            const int rinc = fuz_warp_incl_scan(radv, lane);
            const int qinc = fuz_warp_incl_scan(qadv, lane);
            const int rp0 = carry_rp + rinc - radv;
            const int qp0 = carry_qp + qinc - qadv;
            if (is_m && (long long)qp0 + len > (long long)l_seq) overrun = true;       // IndexError :84
            const uint32_t mmask = __ballot_sync(0xffffffffu, is_m);
            if (__any_sync(0xffffffffu, overrun)) break;     // bad record: stop before reading past SEQ
            // runs of adjacent match ops inside the chunk become one segment each (a run that
            // crosses the chunk border, or is interrupted by N/H/P, gives touching segments:
            // same projection); appended in CIGAR order = reference order
            const uint32_t smask = mmask & ~(mmask << 1), emask = mmask & ~(mmask >> 1);
            const int e = lane + __ffs(emask >> lane) - 1;                             // last op of the run starting here
            const int run_end = __shfl_sync(0xffffffffu, rp0 + len, e & 31);
            if ((smask >> lane) & 1u) {
                const int idx = n_seg + __popc(smask & lt);
                s_rs[idx] = rp0; s_re[idx] = run_end; s_dq[idx] = qp0 - rp0;
            }
            n_seg += __popc(smask);
            __syncwarp();
            if (n_seg > FUZ_SEGCAP - 32) {
                // emit every quad that no later segment can touch: those before the quad holding
                // the end of the last segment; keep the (<= 32) segments that reach into that quad
                const int q_lim = (((s_re[n_seg - 1] - 1) >> 3) - R.W0) >> 2;
                emit_quads(s_rs, s_re, s_dq, n_seg, q_next, q_lim, R, lane, s_low);
                q_next = max(q_next, q_lim);
                const int P = (R.W0 + 4 * q_lim) << 3;
                const int s = n_seg - 32 + lane;
                const bool keep = s >= 0 && s_re[s] > P;
                const uint32_t kmask = __ballot_sync(0xffffffffu, keep);
                int t_rs = 0, t_re = 0, t_dq = 0;
                if (keep) { t_rs = s_rs[s]; t_re = s_re[s]; t_dq = s_dq[s]; }
                __syncwarp();
                if (keep) { const int d = __popc(kmask & lt); s_rs[d] = t_rs; s_re[d] = t_re; s_dq[d] = t_dq; }
                n_seg = __popc(kmask);
                __syncwarp();
            }
            carry_rp += __shfl_sync(0xffffffffu, rinc, 31);
            carry_qp += __shfl_sync(0xffffffffu, qinc, 31);
        }
        overrun = __any_sync(0xffffffffu, overrun);
        if (!overrun) emit_quads(s_rs, s_re, s_dq, n_seg, q_next, n_words >> 2, R, lane, s_low);
        __syncwarp();
        if (overrun && lane == 0) fuz_raise(st, FUZ_E_BADRECORD, r);
    }
    if (lane == 0 && acc_accepted) {
        atomicAdd((unsigned long long *)&st->aligned_bases, (unsigned long long)acc_aligned);
        atomicAdd((unsigned long long *)&st->n_accepted, (unsigned long long)acc_accepted);
    }
}

// ---------------------------------------------------------------- tiled pileup (default)
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#define FUZ_FA(a, b, c, s, cy) do { s = xor3(a, b, c); cy = maj3(a, b, c); } while (0)

// count of base b at position i from the 8 bit planes of the vertical counters
__device__ __forceinline__ uint32_t plane_count(const uint32_t (&acc)[8], int bit) {
    uint32_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) v |= ((acc[k] >> bit) & 1u) << k;
    return v;
}

// Bit-sliced screen before the exact het test: which of a thread's 8 positions CAN be het sites.  Bit 4i+b of plane k is
// bit k of the count of base b at position i.  A het site has total >= 10, hence a second base with count >= 3, and
// 4*c1 > total >= c0 + c1, hence c1 > c0/3 >= 2^h0 / 3 with h0 the top plane of the position: the second base has a bit in
// plane h0-2 or above.  Positions whose nibble of the result is zero are skipped (one in ~10^4 survives at 60x; the exact
// test, phasing.py:112-120, runs on the survivors only).
__device__ __forceinline__ uint32_t nib_spread(uint32_t x) {        // 0xF in every nibble that has a bit set
    x |= x >> 1; x |= x >> 2;
    return (x & 0x11111111u) * 15u;
}
__device__ __forceinline__ uint32_t het_candidates(const uint32_t (&acc)[8]) {
    const uint32_t ge3 = (acc[0] & acc[1]) | acc[2] | acc[3] | acc[4] | acc[5] | acc[6] | acc[7];
    const uint32_t p7 = acc[7], p6 = p7 | acc[6], p5 = p6 | acc[5], p4 = p5 | acc[4], p3 = p4 | acc[3];
    const uint32_t near_top = acc[7] | acc[6] | acc[5] | (acc[4] & ~nib_spread(p7)) | (acc[3] & ~nib_spread(p6)) |
                              (acc[2] & ~nib_spread(p5)) | (acc[1] & ~nib_spread(p4)) | (acc[0] & ~nib_spread(p3));
    const uint32_t q = ge3 & near_top;
    return (q & (q >> 1) & 0x77777777u) | (q & (q >> 2) & 0x33333333u) | (q & (q >> 3) & 0x11111111u);
}

// One CTA per 2048-position tile, one thread per 8-position word.  Every read overlapping
// the tile contributes one ALIGNED word load per thread (no shifting: the projection is on
// the global grid, one-hot A/C/G/T nibbles).  Counting is bit-sliced *vertically*: 15 words
// are reduced by a carry-save adder tree (11 full adders = 22 LOP3) to a 4-bit number per
// (position, base) bit, which is rippled into 8 bit planes held in registers (depth <= 255
// between spills into 16-bit counters).  No atomics, no shared-memory histogram.
__global__ void __launch_bounds__(FUZ_PTILE_THREADS, 5) k_pileup_gather(HetScratch S, int64_t cap_sites,
                                                                       uint32_t *__restrict__ counts_out, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    __shared__ int4 l_ent[FUZ_PTILE_THREADS];             // x = word offset of the projection, y = first word, z = words
    __shared__ uint32_t c16s[16][FUZ_PTILE_THREADS];      // spill counters (depth > 255 only): [4 * base + j][thread]
    __shared__ int s_warp_tot[FUZ_NW];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int t0 = tile * FUZ_PTILE, t1 = t0 + FUZ_PTILE;
    const int Wt = (t0 >> 3) + tid;                       // my word on the global grid
    const int rlo = S.tile_rlo[tile], rhi = S.tile_rhi[tile];
    const uint32_t *__restrict__ proj = S.proj;
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};           // bit planes of the per-(position, base) counters
    // spill: 16-bit counters, base b, positions 2j (low half) / 2j+1 (high half); kept in shared
    // memory because they are touched only when more than 255 reads cover a tile
#define C16(b, j) c16s[4 * (b) + (j)][tid]
    int n_reads_seen = 0, groups_in_acc = 0;
    bool spilled = false;
    for (int cb = rlo; cb < rhi; cb += FUZ_PTILE_THREADS) {
        // compact the records overlapping the tile (any order: only counts matter here)
        const int r = cb + tid;
        const bool ok = r < rhi && S.r_flags[r] && S.r_gend[r] > t0 && S.r_gstart[r] < t1;
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp_tot[warp] = __popc(m);
        __syncthreads();
        int off = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < FUZ_NW; w++) { int v = s_warp_tot[w]; if (w < warp) off += v; tot += v; }
        if (ok) l_ent[off + __popc(m & ((1u << lane) - 1u))] = make_int4(S.r_woff[r], fuz_row_w0(S.r_gstart[r]), S.r_nwords[r], 0);
        __syncthreads();
        n_reads_seen += tot;
        for (int g = 0; g < tot; g += 15) {
            const int ns = min(15, tot - g);
            uint32_t x[15];
#pragma unroll
            for (int s = 0; s < 15; s++) {
                x[s] = 0;
                if (s < ns) {
                    const int4 e = l_ent[g + s];
                    const int j = Wt - e.y;
                    if ((unsigned)j < (unsigned)e.z) x[s] = __ldg(proj + e.x + j);
                }
            }
            // carry-save adder tree: 15 one-bit inputs per bit position -> 4-bit count
            uint32_t s0, s1, s2, s3, s4, s5, k0, k1, k2, k3, k4, k5, k6, ones, t0_, t1_, d0, d1, d2, twos, fours, eights;
            FUZ_FA(x[0], x[1], x[2], s0, k0); FUZ_FA(x[3], x[4], x[5], s1, k1); FUZ_FA(x[6], x[7], x[8], s2, k2);
            FUZ_FA(x[9], x[10], x[11], s3, k3); FUZ_FA(x[12], x[13], x[14], s4, k4);
            FUZ_FA(s0, s1, s2, s5, k5); FUZ_FA(s3, s4, s5, ones, k6);
            FUZ_FA(k0, k1, k2, t0_, d0); FUZ_FA(k3, k4, k5, t1_, d1); FUZ_FA(t0_, t1_, k6, twos, d2);
            FUZ_FA(d0, d1, d2, fours, eights);
            // ripple the 4-bit number into the 8 planes
            uint32_t cy = acc[0] & ones; acc[0] ^= ones;
            uint32_t nc = maj3(acc[1], twos, cy); acc[1] = xor3(acc[1], twos, cy); cy = nc;
            nc = maj3(acc[2], fours, cy); acc[2] = xor3(acc[2], fours, cy); cy = nc;
            nc = maj3(acc[3], eights, cy); acc[3] = xor3(acc[3], eights, cy); cy = nc;
#pragma unroll
            for (int k = 4; k < 8; k++) { nc = acc[k] & cy; acc[k] ^= cy; cy = nc; }
            if (++groups_in_acc == 17) {                 // 17 * 15 = 255: planes are full, spill to 16-bit counters
#pragma unroll
                for (int b = 0; b < 4; b++)
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (!spilled && (i & 1) == 0) C16(b, i >> 1) = 0;
                        C16(b, i >> 1) += plane_count(acc, 4 * i + b) << ((i & 1) * 16);
                    }
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] = 0;
                groups_in_acc = 0;
                spilled = true;
            }
        }
        __syncthreads();
    }
    if (n_reads_seen > 65535 && tid == 0) fuz_raise(st, FUZ_E_DEPTH, tile);
    const int pos_limit = S.tile_limit[tile];
    uint32_t cnt[8][4];
    uint32_t hetmask = 0;
    if (!spilled && !counts_out) {
        // a het site needs two bases with count >= 3 (second allele > 25 % of a depth >= 10):
        // test that on the planes and extract counts only for the few candidate positions
        const uint32_t two = het_candidates(acc);
#pragma unroll
        for (int i = 0; i < 8; i++) {
#pragma unroll
            for (int b = 0; b < 4; b++) cnt[i][b] = 0;
            if ((two >> (4 * i)) & 7u) {
#pragma unroll
                for (int b = 0; b < 4; b++) cnt[i][b] = plane_count(acc, 4 * i + b);
                if (t0 + tid * 8 + i < pos_limit && het_test(cnt[i][0], cnt[i][1], cnt[i][2], cnt[i][3])) hetmask |= 1u << i;
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                cnt[i][b] = (spilled ? (C16(b, i >> 1) >> ((i & 1) * 16)) & 0xFFFFu : 0u) + plane_count(acc, 4 * i + b);
        if (counts_out) {
            uint4 *o = reinterpret_cast<uint4 *>(counts_out) + (size_t)t0 + (size_t)tid * 8;
#pragma unroll
            for (int i = 0; i < 8; i++) o[i] = make_uint4(cnt[i][0], cnt[i][1], cnt[i][2], cnt[i][3]);
        }
        hetmask = het_mask_of(cnt, t0, pos_limit);
    }
    emit_tile_sites(hetmask, cnt, tile, t0, S, cap_sites, st, s_warp_tot, &s_base);
}

#undef C16

#include "fuz_pileup_seg.cuh"

// ---------------------------------------------------------------- cross-check pileup (impl 1)
// One warp per accepted record walks its CIGAR and adds every M/=/X base with a global
// atomic.  Slow by design (L2 atomics); kept as an independent device implementation that
// the tests compare against the tiled kernel.
__global__ void __launch_bounds__(256) k_pileup_atomic(
    const uint8_t *__restrict__ rec_buf, const int64_t *__restrict__ rec_off, int n_rec,
    const int32_t *__restrict__ ctg_rec_off, const int64_t *__restrict__ ctg_goff, int n_ctg, HetScratch S,
    uint32_t *__restrict__ counts, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp_g; r < n_rec; r += n_warps) {
        if (!S.r_flags[r]) continue;
        const uint8_t *rec = rec_buf + rec_off[r];
        int c = fuz_upper_bound(ctg_rec_off, 0, n_ctg + 1, r) - 1;
        const int64_t gend = ctg_goff[c + 1];
        const int l_name = fuz_ld_u32_un(rec + 12) & 0xFF;
        const int n_cig = fuz_ld_u32_un(rec + 16) & 0xFFFF;
        const uint8_t *cig = rec + 36 + l_name;
        const uint8_t *seq = cig + 4 * (int64_t)n_cig;
        int64_t rp = S.r_gstart[r];
        int qp = 0;
        for (int k = 0; k < n_cig; k++) {
            uint32_t cw = fuz_ld_u32_un(cig + 4 * k);
            int len = (int)(cw >> 4);
            uint32_t op = cw & 15;
            if (op == 4 || op == 1) qp += len;
            else if (op == 2) rp += len;
            else if (op_is_match(op)) {
                for (int i = lane; i < len; i += 32) {
                    int q = qp + i;
                    uint32_t byte = seq[q >> 1];
                    uint32_t nib = (q & 1) ? (byte & 15u) : (byte >> 4);
                    int b = nib == 1 ? 0 : nib == 2 ? 1 : nib == 4 ? 2 : nib == 8 ? 3 : -1;
                    int64_t p = rp + i;
                    if (b >= 0 && p < gend) atomicAdd(&counts[4 * p + b], 1u);
                }
                rp += len; qp += len;
            }
        }
    }
}

__global__ void __launch_bounds__(FUZ_PTILE_THREADS) k_het_from_counts(
    const uint32_t *__restrict__ counts, HetScratch S, int64_t cap_sites, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    __shared__ int s_warp_tot[FUZ_NW];
    __shared__ int s_base;
    const int tile = blockIdx.x, tid = threadIdx.x;
    const int t0 = tile * FUZ_PTILE;
    uint32_t cnt[8][4];
    const uint4 *in = reinterpret_cast<const uint4 *>(counts) + (size_t)t0 + (size_t)tid * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint4 v = in[i];
        cnt[i][0] = v.x; cnt[i][1] = v.y; cnt[i][2] = v.z; cnt[i][3] = v.w;
    }
    emit_tile_sites(het_mask_of(cnt, t0, S.tile_limit[tile]), cnt, tile, t0, S, cap_sites, st, s_warp_tot, &s_base);
}

// ---------------------------------------------------------------- ordered sites + rows
// Three steps: exclusive scan of the per-tile site counts (tiles are in position order, so this puts the sites in
// order; fuz_scan_i32 publishes the site count), k_sites_place copies the sites to their ordered slots with the derived
// fields (one thread per site over the whole grid: at 60 000 sites a batch one CTA spent 50 us here on 15 dependent
// rounds), then the exclusive scan of the variant_map rows per site.
__global__ void __launch_bounds__(256) k_sites_place(HetScratch S, const int64_t *__restrict__ ctg_goff, fuz_outputs O, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int total = (int)st->n_sites;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < total; u += gridDim.x * blockDim.x) {
        // claimed slot u of its tile -> ordered slot dd
        const int t = S.us_tile[u];
        const int gp = S.us_gpos[u];
        const uint4 v = reinterpret_cast<const uint4 *>(S.us_cnt)[u];
        const int dd = S.tile_site_off[t] + (u - S.tile_site_base[t]);
        const int c = S.tile_ctg[t];
        const int64_t org = ctg_goff[c];
        const uint32_t k[4] = {v.x << 2, (v.y << 2) | 1, (v.z << 2) | 2, (v.w << 2) | 3};
        uint32_t m0 = max(max(k[0], k[1]), max(k[2], k[3])), m1 = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) if (k[b] != m0) m1 = max(m1, k[b]);
        const int b0 = m0 & 3, b1 = m1 & 3;
        S.s_gpos[dd] = gp;
        O.d_site_ctg[dd] = c;
        O.d_site_pos[dd] = (int32_t)(gp - org) + 1;
        reinterpret_cast<uint4 *>(O.d_site_cnt)[dd] = v;
        // allele order of the association table = order by "ACTG" (SURVEY.md B.1):
        // rank A=0 C=1 T=2 G=3 (two bits each in 0xB4)
        const bool sw = ((0xB4 >> (2 * b0)) & 3) > ((0xB4 >> (2 * b1)) & 3);
        reinterpret_cast<uchar2 *>(O.d_site_top)[dd] = make_uchar2((uint8_t)b0, (uint8_t)b1);
        reinterpret_cast<uchar2 *>(O.d_site_al)[dd] = make_uchar2((uint8_t)(sw ? b1 : b0), (uint8_t)(sw ? b0 : b1));
        S.site_rows[dd] = (int)((m0 >> 2) + (m1 >> 2));
    }
}

// One warp per site: the records covering the site, in record (= file) order, 32 at a
// time; ballot/popc turn "my read carries the major / minor allele" into ordered row
// slots (phasing.py:125-128: all major-allele reads, then all minor-allele reads).
__global__ void __launch_bounds__(256) k_signature(
    const int32_t *__restrict__ rec_qid, const int32_t *__restrict__ ctg_rec_off, HetScratch S, fuz_outputs O,
    fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    const int n_sites = (int)st->n_sites;
    const uint32_t lt = (1u << lane) - 1u;
    for (int s = warp_g; s < n_sites; s += n_warps) {
        const int gp = S.s_gpos[s];
        const int b0 = O.d_site_top[2 * s], b1 = O.d_site_top[2 * s + 1];
        const uint32_t code0 = 1u << b0, code1 = 1u << b1;
        const int n0 = O.d_site_cnt[4 * s + b0], n1 = O.d_site_cnt[4 * s + b1];
        // records that can cover the site = candidates of its pileup tile (file order)
        const int rlo = S.tile_rlo[gp / FUZ_PTILE], rhi = S.tile_rhi[gp / FUZ_PTILE];
        const int64_t off0 = S.site_row_off[s], off1 = off0 + n0;
        int run0 = 0, run1 = 0;
        // the four record fields of a round are independent loads (no short-circuit chain), and those of the NEXT round are
        // requested before this round's projection word is used: two dependent load levels per round instead of five
        int fl = 0, ge = 0, gs = 0, wo = 0;
        if (rlo + lane < rhi) { const int r0 = rlo + lane; fl = S.r_flags[r0]; ge = S.r_gend[r0]; gs = S.r_gstart[r0]; wo = S.r_woff[r0]; }
        for (int rb = rlo; rb < rhi; rb += 32) {
            const int r = rb + lane;
            const bool cov = fl && ge > gp && gs <= gp;
            uint32_t w = 0;
            if (cov) w = S.proj[wo + ((gp >> 3) - fuz_row_w0(gs))];
            fl = 0;
            if (r + 32 < rhi) { const int rn = r + 32; fl = S.r_flags[rn]; ge = S.r_gend[rn]; gs = S.r_gstart[rn]; wo = S.r_woff[rn]; }
            const uint32_t nib = cov ? (w >> (4 * (gp & 7))) & 15u : 0u;
            uint32_t m0 = __ballot_sync(0xffffffffu, nib == code0);
            uint32_t m1 = __ballot_sync(0xffffffffu, nib == code1);
            if (nib == code0 || nib == code1) {
                bool first = nib == code0;
                int64_t i = first ? off0 + run0 + __popc(m0 & lt) : off1 + run1 + __popc(m1 & lt);
                if (i < O.cap_vmap) {
                    O.d_vm_site[i] = s;
                    O.d_vm_base[i] = (uint8_t)(first ? b0 : b1);
                    O.d_vm_qid[i] = rec_qid[r];
                }
            }
            run0 += __popc(m0); run1 += __popc(m1);
        }
        if (lane == 0 && (run0 != n0 || run1 != n1)) fuz_raise(st, FUZ_E_INTERNAL, s);
    }
}

}  // namespace

// ------------------------------------------------------------------ host side
int fuz_het_call_impl(fuz_ctx *ctx, const fuz_batch *in, fuz_outputs *out) {
    if (!ctx || !in || !out) return FUZ_E_ARG;
    if (in->n_ctg < 1 || in->n_rec < 0) return fuz_fail(ctx, FUZ_E_ARG, "fuz_het_call: empty batch");
    if (in->total_glen <= 0 || in->total_glen % FUZ_TILE || in->total_glen > 0x7fff0000LL)
        return fuz_fail(ctx, FUZ_E_ARG, "total_glen %lld must be a positive multiple of %d below 2^31",
                        (long long)in->total_glen, FUZ_TILE);
    cudaStream_t st = ctx->stream;
    const int n_rec = in->n_rec, n_ctg = in->n_ctg;
    const bool seg_path = ctx->pileup_impl == 2, seg_proj = ctx->pileup_impl == 3;     // 3: k_segments + k_project_seg + register pileup
    // tiles of the path that runs: 8192 positions (segment pileup) or 2048 (projection pileup, cross-check het test)
    const int n_tiles = (int)(in->total_glen / (seg_path ? FUZ_TILE : FUZ_PTILE));
    const int64_t cap_sites = out->cap_sites;
    // projection paths (pileup_impl 0, 1): one word per 8 aligned reference positions.  SEQ holds 2 bases per byte, so
    // rec_bytes / 4 words cover every M/=/X base; deletions add to the span and are checked on the device
    // (FUZ_E_CAPACITY, index 6).
    const int64_t proj_cap = seg_path ? 0 : in->rec_bytes / 4 + 8 * (int64_t)n_rec + 64;
    if (proj_cap > 0x7fffffffLL) return fuz_fail(ctx, FUZ_E_ARG, "batch too large: split it (projection exceeds 2^31 words)");
    // segment path (pileup_impl 2): a record with n CIGAR operations takes n / 2 + 1 segment slots; a read adds one
    // entry to every tile it overlaps.  Both are bounded by heuristics on the record bytes (a BAM record spends 1.5
    // bytes per base on SEQ + QUAL); a denser batch fails with FUZ_E_CAPACITY (index 6 / 9) and fuz_status says how
    // much is needed (n_segments, reserved[0]): options "seg_cap" / "ent_cap" raise the reservation.
    int64_t seg_cap = 0, ent_cap = 0;
    if (seg_proj) seg_cap = std::max<int64_t>(in->rec_bytes / 20 + 2 * (int64_t)n_rec + 64, ctx->seg_cap_min);
    if (seg_path) {
        seg_cap = std::max<int64_t>(in->rec_bytes / 20 + 2 * (int64_t)n_rec + 64, ctx->seg_cap_min);
        ent_cap = std::max<int64_t>(in->rec_bytes / (FUZ_TILE / 2) + 3 * (int64_t)n_rec + 64, ctx->ent_cap_min);
        if (seg_cap > 0x7fffffffLL || ent_cap > 0x7fffffffLL)
            return fuz_fail(ctx, FUZ_E_ARG, "batch too large: split it (more than 2^31 segments or tile entries)");
    }
    const int pile_ctas = seg_path ? std::min(n_tiles, 148 * 2) : 0;
    const int gather_ctas = (ctx->pileup_impl == 0 && ctx->gather_tma) ? std::min(n_tiles, 148 * 3) : 0;
    HetScratch S;
    FuzLayout L;
    size_t o_gstart = L.add(4 * (size_t)(n_rec + 1)), o_gend = L.add(4 * (size_t)(n_rec + 1));
    size_t o_nw = L.add(4 * (size_t)(n_rec + 2)), o_woff = L.add(4 * (size_t)(n_rec + 2));
    size_t o_rseq = L.add(8 * (size_t)(n_rec + 1)), o_flags = L.add((size_t)n_rec + 1);
    size_t o_segoff = L.add(4 * (size_t)(n_rec + 2)), o_nseg = L.add(4 * (size_t)(n_rec + 2));
    size_t o_proj = L.add(4 * (size_t)(proj_cap + 4));       // + 4: k_pileup_gather_tma rounds its copies up to 16 bytes
    size_t o_segs = L.add(16 * (size_t)seg_cap), o_ents = L.add(32 * (size_t)ent_cap);
    size_t o_clast = L.add(4 * (size_t)n_ctg), o_cspan = L.add(4 * (size_t)n_ctg);
    size_t o_tctg = L.add(4 * (size_t)n_tiles), o_tlo = L.add(4 * (size_t)n_tiles), o_thi = L.add(4 * (size_t)n_tiles);
    size_t o_tlim = L.add(4 * (size_t)n_tiles);
    size_t o_tbase = L.add(4 * (size_t)n_tiles), o_tcnt = L.add(4 * (size_t)(n_tiles + 1)),
           o_toff = L.add(4 * (size_t)(n_tiles + 2));
    size_t o_usg = L.add(4 * (size_t)cap_sites), o_usc = L.add(16 * (size_t)cap_sites), o_ust = L.add(4 * (size_t)cap_sites);
    size_t o_sg = L.add(4 * (size_t)cap_sites), o_srow = L.add(4 * (size_t)(cap_sites + 1)),
           o_sroff = L.add(4 * (size_t)(cap_sites + 2));
    size_t o_cursor = L.add(16);
    size_t o_spill = L.add(seg_path ? 4 * (size_t)pile_ctas * 64 * FUZ_CONSUMERS : 4 * (size_t)gather_ctas * 16 * FUZ_PTILE_THREADS);
    size_t o_counts = 0;
    const bool need_counts = ctx->pileup_impl == 1 && !out->d_counts;
    if (need_counts) o_counts = L.add(16 * (size_t)in->total_glen);
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    int32_t *keep_row_off; uint8_t *keep_dup;
    if ((rc = fuz_keep_commit(ctx, cap_sites, out->cap_vmap, &keep_row_off, &keep_dup))) return rc;
    S.r_gstart = fuz_at<int32_t>(ctx, o_gstart); S.r_gend = fuz_at<int32_t>(ctx, o_gend);
    S.r_nwords = fuz_at<int32_t>(ctx, o_nw); S.r_woff = fuz_at<int32_t>(ctx, o_woff);
    S.r_seq = fuz_at<int64_t>(ctx, o_rseq); S.r_flags = fuz_at<uint8_t>(ctx, o_flags);
    S.proj = fuz_at<uint32_t>(ctx, o_proj); S.proj_cap = proj_cap;
    S.segs = fuz_at<int4>(ctx, o_segs); S.seg_cap = seg_cap;
    S.ents = fuz_at<FuzTileEnt>(ctx, o_ents); S.ent_cap = ent_cap;
    S.r_seg_off = fuz_at<int32_t>(ctx, o_segoff); S.r_nseg = fuz_at<int32_t>(ctx, o_nseg);
    S.ctg_last_rec = fuz_at<int32_t>(ctx, o_clast); S.ctg_maxspan = fuz_at<int32_t>(ctx, o_cspan);
    S.rec_cursor = fuz_at<int32_t>(ctx, o_cursor); S.tile_cursor = S.rec_cursor + 1;
    S.tile_ctg = fuz_at<int32_t>(ctx, o_tctg); S.tile_rlo = fuz_at<int32_t>(ctx, o_tlo); S.tile_rhi = fuz_at<int32_t>(ctx, o_thi);
    S.tile_ent_base = S.tile_rlo; S.tile_ent_cnt = S.tile_rhi;
    S.tile_limit = fuz_at<int32_t>(ctx, o_tlim);
    S.tile_site_base = fuz_at<int32_t>(ctx, o_tbase); S.tile_site_cnt = fuz_at<int32_t>(ctx, o_tcnt);
    S.tile_site_off = fuz_at<int32_t>(ctx, o_toff);
    S.us_gpos = fuz_at<int32_t>(ctx, o_usg); S.us_cnt = fuz_at<uint32_t>(ctx, o_usc);
    S.us_tile = fuz_at<int32_t>(ctx, o_ust);
    S.s_gpos = fuz_at<int32_t>(ctx, o_sg); S.site_rows = fuz_at<int32_t>(ctx, o_srow);
    S.site_row_off = keep_row_off;       // also the first input of the association stage
    (void)o_sroff;
    S.spill = fuz_at<uint32_t>(ctx, o_spill);
    S.counts = out->d_counts ? out->d_counts : (need_counts ? fuz_at<uint32_t>(ctx, o_counts) : nullptr);
    S.n_tiles = n_tiles;

    fuz_launch(ctx, k_het_init, 1, 256, 0, st, ctx->d_status, S.ctg_last_rec, S.ctg_maxspan, S.rec_cursor, n_ctg);
    FUZ_LAUNCH_CHECK(ctx, "k_het_init");
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (ctx->timing) {
        if (ctx->timing_used == ctx->timing_events.size()) {
            cudaEvent_t a, b;
            FUZ_CUDA(ctx, cudaEventCreate(&a));
            FUZ_CUDA(ctx, cudaEventCreate(&b));
            ctx->timing_events.push_back({a, b});
        }
        ev0 = ctx->timing_events[ctx->timing_used].first;
        ev1 = ctx->timing_events[ctx->timing_used].second;
        ctx->timing_used++;
    }
    if (seg_path) {
        const size_t smem = FUZ_NSTAGE * sizeof(FuzStage) + FUZ_G * (FUZ_TILE / 8) * sizeof(uint32_t) + 2 * FUZ_NSTAGE * sizeof(uint64_t);
        if (!ctx->pileup_attr_set) {
            FUZ_CUDA(ctx, cudaFuncSetAttribute(k_pileup_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ctx->pileup_attr_set = true;
        }
        if (ev0) FUZ_CUDA(ctx, cudaEventRecord(ev0, st));
        if (ctx->trace) ctx->trace[0] = 1;
        if (n_rec > 0) {
            const int ctas = std::min((n_rec + 255) / 256, 148 * 8);
            fuz_launch(ctx, k_segments, ctas, 256, 0, st, in->d_rec_buf, in->d_rec_off, n_rec, in->d_ctg_rec_off, in->d_ctg_goff, n_ctg, S,
                       ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_segments");
        }
        fuz_launch(ctx, k_tile_lists, std::min((n_tiles + 7) / 8, 148 * 8), 256, 0, st, n_tiles, n_ctg, in->d_rec_buf, in->d_ctg_goff,
                   in->d_ctg_rec_off, S, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_tile_lists");
        fuz_launch(ctx, k_pileup_tma, pile_ctas, FUZ_PILEUP_THREADS, smem, st, in->d_rec_buf, S, cap_sites, out->d_counts, ctx->d_status,
                   ctx->pileup_debug, ctx->trace);
        FUZ_LAUNCH_CHECK(ctx, "k_pileup_tma");
        if (ev1) FUZ_CUDA(ctx, cudaEventRecord(ev1, st));
    } else if (seg_proj) {
        if (ev0) FUZ_CUDA(ctx, cudaEventRecord(ev0, st));
        if (n_rec > 0) {
            const int ctas = std::min((n_rec + 255) / 256, 148 * 8);
            fuz_launch(ctx, k_segments, ctas, 256, 0, st, in->d_rec_buf, in->d_rec_off, n_rec, in->d_ctg_rec_off, in->d_ctg_goff, n_ctg, S,
                       ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_segments");
            fuz_launch(ctx, k_project_seg, ctx->project_ctas, 256, 0, st, in->d_rec_buf, n_rec, S, ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_project_seg");
        }
        fuz_launch(ctx, k_tile_ranges, (n_tiles + 7) / 8, 256, 0, st, n_tiles, n_ctg, in->d_ctg_goff, in->d_ctg_rec_off, S, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_tile_ranges");
        fuz_launch(ctx, k_pileup_gather, n_tiles, FUZ_PTILE_THREADS, 0, st, S, cap_sites, out->d_counts, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_pileup_gather");
        if (ev1) FUZ_CUDA(ctx, cudaEventRecord(ev1, st));
    } else if (ctx->pileup_impl == 0) {
        if (ev0) FUZ_CUDA(ctx, cudaEventRecord(ev0, st));
        if (n_rec > 0) {
            fuz_launch(ctx, k_project, ctx->project_ctas, 256, 0, st, in->d_rec_buf, in->d_rec_off, n_rec, in->d_ctg_rec_off, in->d_ctg_goff,
                                                       n_ctg, S, ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_project");
        }
        fuz_launch(ctx, k_tile_ranges, (n_tiles + 7) / 8, 256, 0, st, n_tiles, n_ctg, in->d_ctg_goff, in->d_ctg_rec_off, S, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_tile_ranges");
        if (ctx->gather_tma) {
            const size_t gsm = FUZ_GSTAGES * sizeof(FuzGStage) + 2 * FUZ_GSTAGES * sizeof(uint64_t);
            if (!ctx->gather_attr_set) {
                FUZ_CUDA(ctx, cudaFuncSetAttribute(k_pileup_gather_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
                ctx->gather_attr_set = true;
            }
            fuz_launch(ctx, k_pileup_gather_tma, gather_ctas, FUZ_PTILE_THREADS + 32, gsm, st, S, cap_sites, out->d_counts, ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_pileup_gather_tma");
        } else {
            fuz_launch(ctx, k_pileup_gather, n_tiles, FUZ_PTILE_THREADS, 0, st, S, cap_sites, out->d_counts, ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_pileup_gather");
        }
        if (ev1) FUZ_CUDA(ctx, cudaEventRecord(ev1, st));
    } else {
        FUZ_CUDA(ctx, cudaMemsetAsync(S.counts, 0, 16 * (size_t)in->total_glen, st));
        if (ev0) FUZ_CUDA(ctx, cudaEventRecord(ev0, st));
        if (n_rec > 0) {
            fuz_launch(ctx, k_project, ctx->project_ctas, 256, 0, st, in->d_rec_buf, in->d_rec_off, n_rec, in->d_ctg_rec_off, in->d_ctg_goff,
                                                       n_ctg, S, ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_project");      // filter + the projection the variant_map rows come from
            fuz_launch(ctx, k_pileup_atomic, FUZ_GRID_BLOCKS, 256, 0, st, in->d_rec_buf, in->d_rec_off, n_rec, in->d_ctg_rec_off,
                                                              in->d_ctg_goff, n_ctg, S, S.counts, ctx->d_status);
            FUZ_LAUNCH_CHECK(ctx, "k_pileup_atomic");
        }
        if (ev1) FUZ_CUDA(ctx, cudaEventRecord(ev1, st));
        fuz_launch(ctx, k_tile_ranges, (n_tiles + 7) / 8, 256, 0, st, n_tiles, n_ctg, in->d_ctg_goff, in->d_ctg_rec_off, S, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_tile_ranges");
        fuz_launch(ctx, k_het_from_counts, n_tiles, FUZ_PTILE_THREADS, 0, st, S.counts, S, cap_sites, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_het_from_counts");
    }
    if ((rc = fuz_scan_i32_wide(ctx, S.tile_site_cnt, S.tile_site_off, n_tiles, FUZ_FIN_SITES, cap_sites))) return rc;
    fuz_launch(ctx, k_sites_place, FUZ_GRID_BLOCKS, 256, 0, st, S, in->d_ctg_goff, *out, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_sites_place");
    if ((rc = fuz_scan_i32_wide(ctx, S.site_rows, S.site_row_off, cap_sites, FUZ_FIN_VMAP, out->cap_vmap, &ctx->d_status->n_sites))) return rc;
    if (ctx->join_pending) {                  // q_ids assigned on the side stream (fuz_phase_batch)
        FUZ_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_join, 0));
        ctx->join_pending = false;
    }
    if (seg_path) {
        fuz_launch(ctx, k_signature_seg, FUZ_GRID_BLOCKS, 256, 0, st, in->d_rec_buf, in->d_rec_qid, S, *out, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_signature_seg");
    } else {
        fuz_launch(ctx, k_signature, 148 * ctx->grid_sig, 256, 0, st, in->d_rec_qid, in->d_ctg_rec_off, S, *out, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_signature");
    }
    return FUZ_OK;
}

extern "C" int fuz_het_call(fuz_ctx *ctx, const fuz_batch *in, fuz_outputs *out) {
    if (ctx && in && !in->d_rec_qid) return fuz_fail(ctx, FUZ_E_ARG, "fuz_het_call needs d_rec_qid (fuz_assign_qids or fuz_phase_batch)");
    return fuz_het_call_impl(ctx, in, out);
}
