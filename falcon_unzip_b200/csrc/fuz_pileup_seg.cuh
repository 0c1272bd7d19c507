// Segment-list pileup (pileup_impl 0): included by fuz_hetcall.cu inside its anonymous namespace.
//
// make_het_call (falcon_unzip/phasing.py:63-129) without a reference-aligned copy of the reads:
//   k_segments    one LANE per record walks the CIGAR once: filter totals of phasing.py:63-75, validation, and the
//                 match segments (maximal stretches of M/=/X with a constant query - reference offset; N/H/P advance
//                 nothing, phasing.py:77-96) as int4 {ref start, ref end, query - ref, 0} in global memory
//   k_tile_lists  one warp per 8192-position tile: the accepted reads overlapping the tile, in file order, each as a
//                 32-byte entry: which of its segments intersect the tile and which 16-byte aligned slice of its
//                 4-bit SEQ holds their bases
//   k_pileup_tma  persistent CTAs, 1 producer warp + 8 consumer warps.  The producer turns entries into bulk copies
//                 (cp.async.bulk global -> shared, completion on an mbarrier): SEQ slice + segment slice of 4 reads
//                 per pipeline stage, running ahead across tile borders.  A consumer thread owns one QUAD of the tile
//                 (32 positions = 128 bits of 4-bit codes).  Per stage: (1) the segment STARTS inside the tile are dealt
//                 to the threads as dense tasks (one per thread): the piece of the segment inside the quad it starts in
//                 is cut and OR-ed into a shared-memory fix-up row; (2) every thread finds the segment covering the
//                 start of its quad (the sorted segment ends sit one per lane: shuffle binary search), cuts its 32
//                 nibbles out of the staged SEQ (two 128-bit loads, word select, byte reversal, funnel shift), masks
//                 them to the segment, ORs the fix-up words and feeds the four words to bit-sliced Harley-Seal
//                 counters.  No divergent second pass at segment borders.  Het test and ordered site compaction at
//                 the end of a tile
//   k_signature_seg  one warp per het site, one lane per read of the tile's list: segment search + one SEQ byte
#define FUZ_G 4                    // reads per pipeline stage
#define FUZ_SLICE_CAP 4352         // bytes of SEQ staged per (read, tile): 4096 + alignment margins + insertions
#define FUZ_SEGW 96                // segments staged per (read, tile)
#define FUZ_NSTAGE 4
#define FUZ_ENT_SLOW 1             // the slice or the segment list exceed a stage slot: consumers read global memory
#define FUZ_CONSUMERS 256          // 32 positions each
#define FUZ_CW (FUZ_CONSUMERS / 32)
#define FUZ_PILEUP_THREADS (FUZ_CONSUMERS + 32)
static_assert(FUZ_TILE == 32 * FUZ_CONSUMERS, "one quad per consumer thread");

struct __align__(16) FuzTileEnt {
    int64_t seq_off;               // offset in rec_buf of the slice; rec_buf + seq_off is 16-byte aligned
    int32_t seq_bytes;             // multiple of 16
    int32_t dqb;                   // nibble index inside the slice of query base q is q + dqb
    int32_t seg_src;               // first segment of the read (index into segs) that ends behind the tile start
    int32_t nseg;                  // segments of the read intersecting the tile
    int32_t rec;                   // record index
    int32_t flags;
};
static_assert(sizeof(FuzTileEnt) == 32, "tile entry layout");

struct __align__(16) FuzStage {
    uint8_t seq[FUZ_G][FUZ_SLICE_CAP];
    int4 segs[FUZ_G][FUZ_SEGW];
    FuzTileEnt ent[FUZ_G];
    int32_t n, tile, last, pad;
};

// progress markers into a host-mapped buffer (option "trace_ptr", debugging only): one slot per (CTA, warp)
__device__ __forceinline__ void fuz_trace(volatile uint32_t *tr, int slot, uint32_t v) {
    if (tr && slot < 4096) { tr[slot] = v; __threadfence_system(); }
}

// ---------------------------------------------------------------- K1: CIGAR -> segments
// 32-bit arithmetic relative to the record start; a CIGAR whose lengths sum to 2^31 or more is rejected.
__global__ void __launch_bounds__(256) k_segments(
    const uint8_t *__restrict__ rec_buf, const int64_t *__restrict__ rec_off, int n_rec,
    const int32_t *__restrict__ ctg_rec_off, const int64_t *__restrict__ ctg_goff, int n_ctg, HetScratch S, fuz_status *st) {
    fuz_pdl_enter();
    const int lane = threadIdx.x & 31;
    long long acc_aligned = 0;
    int acc_accepted = 0;
    const int n_round = (n_rec + 31) & ~31;                         // warp-uniform trip count (warp-aggregated atomics)
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += gridDim.x * blockDim.x) {
        const bool live = r < n_rec;
        bool ok = false;
        int64_t off_r = 0, gstart64 = 0;
        int32_t pos = 0, l_seq = 0, n_cig = 0, l_name = 0, c = -1, alloc = 0;
        if (live) {
            off_r = rec_off[r];
            const int64_t off_n = rec_off[r + 1], off_p = r > 0 ? rec_off[r - 1] : 0;
            const uint8_t *rec = rec_buf + off_r;
            const int32_t block_size = (int32_t)fuz_ld_u32_un(rec);
            pos = (int32_t)fuz_ld_u32_un(rec + 8);
            const uint32_t w12 = fuz_ld_u32_un(rec + 12), w16 = fuz_ld_u32_un(rec + 16);
            l_seq = (int32_t)fuz_ld_u32_un(rec + 20);
            const int32_t prev_pos = r > 0 ? (int32_t)fuz_ld_u32_un(rec_buf + off_p + 8) : 0;
            c = fuz_upper_bound(ctg_rec_off, 0, n_ctg + 1, r) - 1;
            l_name = w12 & 0xFF;
            n_cig = w16 & 0xFFFF;
            S.r_flags[r] = 0; S.r_gstart[r] = 0; S.r_gend[r] = 0; S.r_nseg[r] = 0; S.r_seg_off[r] = 0;
            if (c < 0 || c >= n_ctg || pos < 0 || l_seq < 0 || off_n - off_r != (int64_t)block_size + 4 ||
                36 + l_name + 4 * (int64_t)n_cig + ((int64_t)l_seq + 1) / 2 + l_seq > (int64_t)block_size + 4) {
                fuz_raise(st, FUZ_E_BADRECORD, r);
            } else {
                ok = true;
                gstart64 = ctg_goff[c] + pos;
                if (r > ctg_rec_off[c] && prev_pos > pos) fuz_raise(st, FUZ_E_UNSORTED, r);    // coordinate order inside the contig
                alloc = n_cig / 2 + 1;                   // segments are separated by at least one I / D / S operation
            }
        }
        // room for the segments: one atomic per warp
        const int incl = fuz_warp_incl_scan(alloc, lane);
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        long long wbase = 0;
        if (lane == 31 && tot) wbase = (long long)atomicAdd((unsigned long long *)&st->n_segments, (unsigned long long)tot);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        const long long seg_base = wbase + incl - alloc;
        if (ok && seg_base + alloc > S.seg_cap) { fuz_raise(st, FUZ_E_CAPACITY, 6); ok = false; }
        bool accept = false;
        uint32_t span = 0;
        if (ok) {
            const uintptr_t ca = reinterpret_cast<uintptr_t>(rec_buf + off_r + 36 + l_name);
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(ca & ~(uintptr_t)3);
            const uint32_t sh = (uint32_t)(ca & 3) * 8;
            int4 *out = S.segs + seg_base;
            const int4 *out0 = out;
            const int gs = (int)gstart64;
            uint32_t rp = 0, qp = 0, skip = 0, aligned = 0;                  // relative to the record start
            int seg_rs = 0, seg_dq = 0;                                      // absolute start, query - reference
            unsigned long long total = 0;
            bool open = false, overrun = false;
            uint32_t opmax = 0;
            uint32_t prev = __ldg(wp);
            // class of an operation, two bits each: 1 = M = X (advance both), 2 = I S (query), 3 = D (reference), 0 = N H P
            // and unknown codes (nothing, phasing.py:77-96).  Bit 0 = advances the reference, bit0 ^ bit1 = the query
            const uint32_t kClass = 1u | 2u << 2 | 3u << 4 | 2u << 8 | 1u << 14 | 1u << 16;
#pragma unroll 4
            for (int k = 0; k < n_cig; k++) {
                const uint32_t cur = __ldg(wp + k + 1);
                const uint32_t cw = __funnelshift_r(prev, cur, sh);
                prev = cur;
                const uint32_t len = cw >> 4, op = cw & 15;
                const uint32_t cls = len ? (kClass >> (2 * op)) & 3u : 0u;
                opmax = max(opmax, op);
                total += len;
                if (op == 4) skip += len;
                if (cls == 1) {
                    if (!open) { open = true; seg_rs = gs + (int)rp; seg_dq = (int)(qp - rp) - gs; }
                } else if (cls && open) {                                    // I D S end a segment
                    *out++ = make_int4(seg_rs, gs + (int)rp, seg_dq, 0);
                    aligned += (uint32_t)(gs + (int)rp - seg_rs);
                    overrun |= qp > (uint32_t)l_seq;                         // IndexError phasing.py:84
                    open = false;
                }
                rp += (cls & 1u) ? len : 0u;
                qp += ((cls ^ (cls >> 1)) & 1u) ? len : 0u;
            }
            if (open) {
                *out++ = make_int4(seg_rs, gs + (int)rp, seg_dq, 0);
                aligned += (uint32_t)(gs + (int)rp - seg_rs);
                overrun |= qp > (uint32_t)l_seq;
            }
            const int n_seg = (int)(out - out0);
            span = rp;
            if (opmax > 8 || total == 0 || total >= 0x80000000ull || gstart64 + rp > 0x7fffffffLL) {   // unknown op / ZeroDivisionError phasing.py:72
                fuz_raise(st, FUZ_E_BADRECORD, r);
                ok = false;
            } else {
                // phasing.py:72-75 in IEEE double, same operation order as the reference
                accept = (skip == 0 || !(1.0 - 1.0 * (double)skip / (double)total < 0.1)) && !(total < 2000);
                if (accept && overrun) { fuz_raise(st, FUZ_E_BADRECORD, r); accept = false; }
                S.r_gstart[r] = gs;
                S.r_gend[r] = accept ? gs + (int32_t)span : gs;
                S.r_flags[r] = accept ? 1 : 0;
                S.r_seg_off[r] = (int32_t)seg_base;
                S.r_nseg[r] = accept ? n_seg : 0;
                S.r_seq[r] = off_r + 36 + l_name + 4 * (int64_t)n_cig;
                if (accept) { acc_aligned += aligned; acc_accepted += 1; }
            }
        }
        // the words of the reference-aligned projection (k_project_seg): whole quads, one atomic per warp
        if (S.proj_cap > 0) {
            const int nw = (accept && span > 0) ? (int)((((gstart64 + span - 1) >> 3) - ((gstart64 >> 3) & ~3LL) + 4) & ~3LL) : 0;
            const int incl_w = fuz_warp_incl_scan(nw, lane);
            const int tot_w = __shfl_sync(0xffffffffu, incl_w, 31);
            long long pbase = 0;
            if (lane == 31 && tot_w) pbase = (long long)atomicAdd((unsigned long long *)&st->reserved[0], (unsigned long long)tot_w);
            pbase = __shfl_sync(0xffffffffu, pbase, 31) + incl_w - nw;
            if (live) {
                if (nw && pbase + nw > S.proj_cap) { fuz_raise(st, FUZ_E_CAPACITY, 6); S.r_nwords[r] = 0; S.r_woff[r] = 0; S.r_flags[r] = 0; }
                else { S.r_nwords[r] = nw; S.r_woff[r] = (int32_t)pbase; }
            }
        }
        // last accepted record and longest span of the contig: one atomic per warp when the warp sits in one contig
        const int c0 = __shfl_sync(0xffffffffu, c, 0);
        const int rmax = accept ? r : -1, smax = accept ? (int)span : 0;
        if (__all_sync(0xffffffffu, c == c0 || !live)) {
            const int wr = __reduce_max_sync(0xffffffffu, rmax), ws = __reduce_max_sync(0xffffffffu, smax);
            if (lane == 0 && wr >= 0) { atomicMax(&S.ctg_last_rec[c0], wr); atomicMax(&S.ctg_maxspan[c0], ws); }
        } else if (accept) {
            atomicMax(&S.ctg_last_rec[c], r);
            atomicMax(&S.ctg_maxspan[c], smax);
        }
    }
    acc_aligned = fuz_warp_sum64(acc_aligned);
    acc_accepted = fuz_warp_sum(acc_accepted);
    if (lane == 0 && acc_accepted) {
        atomicAdd((unsigned long long *)&st->aligned_bases, (unsigned long long)acc_aligned);
        atomicAdd((unsigned long long *)&st->n_accepted, (unsigned long long)acc_accepted);
    }
}

// ---------------------------------------------------------------- K1b: reads of every tile
// One warp per tile.  The limit is POS_last of the contig (see k_tile_ranges).  Entries of a tile are contiguous and in
// file order (k_signature_seg depends on it); the tiles claim their ranges with one atomic each.
__global__ void __launch_bounds__(256) k_tile_lists(int n_tiles, int n_ctg, const uint8_t *__restrict__ rec_buf,
                                                    const int64_t *__restrict__ ctg_goff, const int32_t *__restrict__ ctg_rec_off,
                                                    HetScratch S, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const uint32_t lt = (1u << lane) - 1u;
    const uintptr_t buf_a = reinterpret_cast<uintptr_t>(rec_buf);
    for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n_tiles; t += (gridDim.x * blockDim.x) >> 5) {
        const int64_t t0l = (int64_t)t * FUZ_TILE;
        int c = fuz_warp_lower_bound<int64_t>(ctg_goff, 0, n_ctg + 1, t0l + 1, lane) - 1;
        if (c >= n_ctg) c = n_ctg - 1;
        const int r0 = ctg_rec_off[c], r1 = ctg_rec_off[c + 1];
        const int t0 = (int)t0l, t1 = (int)(t0l + FUZ_TILE);
        const int span = S.ctg_maxspan[c], lr = S.ctg_last_rec[c];
        const int rlo = fuz_warp_lower_bound<int32_t>(S.r_gstart, r0, r1, t0 - span + 1, lane);
        const int rhi = fuz_warp_lower_bound<int32_t>(S.r_gstart, r0, r1, t1, lane);
        int cnt = 0;
        for (int rb = rlo; rb < rhi; rb += 32) {
            const int r = rb + lane;
            const bool ov = r < rhi && S.r_flags[r] && S.r_gend[r] > t0 && S.r_gstart[r] < t1;
            cnt += __popc(__ballot_sync(0xffffffffu, ov));
        }
        long long base = 0;
        if (lane == 0 && cnt) base = (long long)atomicAdd((unsigned long long *)&st->reserved[0], (unsigned long long)cnt);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base + cnt > S.ent_cap) {
            if (lane == 0) fuz_raise(st, FUZ_E_CAPACITY, 9);
            cnt = 0;
        }
        if (lane == 0) {
            S.tile_ctg[t] = c;
            S.tile_limit[t] = lr >= 0 ? S.r_gstart[lr] : (int32_t)ctg_goff[c];
            S.tile_ent_base[t] = (int32_t)base;
            S.tile_ent_cnt[t] = cnt;
        }
        if (cnt == 0) continue;
        int run = 0;
        for (int rb = rlo; rb < rhi; rb += 32) {
            const int r = rb + lane;
            const bool ov = r < rhi && S.r_flags[r] && S.r_gend[r] > t0 && S.r_gstart[r] < t1;
            const uint32_t m = __ballot_sync(0xffffffffu, ov);
            if (ov) {
                const int so = S.r_seg_off[r], ns = S.r_nseg[r];
                const int4 *sg = S.segs + so;
                int lo = 0, hi = ns;                                  // first segment ending behind the tile start
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&sg[mid].y) <= t0) lo = mid + 1; else hi = mid; }
                const int s_lo = lo;
                hi = ns;                                              // first segment starting at or behind the tile end
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&sg[mid].x) < t1) lo = mid + 1; else hi = mid; }
                const int nst = lo - s_lo;
                FuzTileEnt e;
                e.seq_off = 0; e.seq_bytes = 0; e.dqb = 0; e.seg_src = so + s_lo; e.nseg = nst; e.rec = r; e.flags = 0;
                if (nst > 0) {
                    const int4 first = __ldg(&sg[s_lo]), last = __ldg(&sg[lo - 1]);
                    const long long qlo = (long long)max(t0, first.x) + first.z;          // first and last query base used
                    const long long qhi = (long long)min(t1, last.y) + last.z - 1;
                    const uintptr_t seq_a = buf_a + (uintptr_t)S.r_seq[r];
                    // >= 16 bytes before the first base (a quad may start up to 31 positions before its segment) and 32
                    // behind the last (two 16-byte chunks are read per quad); rec_buf carries 64 bytes of slack
                    const uintptr_t A = (seq_a + (uintptr_t)(qlo >> 1) - 16) & ~(uintptr_t)15;
                    const uintptr_t B = (seq_a + (uintptr_t)(qhi >> 1) + 32 + 15) & ~(uintptr_t)15;
                    e.seq_off = (int64_t)(A - buf_a);
                    e.seq_bytes = (int32_t)min((long long)(B - A), 0x7ffffff0LL);
                    e.dqb = (int32_t)(2 * ((long long)seq_a - (long long)A));
                    if (B - A > FUZ_SLICE_CAP || nst > FUZ_SEGW) e.flags = FUZ_ENT_SLOW;
                }
                int4 *dst = reinterpret_cast<int4 *>(S.ents + base + run + __popc(m & lt));
                const int4 *src = reinterpret_cast<const int4 *>(&e);
                dst[0] = src[0]; dst[1] = src[1];
            }
            run += __popc(m);
        }
    }
}

// ---------------------------------------------------------------- segment-major projection (pileup_impl 3)
// The reference-aligned 4-bit projection of k_project (one word per 8 positions, A=1 C=2 G=4 T=8), produced from the
// segment lists of k_segments: one warp per record, one LANE per segment, segments handed out from a warp-local counter.
// Inside a segment the query - reference offset is constant, so the lane walks its quads in order with a loop-invariant
// shift and a sliding window over the 4-bit SEQ: four new words, nibble swap, funnel shift, one 128-bit store per quad;
// no search for the segment of a quad and no masks except in the first and last quad of the segment.  Every quad of the
// record is written exactly once, by the LAST segment that touches it: that lane ORs in the tails of the earlier segments
// ending in the quad (a deletion or insertion inside the quad) and writes the zero quads in front of it (long deletions,
// the padding of the record).
__global__ void __launch_bounds__(256) k_project_seg(const uint8_t *__restrict__ rec_buf, int n_rec, HetScratch S, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    __shared__ int s_next[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    int r_next = 0;
    for (int r = warp_g; r < n_rec; r = r_next) {
        if (lane == 0) r_next = n_warps + atomicAdd(S.rec_cursor, 1);
        r_next = __shfl_sync(0xffffffffu, r_next, 0);
        const int n_words = S.r_nwords[r];
        if (!S.r_flags[r] || n_words == 0) continue;
        const int ns = S.r_nseg[r], n_quads = n_words >> 2;
        const int4 *__restrict__ segs = S.segs + S.r_seg_off[r];
        ProjRec R;
        {
            const uintptr_t sa = reinterpret_cast<uintptr_t>(rec_buf + S.r_seq[r]);
            R.base4 = reinterpret_cast<const uint32_t *>(sa & ~(uintptr_t)3);
            R.nphase = (int)(sa & 3) * 2;
            R.W0 = fuz_row_w0(S.r_gstart[r]);
            R.out = S.proj + S.r_woff[r];
        }
        auto quad_of = [&](int pos) { return ((pos >> 3) - R.W0) >> 2; };
        auto quad_pos = [&](int q) { return (R.W0 + 4 * q) << 3; };
        uint4 *out4 = reinterpret_cast<uint4 *>(R.out);
        if (ns == 0) {                                                       // a record of deletions only: nothing aligned
            for (int q = lane; q < n_quads; q += 32) out4[q] = make_uint4(0u, 0u, 0u, 0u);
            continue;
        }
        __syncwarp();
        if (lane == 0) s_next[wib] = 32;
        __syncwarp();
        int j = lane;
        while (j < ns) {
            const int4 sg = __ldg(segs + j);
            const int qf = quad_of(sg.x), ql = quad_of(sg.y - 1);
            const int q_prev = j > 0 ? quad_of(__ldg(&segs[j - 1].y) - 1) : -1;             // last quad of the previous segment
            const int qn = j + 1 < ns ? quad_of(__ldg(&segs[j + 1].x)) : 0x7fffffff;         // first quad of the next one
            for (int q = q_prev + 1; q < qf; q++) out4[q] = make_uint4(0u, 0u, 0u, 0u);      // nothing aligned there
            // the quads I write: qf .. ql, except ql when the next segment starts inside it
            const int q_end = qn == ql ? ql - 1 : ql;
            if (qf <= q_end) {
                // first quad: masked to my start (and to my end when it is also my last), plus the tails of earlier segments
                uint32_t v[4];
                quad_piece(R, quad_pos(qf), sg.x, sg.y, sg.z, v);
                for (int k = j - 1; k >= 0; k--) {
                    const int4 sk = __ldg(segs + k);
                    if (quad_of(sk.y - 1) != qf) break;
                    uint32_t u[4];
                    quad_piece(R, quad_pos(qf), sk.x, sk.y, sk.z, u);
                    v[0] |= u[0]; v[1] |= u[1]; v[2] |= u[2]; v[3] |= u[3];
                }
                out4[qf] = make_uint4(v[0], v[1], v[2], v[3]);
                // middle quads: whole quads of my segment, constant shift, sliding window over SEQ
                const int q_mid_end = min(q_end, ql - 1);                    // ql itself may be partial: handled below
                if (qf + 1 <= q_mid_end) {
                    const int n0 = quad_pos(qf + 1) + sg.z + R.nphase;       // nibble of the first middle position relative to base4
                    const uint32_t *src = R.base4 + (n0 >> 3);
                    const uint32_t sh = (uint32_t)(n0 & 7) * 4;
                    uint32_t m0 = swap_nibbles(__ldg(src));
                    for (int q = qf + 1; q <= q_mid_end; q++, src += 4) {
                        const uint32_t m1 = swap_nibbles(__ldg(src + 1)), m2 = swap_nibbles(__ldg(src + 2)),
                                       m3 = swap_nibbles(__ldg(src + 3)), m4 = swap_nibbles(__ldg(src + 4));
                        uint32_t w0 = __funnelshift_r(m0, m1, sh), w1 = __funnelshift_r(m1, m2, sh),
                                 w2 = __funnelshift_r(m2, m3, sh), w3 = __funnelshift_r(m3, m4, sh);
                        if (ambiguous_nibbles(w0) | ambiguous_nibbles(w1) | ambiguous_nibbles(w2) | ambiguous_nibbles(w3)) {   // ambiguity codes: rare
                            w0 = keep_acgt(w0); w1 = keep_acgt(w1); w2 = keep_acgt(w2); w3 = keep_acgt(w3);
                        }
                        out4[q] = make_uint4(w0, w1, w2, w3);
                        m0 = m4;
                    }
                }
                // last quad (when it is not the first): masked to my end
                if (ql > qf && q_end == ql) {
                    quad_piece(R, quad_pos(ql), sg.x, sg.y, sg.z, v);
                    out4[ql] = make_uint4(v[0], v[1], v[2], v[3]);
                }
            }
            if (j + 1 == ns)                                                 // the padding behind the last segment
                for (int q = ql + 1; q < n_quads; q++) out4[q] = make_uint4(0u, 0u, 0u, 0u);
            j = atomicAdd(&s_next[wib], 1);
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------- mbarrier / bulk copy (sm_90+ PTX)
__device__ __forceinline__ uint32_t fuz_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fuz_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fuz_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fuz_mbar_arrive_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fuz_mbar_wait(uint32_t bar, uint32_t parity, unsigned sleep_ns = 100) {
    uint32_t done;
    uint32_t spins = 0;
    for (;;) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(sleep_ns);                                             // leave the issue slots to the other warps
        if (++spins > (1u << 24)) __trap();                           // a lost arrival becomes a CUDA error, not a hang
    }
}
__device__ __forceinline__ void fuz_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// barrier of the 256 consumer threads.  bar.sync is warp-ALIGNED: a warp that reaches it diverged (lanes leave the
// mbarrier spin loop or the task loop at different times) would be counted once per arriving fragment, so the warp is
// reconverged first
__device__ __forceinline__ void fuz_consumer_sync() {
    __syncwarp();
    asm volatile("bar.sync 1, %0;" ::"n"(FUZ_CONSUMERS) : "memory");
}

// ---------------------------------------------------------------- TMA-fed register pileup over the projection (pileup_impl 0)
// k_pileup_gather as a persistent producer / consumer pipeline.  The projection rows are reference aligned, so the part of a
// read inside a 2048-position tile is one contiguous run of <= 256 words: the producer warp finds the reads of a tile, turns
// each into one bulk copy (cp.async.bulk, 16-byte aligned start, completion on the stage's mbarrier) into a ring of 4 stages
// of 15 reads, and runs ahead across tile borders; the 256 consumer threads (one 8-position word each) read their word of
// every read from shared memory and feed the carry-save tree.  The dependent chain of the plain kernel (tile range ->
// record fields -> compaction -> loads, 3.3 waves of short-lived CTAs) is what kept it at 3.5 TB/s.
#define FUZ_GG 15                   // reads per stage = inputs of one carry-save tree
#define FUZ_GSLOT 256               // words per read slot = words of a tile: tile word j of the read sits at slot[j]
#define FUZ_GSTAGES 4
struct __align__(16) FuzGStage {
    uint32_t slot[FUZ_GG][FUZ_GSLOT];
    int2 ent[FUZ_GG + 1];           // x = first tile word of the read, y = its words inside the tile (0: empty slot)
    int32_t n, tile, last, pad;
};

__global__ void __launch_bounds__(FUZ_PTILE_THREADS + 32, 3) k_pileup_gather_tma(HetScratch S, int64_t cap_sites, uint32_t *__restrict__ counts_out,
                                                                                  fuz_status *st) {
    fuz_pdl_enter();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FuzGStage *stages = reinterpret_cast<FuzGStage *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + FUZ_GSTAGES * sizeof(FuzGStage));     // full[], empty[]
    __shared__ int s_warp_tot[FUZ_NW];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < FUZ_GSTAGES; i++) {
            fuz_mbar_init(fuz_smem_u32(&bars[i]), 1);
            fuz_mbar_init(fuz_smem_u32(&bars[FUZ_GSTAGES + i]), FUZ_NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool dead = st->error != 0;
    if (warp == FUZ_NW) {
        // ------------------------------------------------------------ producer
        int stage = 0, fill = 0;
        uint32_t phase = 0;
        const uint32_t lt = (1u << lane) - 1u;
        bool open = false;                                            // the current stage has been waited for
        // Software pipeline: the record fields of the NEXT round of 32 candidates (of this tile or the next) and the tile
        // after next are requested before the current round is turned into copies -- no load latency between rounds.
        int f_tile = 0, f_rlo = 0, f_rhi = 0;                         // lane 0: the fetched tile (loads may be in flight)
        auto fetch_tile = [&]() {
            if (lane == 0) {
                const int t = dead ? S.n_tiles : atomicAdd(S.tile_cursor, 1);
                const bool e = t >= S.n_tiles;
                f_tile = t; f_rlo = e ? 0 : S.tile_rlo[t]; f_rhi = e ? 0 : S.tile_rhi[t];
            }
        };
        struct Round { int fl, gs, ge, nw, wo; };
        auto load_round = [&](bool end_, int rhi_, int cb_) {
            const int r = cb_ + lane;
            const bool in = !end_ && r < rhi_;
            Round q;
            q.fl = in ? S.r_flags[r] : 0; q.gs = in ? S.r_gstart[r] : 0; q.ge = in ? S.r_gend[r] : 0;
            q.nw = in ? S.r_nwords[r] : 0; q.wo = in ? S.r_woff[r] : 0;
            return q;
        };
        fetch_tile();
        int tile = __shfl_sync(0xffffffffu, f_tile, 0), rlo = __shfl_sync(0xffffffffu, f_rlo, 0), rhi = __shfl_sync(0xffffffffu, f_rhi, 0);
        bool end = tile >= S.n_tiles;
        if (!end) fetch_tile();
        int cb = rlo;
        Round cur = load_round(end, rhi, cb);
        for (;;) {
            const bool last_round = cb + 32 >= rhi;
            int n_tile = tile, n_rlo = rlo, n_rhi = rhi, n_cb = cb + 32;
            bool n_end = end;
            if (last_round && !end) {
                n_tile = __shfl_sync(0xffffffffu, f_tile, 0); n_rlo = __shfl_sync(0xffffffffu, f_rlo, 0); n_rhi = __shfl_sync(0xffffffffu, f_rhi, 0);
                n_end = n_tile >= S.n_tiles;
                if (!n_end) fetch_tile();
                n_cb = n_rlo;
            }
            const Round nxt = load_round(n_end || (last_round && end), n_rhi, n_cb);
            {
                const int t0 = tile * FUZ_PTILE, t1 = t0 + FUZ_PTILE, Wt0 = t0 >> 3;
                const bool ok = cur.fl && cur.ge > t0 && cur.gs < t1;                      // (fl == 0 outside the range)
                int2 e = make_int2(0, 0);
                const uint32_t *src = nullptr;
                uint32_t bytes = 0;
                if (ok) {
                    // rows start on 4-word boundaries of the global grid (fuz_row_w0) and are whole quads long: the part of
                    // the row inside the tile is a 16-byte aligned run, copied to its place in the slot
                    const int W0 = fuz_row_w0(cur.gs);
                    const int lo = max(W0, Wt0) - Wt0, hi = min(W0 + cur.nw, Wt0 + FUZ_PTILE / 8) - Wt0;
                    src = S.proj + ((int64_t)cur.wo + (Wt0 + lo - W0));
                    bytes = 4u * (uint32_t)(hi - lo);
                    e = make_int2(lo, hi - lo);
                }
                uint32_t m = __ballot_sync(0xffffffffu, ok);
                do {
                    FuzGStage &sg = stages[stage];
                    const uint32_t full = fuz_smem_u32(&bars[stage]), empty = fuz_smem_u32(&bars[FUZ_GSTAGES + stage]);
                    if (!open) { fuz_mbar_wait(empty, phase ^ 1u, 200); open = true; }
                    const int take = min(FUZ_GG - fill, __popc(m));
                    // the `take` lowest set lanes of m go into slots fill .. fill + take - 1
                    const int my = __popc(m & lt);
                    const bool mine = ((m >> lane) & 1u) && my < take;
                    const uint32_t tx = __reduce_add_sync(0xffffffffu, mine ? bytes : 0u);
                    if (lane == 0 && tx) asm volatile("mbarrier.expect_tx.shared::cta.b64 [%0], %1;" ::"r"(full), "r"(tx) : "memory");
                    __syncwarp();
                    if (mine) {
                        sg.ent[fill + my] = e;
                        fuz_bulk_g2s(fuz_smem_u32(&sg.slot[fill + my][e.x]), src, bytes, full);
                    }
                    m &= ~__ballot_sync(0xffffffffu, mine);
                    fill += take;
                    const bool tile_done = last_round && m == 0;
                    if (fill == FUZ_GG || tile_done) {                // post the stage
                        if (lane == 0) { sg.n = end ? -1 : fill; sg.tile = tile; sg.last = tile_done; }
                        if (lane >= fill && lane <= FUZ_GG) sg.ent[lane] = make_int2(0, 0);      // empty slots contribute nothing
                        __syncwarp();
                        if (lane == 0) fuz_mbar_arrive(full);
                        fill = 0; open = false;
                        if (++stage == FUZ_GSTAGES) { stage = 0; phase ^= 1u; }
                    }
                } while (m);
            }
            if (last_round && end) break;
            tile = n_tile; rlo = n_rlo; rhi = n_rhi; cb = n_cb; end = n_end;
            cur = nxt;
        }
        return;
    }
    // ---------------------------------------------------------------- consumers: thread = one 8-position word of the tile
    uint32_t *spill = S.spill + (size_t)blockIdx.x * 16 * FUZ_PTILE_THREADS;
#define C16G(b, j) spill[(4 * (b) + (j)) * FUZ_PTILE_THREADS + tid]
    uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int n_reads_seen = 0, groups_in_acc = 0;
    bool spilled = false;
    int stage = 0;
    uint32_t phase = 0;
    for (;;) {
        FuzGStage &sg = stages[stage];
        fuz_mbar_wait(fuz_smem_u32(&bars[stage]), phase);
        const int n = sg.n, tile = sg.tile, last = sg.last;
        if (n < 0) break;
        // branch free: every slot is read (tile word j of a read sits at slot[j]); what lies outside the read's words in
        // this tile -- and every empty slot -- is masked with the entry (two entries per 128-bit load)
        uint32_t x[FUZ_GG];
#pragma unroll
        for (int s = 0; s < FUZ_GG; s++) x[s] = sg.slot[s][tid];
#pragma unroll
        for (int s2 = 0; s2 < FUZ_GG + 1; s2 += 2) {
            const int4 e = *reinterpret_cast<const int4 *>(&sg.ent[s2]);
            if ((unsigned)(tid - e.x) >= (unsigned)e.y) x[s2] = 0;
            if (s2 + 1 < FUZ_GG && (unsigned)(tid - e.z) >= (unsigned)e.w) x[s2 + 1] = 0;
        }
        __syncwarp();
        if (lane == 0) fuz_mbar_arrive(fuz_smem_u32(&bars[FUZ_GSTAGES + stage]));
        if (++stage == FUZ_GSTAGES) { stage = 0; phase ^= 1u; }
        n_reads_seen += n;
        if (n > 0) {
            uint32_t s0, s1, s2, s3, s4, s5, k0, k1, k2, k3, k4, k5, k6, ones, t0_, t1_, d0, d1, d2, twos, fours, eights;
            FUZ_FA(x[0], x[1], x[2], s0, k0); FUZ_FA(x[3], x[4], x[5], s1, k1); FUZ_FA(x[6], x[7], x[8], s2, k2);
            FUZ_FA(x[9], x[10], x[11], s3, k3); FUZ_FA(x[12], x[13], x[14], s4, k4);
            FUZ_FA(s0, s1, s2, s5, k5); FUZ_FA(s3, s4, s5, ones, k6);
            FUZ_FA(k0, k1, k2, t0_, d0); FUZ_FA(k3, k4, k5, t1_, d1); FUZ_FA(t0_, t1_, k6, twos, d2);
            FUZ_FA(d0, d1, d2, fours, eights);
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
                        if (!spilled && (i & 1) == 0) C16G(b, i >> 1) = 0;
                        C16G(b, i >> 1) += plane_count(acc, 4 * i + b) << ((i & 1) * 16);
                    }
#pragma unroll
                for (int k = 0; k < 8; k++) acc[k] = 0;
                groups_in_acc = 0;
                spilled = true;
            }
        }
        if (!last) continue;
        // ------------------------------------------------------------ end of the tile: het test, ordered sites (as k_pileup_gather)
        if (n_reads_seen > 65535 && tid == 0) fuz_raise(st, FUZ_E_DEPTH, tile);
        const int t0 = tile * FUZ_PTILE;
        const int pos_limit = S.tile_limit[tile];
        uint32_t cnt[8][4];
        uint32_t hetmask = 0;
        if (!spilled && !counts_out) {
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
                    cnt[i][b] = (spilled ? (C16G(b, i >> 1) >> ((i & 1) * 16)) & 0xFFFFu : 0u) + plane_count(acc, 4 * i + b);
            if (counts_out) {
                uint4 *o = reinterpret_cast<uint4 *>(counts_out) + (size_t)t0 + (size_t)tid * 8;
#pragma unroll
                for (int i = 0; i < 8; i++) o[i] = make_uint4(cnt[i][0], cnt[i][1], cnt[i][2], cnt[i][3]);
            }
            hetmask = het_mask_of(cnt, t0, pos_limit);
        }
        emit_tile_sites<true>(hetmask, cnt, tile, t0, S, cap_sites, st, s_warp_tot, &s_base);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] = 0;
        n_reads_seen = 0; groups_in_acc = 0; spilled = false;
    }
#undef C16G
}

// ---------------------------------------------------------------- the cut
// The 32 nibbles of the read at reference positions Q0 .. Q0 + 31 under segment sg (position 8k + i of the quad in
// bits 28 - 4i of v[k]: the BAM nibble order after a byte reversal), masked to the part of the quad the segment
// covers.  seq = the slice as 16-byte chunks; the slice starts >= 16 bytes before the first base it is asked for.
template <bool GLOBAL>
__device__ __forceinline__ void fuz_cut_quad(const int4 sg, const uint4 *seq, int dqb, int Q0, uint32_t (&v)[4]) {
    const int n = (int)((uint32_t)Q0 + (uint32_t)sg.z + (uint32_t)dqb);      // nibble of position Q0 inside the slice (>= 1)
    const uint4 *cp = seq + (n >> 5);
    const uint4 a = GLOBAL ? __ldg(cp) : cp[0], b = GLOBAL ? __ldg(cp + 1) : cp[1];
    uint32_t m0 = a.x, m1 = a.y, m2 = a.z, m3 = a.w, m4 = b.x, m5 = b.y;
    if (n & 16) { m0 = m2; m1 = m3; m2 = m4; m3 = m5; m4 = b.z; m5 = b.w; }
    if (n & 8) { m0 = m1; m1 = m2; m2 = m3; m3 = m4; m4 = m5; }
    const uint32_t sh = (uint32_t)(n & 7) * 4;
    const uint32_t b0 = __byte_perm(m0, 0, 0x0123), b1 = __byte_perm(m1, 0, 0x0123), b2 = __byte_perm(m2, 0, 0x0123),
                   b3 = __byte_perm(m3, 0, 0x0123), b4 = __byte_perm(m4, 0, 0x0123);
    v[0] = __funnelshift_l(b1, b0, sh); v[1] = __funnelshift_l(b2, b1, sh);
    v[2] = __funnelshift_l(b3, b2, sh); v[3] = __funnelshift_l(b4, b3, sh);
    if (sg.y < Q0 + 32) {                                                    // the segment ends inside the quad
        const int hi = 4 * max(sg.y - Q0, 0);                                // in bits
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] &= ~__funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(hi - 32 * k, 0));
    }
    if (sg.x > Q0) {                                                         // ... starts inside (tasks; starts of reads)
        const int lo = 4 * (sg.x - Q0);
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] &= __funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(lo - 32 * k, 0));
    }
}

// The same cut without branches (shared memory only): nothing survives when `valid` is false.  Four of these run
// interleaved in the main pass of a stage.
__device__ __forceinline__ void fuz_cut_quad_bf(const int4 sg, const uint4 *seq, int dqb, int Q0, bool valid, uint32_t (&v)[4]) {
    const int n = (int)((uint32_t)Q0 + (uint32_t)sg.z + (uint32_t)dqb);
    const uint4 *cp = seq + (valid ? n >> 5 : 0);
    const uint4 a = cp[0], b = cp[1];
    uint32_t m0 = a.x, m1 = a.y, m2 = a.z, m3 = a.w, m4 = b.x, m5 = b.y;
    if (n & 16) { m0 = m2; m1 = m3; m2 = m4; m3 = m5; m4 = b.z; m5 = b.w; }
    if (n & 8) { m0 = m1; m1 = m2; m2 = m3; m3 = m4; m4 = m5; }
    const uint32_t sh = (uint32_t)(n & 7) * 4;
    const uint32_t b0 = __byte_perm(m0, 0, 0x0123), b1 = __byte_perm(m1, 0, 0x0123), b2 = __byte_perm(m2, 0, 0x0123),
                   b3 = __byte_perm(m3, 0, 0x0123), b4 = __byte_perm(m4, 0, 0x0123);
    const int hi = valid ? 4 * min(sg.y - Q0, 32) : 0, lo = 4 * min(max(sg.x - Q0, 0), 32);     // in bits
    v[0] = __funnelshift_l(b1, b0, sh) & __funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)lo) & ~__funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(hi, 0));
    v[1] = __funnelshift_l(b2, b1, sh) & __funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(lo - 32, 0)) & ~__funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(hi - 32, 0));
    v[2] = __funnelshift_l(b3, b2, sh) & __funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(lo - 64, 0)) & ~__funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(hi - 64, 0));
    v[3] = __funnelshift_l(b4, b3, sh) & __funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(lo - 96, 0)) & ~__funnelshift_rc(0xFFFFFFFFu, 0u, (uint32_t)max(hi - 96, 0));
}

__device__ __forceinline__ void fuz_fa(uint32_t a, uint32_t b, uint32_t c, uint32_t &s, uint32_t &cy) { s = xor3(a, b, c); cy = maj3(a, b, c); }

// Harley-Seal step of one word: four more one-bit inputs per bit position.  p[0..3] hold the weights 1, 2, 4, 8, the
// carries of weight 4 / 8 wait in pend4 / pend8 until their partner arrives (phase = stage index & 3).  In phase 3 the
// carry of weight 16 comes back in c16 and the caller ripples it into p[4..7] (fuz_hs_ripple, under a real branch);
// after that the value is an ordinary binary number in the planes p[0..7].
__device__ __forceinline__ void fuz_hs_add4(uint32_t (&p)[8], uint32_t &pend4, uint32_t &pend8, uint32_t &c16, uint32_t a, uint32_t b,
                                            uint32_t c, uint32_t d, int phase) {
    uint32_t cA, cB, c2, c4 = 0, s2, s3;
    fuz_fa(p[0], a, b, p[0], cA);
    fuz_fa(p[0], c, d, p[0], cB);
    fuz_fa(p[1], cA, cB, p[1], c2);
    // phase 0, 2: c2 waits; phase 1, 3: joins its partner
    fuz_fa(p[2], pend4, c2, s2, c4);
    if (phase & 1) p[2] = s2; else { pend4 = c2; c4 = 0; }
    fuz_fa(p[3], pend8, c4, s3, c16);
    if (phase == 3) p[3] = s3; else { if (phase == 1) pend8 = c4; c16 = 0; }
}
__device__ __forceinline__ void fuz_hs_ripple(uint32_t (&p)[8], uint32_t c16) {
#pragma unroll
    for (int k = 4; k < 8; k++) { const uint32_t nc = p[k] & c16; p[k] ^= c16; c16 = nc; }
}

// ---------------------------------------------------------------- K2: TMA-fed tile pileup
// debug (option "pileup_debug", experiments only): 1 = consumers skip the cuts (pipeline alone; results invalid), 2 = every
// entry takes the global-memory path (no bulk copies), 8 = no segment-start tasks, 16 = no main cut, 32 = no het test
__global__ void __launch_bounds__(FUZ_PILEUP_THREADS, 2) k_pileup_tma(const uint8_t *__restrict__ rec_buf, HetScratch S, int64_t cap_sites,
                                                                      uint32_t *__restrict__ counts_out, fuz_status *st, int debug, uint32_t *trace) {
    fuz_pdl_enter();
    volatile uint32_t *tr = trace;
    const int tslot = 16 + (blockIdx.x * 9 + (threadIdx.x >> 5)) * 2;
    uint32_t tcount = 0;
#ifdef FUZ_TRACE
#define FUZ_TR(code) do { if (tr && (threadIdx.x & 31) == 0) { fuz_trace(tr, tslot, (uint32_t)(code)); fuz_trace(tr, tslot + 1, ++tcount); } } while (0)
#else
#define FUZ_TR(code) do { (void)tr; (void)tslot; (void)tcount; } while (0)
#endif
    FUZ_TR(1);
    extern __shared__ __align__(128) uint8_t smem_raw[];
    FuzStage *stages = reinterpret_cast<FuzStage *>(smem_raw);
    uint32_t (*fix)[FUZ_G][FUZ_TILE / 8] = reinterpret_cast<uint32_t (*)[FUZ_G][FUZ_TILE / 8]>(smem_raw + FUZ_NSTAGE * sizeof(FuzStage));
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + FUZ_NSTAGE * sizeof(FuzStage) + sizeof(fix[0]));   // full[], empty[]
    __shared__ int s_warp_tot[FUZ_CW];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < FUZ_NSTAGE; i++) {
            fuz_mbar_init(fuz_smem_u32(&bars[i]), 1);
            fuz_mbar_init(fuz_smem_u32(&bars[FUZ_NSTAGE + i]), FUZ_CW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < FUZ_G * (FUZ_TILE / 8); i += FUZ_PILEUP_THREADS) (&fix[0][0][0])[i] = 0;
    __syncthreads();
    const bool dead = st->error != 0;                                 // an earlier kernel failed: no tile is processed
    if (warp == FUZ_CW) {
        // ------------------------------------------------------------ producer
        // Everything with a global-memory latency is requested one step ahead: the id and entry range of the next tile
        // at the start of a tile, the entries of the next window (32 / FUZ_G groups; the first window of the next tile
        // behind the last one of this tile) before the groups of the current window are issued.
        int stage = 0;
        uint32_t phase = 0;
        auto fetch_tile = [&]() {
            int t = 0;
            if (lane == 0) t = dead ? S.n_tiles : atomicAdd(S.tile_cursor, 1);
            return __shfl_sync(0xffffffffu, t, 0);
        };
        auto load_window = [&](int base, int cnt, int g0, FuzTileEnt &e) {
            const int idx = g0 * FUZ_G + lane;
            const bool have = idx < cnt;
            if (have) {
                const int4 *src = reinterpret_cast<const int4 *>(S.ents + base + idx);
                int4 *ev = reinterpret_cast<int4 *>(&e);
                ev[0] = __ldg(src); ev[1] = __ldg(src + 1);
                if (debug & 2) ev[1].w |= FUZ_ENT_SLOW;
            }
            return have;
        };
        int tile = fetch_tile();
        int base = tile < S.n_tiles ? S.tile_ent_base[tile] : 0, cnt = tile < S.n_tiles ? S.tile_ent_cnt[tile] : 0;
        FuzTileEnt e, e_nxt;
        bool have = load_window(base, cnt, 0, e), have_nxt = false;
        for (;;) {
            FUZ_TR(20);
            const bool end = tile >= S.n_tiles;
            const int ntile = end ? tile : fetch_tile();
            const int nbase = ntile < S.n_tiles ? S.tile_ent_base[ntile] : 0, ncnt = ntile < S.n_tiles ? S.tile_ent_cnt[ntile] : 0;
            const int n_groups = end ? 1 : max(1, (cnt + FUZ_G - 1) / FUZ_G);
            for (int g0 = 0; g0 < n_groups; g0 += 32 / FUZ_G) {
                if (g0 + 32 / FUZ_G < n_groups) have_nxt = load_window(base, cnt, g0 + 32 / FUZ_G, e_nxt);
                else have_nxt = !end && load_window(nbase, ncnt, 0, e_nxt);
                const int g_cnt = min(32 / FUZ_G, n_groups - g0);
                for (int gi = 0; gi < g_cnt; gi++) {
                    FuzStage &sg = stages[stage];
                    const uint32_t full = fuz_smem_u32(&bars[stage]), empty = fuz_smem_u32(&bars[FUZ_NSTAGE + stage]);
                    FUZ_TR(21);
                    fuz_mbar_wait(empty, phase ^ 1u, 400);
                    FUZ_TR(22);
                    const bool mine = have && !end && lane / FUZ_G == gi;
                    const int slot = lane % FUZ_G;
                    uint32_t bytes = 0;
                    if (mine) {
                        int4 *dst = reinterpret_cast<int4 *>(&sg.ent[slot]);
                        const int4 *ev = reinterpret_cast<const int4 *>(&e);
                        dst[0] = ev[0]; dst[1] = ev[1];
                        if (!e.flags && e.nseg > 0) bytes = (uint32_t)e.seq_bytes + 16u * (uint32_t)e.nseg;
                    }
                    if (lane == 0) {
                        sg.n = end ? -1 : max(0, min(FUZ_G, cnt - (g0 + gi) * FUZ_G));
                        sg.tile = tile;
                        sg.last = g0 + gi == n_groups - 1;
                    }
                    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
                    __syncwarp();
                    if (lane == 0) fuz_mbar_arrive_expect(full, total);
                    __syncwarp();
                    if (bytes) {
                        fuz_bulk_g2s(fuz_smem_u32(&sg.seq[slot][0]), rec_buf + e.seq_off, (uint32_t)e.seq_bytes, full);
                        fuz_bulk_g2s(fuz_smem_u32(&sg.segs[slot][0]), S.segs + e.seg_src, 16u * (uint32_t)e.nseg, full);
                    }
                    if (++stage == FUZ_NSTAGE) { stage = 0; phase ^= 1u; }
                }
                e = e_nxt; have = have_nxt;
            }
            if (end) break;
            tile = ntile; base = nbase; cnt = ncnt;
        }
        FUZ_TR(29);
        return;
    }
    // ---------------------------------------------------------------- consumers: thread = one quad (32 positions)
    // 16-bit overflow counters beyond depth 240: [word][base][pair of positions] per thread, in global scratch
    uint32_t *spill = S.spill + (size_t)blockIdx.x * 64 * FUZ_CONSUMERS;
#define FUZ_C16(w, b, j) spill[(((w) * 4 + (b)) * 4 + (j)) * FUZ_CONSUMERS + tid]
    uint32_t P[4][8], pend4[4], pend8[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        pend4[w] = pend8[w] = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) P[w][k] = 0;
    }
    int n_reads_seen = 0, groups16 = 0, sit = 0;                      // sit = stages of the tile so far
    bool spilled = false;
    int stage = 0;
    uint32_t phase = 0;
    auto flush16 = [&]() {                                            // a group of 16 reads is complete: planes hold <= 15 * 16 + 15
        if (++groups16 == 15) {
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int b = 0; b < 4; b++)
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        if (!spilled && (i & 1) == 0) FUZ_C16(w, b, i >> 1) = 0;
                        FUZ_C16(w, b, i >> 1) += plane_count(P[w], 4 * (7 - i) + b) << ((i & 1) * 16);
                    }
#pragma unroll
            for (int w = 0; w < 4; w++)
#pragma unroll
                for (int k = 0; k < 8; k++) P[w][k] = 0;
            groups16 = 0;
            spilled = true;
        }
    };
    // Every consumer warp works through the stages on its own (no CTA barrier per stage): it owns one ROW of the tile
    // (32 quads = 1024 positions), handles the segment starts inside its row as warp-private tasks (fix-up words in its
    // own slice of `fix`) and only meets the other warps at the end of a tile.
    uint32_t (*wfix)[128] = reinterpret_cast<uint32_t (*)[128]>(&fix[0][0][0]) + warp * FUZ_G;      // [read][32 quads x 4 words]
    for (;;) {
        FuzStage &sg = stages[stage];
        FUZ_TR(2);
        fuz_mbar_wait(fuz_smem_u32(&bars[stage]), phase);
        const int n = sg.n, tile = sg.tile, last = sg.last;
        FUZ_TR(3 | (n << 8) | (tile << 16));
        if (n < 0) break;
        const int t0 = tile * FUZ_TILE;
        const int Q0 = t0 + 32 * tid, R0 = t0 + 1024 * warp;
        uint32_t x[FUZ_G][4];
        bool fast = !(debug & 17);
#pragma unroll
        for (int s = 0; s < FUZ_G; s++) fast = fast && !(s < n && sg.ent[s].flags);
        if (fast) {
            // ---- (0) lanes 0 .. 2 FUZ_G - 1: segments of read s starting before the row (even lane) / before its end (odd lane)
            int cnt_lt = 0, end_prev = 0;
            {
                const int s = (lane >> 1) & (FUZ_G - 1);
                const int bound = R0 + ((lane & 1) ? 1024 : 0);
                int hi = (lane < 2 * FUZ_G && s < n) ? sg.ent[s].nseg : 0;
                while (cnt_lt < hi) { const int mid = (cnt_lt + hi) >> 1; if (sg.segs[s][mid].x < bound) cnt_lt = mid + 1; else hi = mid; }
                if (cnt_lt > 0) end_prev = sg.segs[s][cnt_lt - 1].y;       // end of the last segment starting before the bound
            }
            int a[FUZ_G], tb[FUZ_G + 1], c0[FUZ_G];
            tb[0] = 0;
#pragma unroll
            for (int s = 0; s < FUZ_G; s++) {
                a[s] = __shfl_sync(0xffffffffu, cnt_lt, 2 * s);
                const int b = __shfl_sync(0xffffffffu, cnt_lt, 2 * s + 1);
                const int ep = __shfl_sync(0xffffffffu, end_prev, 2 * s);
                c0[s] = a[s] - (a[s] > 0 && ep > R0 ? 1 : 0);            // segments ending at or before the row start
                a[s] = max(a[s], 1);                                     // segment 0 is never a task (the main pass masks its start)
                tb[s + 1] = tb[s] + max(b - a[s], 0);
            }
            // ---- (1) segment starts inside my row as dense tasks: the piece of the segment inside the quad it starts in
            if (!(debug & 8)) {
                for (int k = lane; k < tb[FUZ_G]; k += 32) {
                    static_assert(FUZ_G == 4, "task table written for 4 reads per stage");
                    const int s = (k >= tb[1]) + (k >= tb[2]) + (k >= tb[3]);
                    const int j = k - (s == 0 ? 0 : s == 1 ? tb[1] : s == 2 ? tb[2] : tb[3]) + (s == 0 ? a[0] : s == 1 ? a[1] : s == 2 ? a[2] : a[3]);
                    const int4 seg = sg.segs[s][j];
                    if ((seg.x & 31) == 0) continue;                      // starts a quad: the main pass covers it
                    const int q0 = seg.x & ~31;
                    uint32_t v[4];
                    fuz_cut_quad<false>(seg, reinterpret_cast<const uint4 *>(sg.seq[s]), sg.ent[s].dqb, q0, v);
                    uint32_t *dst = &wfix[s][(q0 - R0) >> 3];
#pragma unroll
                    for (int w = 0; w < 4; w++)
                        if (v[w]) atomicOr(dst + w, v[w]);
                }
            }
            __syncwarp();
            // ---- (2) the segment covering the start of my quad: four branch-free searches and cuts, interleaved
            bool ovf = false;
#pragma unroll
            for (int s = 0; s < FUZ_G; s++) {
                const int nseg = s < n ? sg.ent[s].nseg : 0;
                const int E = c0[s] + lane < nseg ? sg.segs[s][c0[s] + lane].y : 0x7fffffff;
                const int e31 = __shfl_sync(0xffffffffu, E, 31);
                int idx = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int t = __shfl_sync(0xffffffffu, E, idx + step - 1);
                    idx += t <= Q0 ? step : 0;
                }
                ovf = ovf || (idx == 31 && e31 <= Q0);                // more than 31 ends inside the row: generic path
                const int c = c0[s] + idx;
                const int4 seg = sg.segs[s][min(c, FUZ_SEGW - 1)];
                fuz_cut_quad_bf(seg, reinterpret_cast<const uint4 *>(sg.seq[s]), sg.ent[s].dqb, Q0, c < nseg && seg.x < Q0 + 32, x[s]);
            }
            ovf = __any_sync(0xffffffffu, ovf);
            uint32_t bad = 0;
#pragma unroll
            for (int s = 0; s < FUZ_G; s++) {
                uint4 *fp = reinterpret_cast<uint4 *>(&wfix[s][4 * lane]);
                const uint4 f = *fp;
                *fp = make_uint4(0, 0, 0, 0);
                x[s][0] |= f.x; x[s][1] |= f.y; x[s][2] |= f.z; x[s][3] |= f.w;
                bad |= ambiguous_nibbles(x[s][0]) | ambiguous_nibbles(x[s][1]) | ambiguous_nibbles(x[s][2]) | ambiguous_nibbles(x[s][3]);
            }
            if (ovf) fast = false;
            else if (bad) {                                           // ambiguity codes never count (phasing.py:108-111)
#pragma unroll
                for (int s = 0; s < FUZ_G; s++)
#pragma unroll
                    for (int w = 0; w < 4; w++) x[s][w] = keep_acgt(x[s][w]);
            }
        }
        if (!fast) {
            // generic path (reads whose slice or segment list exceed a stage slot, rows with more than 31 segment ends): every
            // lane walks the segments that intersect its quad
#pragma unroll
            for (int s = 0; s < FUZ_G; s++) {
#pragma unroll
                for (int w = 0; w < 4; w++) x[s][w] = 0;
                if (s >= n || (debug & 17)) continue;
                const int nseg = sg.ent[s].nseg, dqb = sg.ent[s].dqb;
                const bool slow = sg.ent[s].flags != 0;
                const int4 *segs = slow ? S.segs + sg.ent[s].seg_src : &sg.segs[s][0];
                const uint4 *seq = slow ? reinterpret_cast<const uint4 *>(rec_buf + sg.ent[s].seq_off) : reinterpret_cast<const uint4 *>(sg.seq[s]);
                int c = 0, hi = nseg;
                while (c < hi) { const int mid = (c + hi) >> 1; if (segs[mid].y <= Q0) c = mid + 1; else hi = mid; }
                for (; c < nseg; c++) {
                    const int4 seg = segs[c];
                    if (seg.x >= Q0 + 32) break;
                    uint32_t v[4];
                    if (slow) fuz_cut_quad<true>(seg, seq, dqb, Q0, v); else fuz_cut_quad<false>(seg, seq, dqb, Q0, v);
#pragma unroll
                    for (int w = 0; w < 4; w++) x[s][w] |= keep_acgt(v[w]);
                }
            }
        }
        FUZ_TR(6);
        __syncwarp();
        if (lane == 0) fuz_mbar_arrive(fuz_smem_u32(&bars[FUZ_NSTAGE + stage]));      // the stage may be refilled
        if (++stage == FUZ_NSTAGE) { stage = 0; phase ^= 1u; }
        n_reads_seen += n;
        {
            uint32_t c16[4];
#pragma unroll
            for (int w = 0; w < 4; w++) fuz_hs_add4(P[w], pend4[w], pend8[w], c16[w], x[0][w], x[1][w], x[2][w], x[3][w], sit & 3);
            if ((++sit & 3) == 0) {
#pragma unroll
                for (int w = 0; w < 4; w++) fuz_hs_ripple(P[w], c16[w]);
                flush16();
            }
        }
        if (!last) continue;
        FUZ_TR(7);
        // ------------------------------------------------------------ end of the tile: het test, ordered sites
        for (; sit & 3; sit++) {                                      // complete the group of 16 with empty inputs
#pragma unroll
            for (int w = 0; w < 4; w++) {
                uint32_t c16;
                fuz_hs_add4(P[w], pend4[w], pend8[w], c16, 0u, 0u, 0u, 0u, sit & 3);
                if ((sit & 3) == 3) fuz_hs_ripple(P[w], c16);
            }
        }
        if (n_reads_seen > 65535 && tid == 0) fuz_raise(st, FUZ_E_DEPTH, tile);
        const int pos_limit = S.tile_limit[tile];
        auto counts_of = [&](int w, int i, uint32_t (&cn)[4]) {       // position 8w + i of my quad
#pragma unroll
            for (int b = 0; b < 4; b++)
                cn[b] = plane_count(P[w], 4 * (7 - i) + b) + (spilled ? (FUZ_C16(w, b, i >> 1) >> ((i & 1) * 16)) & 0xFFFFu : 0u);
        };
        uint32_t hetmask = 0;                                         // bit 8w + i
#pragma unroll
        for (int w = 0; w < 4; w++) {
            if (debug & 32) break;
            // a het site needs two bases with count >= 3 (second allele > 25 % of a depth >= 10): test that on the
            // planes and extract counts only for the few candidate positions
            const uint32_t ge3 = (P[w][0] & P[w][1]) | P[w][2] | P[w][3] | P[w][4] | P[w][5] | P[w][6] | P[w][7];
            const uint32_t two = (ge3 & (ge3 >> 1) & 0x77777777u) | (ge3 & (ge3 >> 2) & 0x33333333u) | (ge3 & (ge3 >> 3) & 0x11111111u);
            const bool all = spilled || counts_out != nullptr;
            if (two || all) {
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    if (all || ((two >> (4 * (7 - i))) & 7u)) {
                        uint32_t cn[4];
                        counts_of(w, i, cn);
                        if (counts_out) reinterpret_cast<uint4 *>(counts_out)[(size_t)Q0 + 8 * w + i] = make_uint4(cn[0], cn[1], cn[2], cn[3]);
                        if (Q0 + 8 * w + i < pos_limit && het_test(cn[0], cn[1], cn[2], cn[3])) hetmask |= 1u << (8 * w + i);
                    }
                }
            }
        }
        // ordered compaction of the tile's sites (position order = thread order, then bit order)
        {
            const int nh = __popc(hetmask);
            const int incl = fuz_warp_incl_scan(nh, lane);
            if (lane == 31) s_warp_tot[warp] = incl;
            FUZ_TR(8);
            fuz_consumer_sync();
            FUZ_TR(9);
            if (warp == 0) {
                const int t = lane < FUZ_CW ? s_warp_tot[lane] : 0;
                const int ti = fuz_warp_incl_scan(t, lane);
                if (lane < FUZ_CW) s_warp_tot[lane] = ti - t;
                if (lane == FUZ_CW - 1) {
                    int sbase = 0;
                    if (ti > 0) sbase = (int)atomicAdd((unsigned long long *)&st->need_sites, (unsigned long long)ti);
                    S.tile_site_base[tile] = sbase;
                    S.tile_site_cnt[tile] = ti;
                    s_base = sbase;
                }
            }
            fuz_consumer_sync();
            int64_t o = (int64_t)s_base + s_warp_tot[warp] + (incl - nh);
            for (uint32_t hm = hetmask; hm; hm &= hm - 1) {
                const int bit = __ffs(hm) - 1;
                if (o < cap_sites) {
                    uint32_t cn[4];
                    uint32_t pw[8];
                    // counts of position `bit` (dynamic word index: select the planes first)
#pragma unroll
                    for (int k = 0; k < 8; k++) pw[k] = (bit >> 3) == 0 ? P[0][k] : (bit >> 3) == 1 ? P[1][k] : (bit >> 3) == 2 ? P[2][k] : P[3][k];
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        cn[b] = plane_count(pw, 4 * (7 - (bit & 7)) + b);
                        if (spilled) cn[b] += (FUZ_C16(bit >> 3, b, (bit & 7) >> 1) >> (((bit & 7) & 1) * 16)) & 0xFFFFu;
                    }
                    S.us_gpos[o] = Q0 + bit;
                    S.us_tile[o] = tile;
                    reinterpret_cast<uint4 *>(S.us_cnt)[o] = make_uint4(cn[0], cn[1], cn[2], cn[3]);
                }
                o++;
            }
        }
#pragma unroll
        for (int w = 0; w < 4; w++) {
            pend4[w] = pend8[w] = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) P[w][k] = 0;
        }
        n_reads_seen = 0; groups16 = 0; sit = 0; spilled = false;
        FUZ_TR(10);
    }
    FUZ_TR(11);
#undef FUZ_C16
#undef FUZ_TR
}

// ---------------------------------------------------------------- variant_map rows
// One warp per site: the reads of the site's tile in file order, 32 at a time; a lane finds the segment of its read
// that holds the position and reads the base from SEQ; ballot/popc turn "my read carries the major / minor allele"
// into ordered row slots (phasing.py:125-128: all major-allele reads, then all minor-allele reads).
__global__ void __launch_bounds__(256) k_signature_seg(const uint8_t *__restrict__ rec_buf, const int32_t *__restrict__ rec_qid,
                                                       HetScratch S, fuz_outputs O, fuz_status *st) {
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
        const int tile = gp / FUZ_TILE;
        const int base = S.tile_ent_base[tile], cnt = S.tile_ent_cnt[tile];
        const int64_t off0 = S.site_row_off[s], off1 = off0 + n0;
        int run0 = 0, run1 = 0;
        for (int kb = 0; kb < cnt; kb += 32) {
            const int k = kb + lane;
            uint32_t nib = 0;
            int r = 0;
            if (k < cnt) {
                const int4 *ep = reinterpret_cast<const int4 *>(S.ents + base + k);
                const int4 ea = __ldg(ep), eb = __ldg(ep + 1);       // seq_off (x, y), seq_bytes, dqb | seg_src, nseg, rec, flags
                r = eb.z;
                const int4 *gs = S.segs + eb.x;
                int lo = 0, hi = eb.y;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (__ldg(&gs[mid].y) <= gp) lo = mid + 1; else hi = mid; }
                if (lo < eb.y) {
                    const int4 sg = __ldg(gs + lo);
                    if (sg.x <= gp) {
                        const int64_t so = (int64_t)(((uint64_t)(uint32_t)ea.y << 32) | (uint32_t)ea.x);
                        const int n = gp + sg.z + ea.w;
                        const uint32_t byte = rec_buf[so + (n >> 1)];
                        const uint32_t nb = (n & 1) ? (byte & 15u) : (byte >> 4);
                        nib = (nb & (nb - 1)) ? 0u : nb;
                    }
                }
            }
            const uint32_t m0 = __ballot_sync(0xffffffffu, nib == code0);
            const uint32_t m1 = __ballot_sync(0xffffffffu, nib == code1);
            if (nib == code0 || nib == code1) {
                const bool first = nib == code0;
                const int64_t i = first ? off0 + run0 + __popc(m0 & lt) : off1 + run1 + __popc(m1 & lt);
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
