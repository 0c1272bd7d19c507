// BAM ingest on the device: BGZF block inflate (RFC 1951), CRC32 check, record index and the
// record range of every contig.  Replaces the `samtools view <bam> <ctg>` pipe of reference
// falcon_unzip/phasing.py:27: the compressed file image crosses PCIe,
// everything after that stays in HBM.
//
//   k_bgzf_inflate   one warp per BGZF block (<= 64 KiB of payload).  All 32 lanes run the decode
//                    loop of fuz_inflate_core.h redundantly on identical bit buffers (uniform
//                    control flow, nothing to broadcast).  The compressed words reach the lanes as
//                    a 32-word register ring (one coalesced load per 128 B, fetched one ring ahead,
//                    handed out by shuffle); literals are staged one byte per lane and leave as
//                    32-byte sector stores; a match is copied by all lanes at once, sources taken
//                    modulo the distance so that overlapping matches need no ordering.  Then the
//                    warp re-reads its payload and checks the CRC32 of the BGZF trailer (32 partial
//                    CRCs combined with the GF(2) operators x^(8 * length)).
//   k_bam_anchor / k_bam_resolve / k_bam_fill
//                    the block_size chain of the record stream is sequential; it is cut into
//                    64 KiB regions: every region finds its first offset that looks like a record
//                    header (strong test) and walks the chain from there; one thread then threads
//                    the regions together starting from offset 0 (a region whose guess is not the
//                    point where the true chain enters it is re-walked), and the regions write
//                    their slice of rec_off.
//   k_bam_ctg_ranges record range of every reference id (records are sorted by refID).
#include <string.h>

#include "fuz_inflate_core.h"
#include "fuz_internal.cuh"

namespace {

#define FUZ_INF_WARPS 4            // warps (= BGZF blocks) per CTA
#define FUZ_BAM_REGION (1 << 16)   // bytes of record stream per chain region

struct CrcTables {
    uint32_t byte_tab[256];        // reflected CRC-32 (poly 0xEDB88320), one byte per step
    uint32_t x2n[32];              // x^(2^k) mod P
};

struct DevIO {
    const uint32_t *words; int64_t n_words;
    int64_t cbase;                 // absolute word index of cur[lane 0]
    uint32_t cur, nxt;             // words cbase + lane, cbase + 32 + lane
    int wi;                        // next word to hand out = cbase + wi
    uint8_t *ob;                   // 32-byte aligned address at or below the first payload byte
    int p0, pos, lim;              // positions relative to ob: payload start, next byte, payload end
    // Output is staged in registers, one byte per lane (lane = position & 31): `pend` holds the
    // window being filled, [pstart, pos); `pendA` the previous, complete window [a0, next multiple
    // of 32), which is stored only when the current one completes -- by then the loads that feed
    // it (match bytes fetched from earlier output) have had a whole window of decoding to return.
    int pstart, a0;
    uint32_t pend, pendA;
    int ln;

    __device__ __forceinline__ uint32_t ldw(int64_t i) const { return i < n_words ? __ldg(words + i) : 0u; }
    __device__ __forceinline__ int seek(int64_t b) {
        cbase = b >> 2;
        wi = 0;
        cur = ldw(cbase + ln);
        nxt = ldw(cbase + 32 + ln);
        return (int)(b & 3);
    }
    __device__ __forceinline__ uint32_t next_word() {
        const uint32_t w = __shfl_sync(0xffffffffu, cur, wi);
        if (++wi == 32) { wi = 0; cur = nxt; cbase += 32; nxt = ldw(cbase + 32 + ln); }
        return w;
    }
    __device__ __forceinline__ int64_t word_pos() const { return cbase + wi; }
    // bytes past the expected size are never stored; the caller sees pos != lim in the end
    __device__ __forceinline__ void store_a() {
        if (a0 >= 0) {
            const int p = (a0 & ~31) + ln;
            if (p >= a0 && p < lim) ob[p] = (uint8_t)pendA;
            a0 = -1;
        }
    }
    __device__ __forceinline__ bool rotate() {                 // the current window is complete
        store_a();
        pendA = pend; a0 = pstart; pstart = pos;
        return pos <= lim;
    }
    __device__ __forceinline__ void flush() {
        store_a();
        const int p = (pstart & ~31) + ln;
        if (p >= pstart && p < min(pos, lim)) ob[p] = (uint8_t)pend;
        pstart = pos;
    }
    __device__ __forceinline__ bool put(uint8_t b) {
        if ((pos & 31) == ln) pend = b;
        pos++;
        if ((pos & 31) == 0) return rotate();
        return true;
    }
    __device__ __forceinline__ bool copy(int len, int dist) {
        if (dist > pos - p0 || pos + len > lim) return false;
        if (dist == 1) {                           // a run of one byte (the 0xFF QUAL of PacBio BAMs)
            uint32_t v;
            if (pstart < pos) v = __shfl_sync(0xffffffffu, pend, (pos - 1) & 31);
            else if (a0 >= 0) v = __shfl_sync(0xffffffffu, pendA, 31);
            else { __syncwarp(); v = __ldcg(ob + pos - 1); }
            while (len > 0) {
                const int w = pos & 31, n = min(len, 32 - w);
                if (ln >= w && ln < w + n) pend = v;
                pos += n; len -= n;
                if ((pos & 31) == 0) rotate();
            }
            return true;
        }
        const int unstored = pos - (a0 >= 0 ? a0 : pstart);
        if (dist >= unstored + len) {              // every source byte is in memory already
            __syncwarp();
            const uint8_t *src = ob - dist;
            while (len > 0) {
                const int w = pos & 31, n = min(len, 32 - w);
                if (ln >= w && ln < w + n) pend = __ldcg(src + (pos - w + ln));
                pos += n; len -= n;
                if ((pos & 31) == 0) rotate();
            }
            return true;
        }
        // near match: sources among the staged bytes or inside the match itself
        flush();
        __syncwarp();
        const uint8_t *src = ob + pos - dist;
        uint8_t *dst = ob + pos;
        if (dist >= len) {
            for (int i = ln; i < len; i += 32) dst[i] = __ldcg(src + i);
        } else {                                   // overlapping: the pattern repeats with period dist
            for (int i = ln; i < len; i += 32) dst[i] = __ldcg(src + (i % dist));
        }
        pos += len;
        pstart = pos;
        return true;
    }
    __device__ __forceinline__ bool copy_in(int64_t src_byte, int len) {
        if (pos + len > lim) return false;
        flush();
        const uint8_t *src = reinterpret_cast<const uint8_t *>(words) + src_byte;
        for (int i = ln; i < len; i += 32) ob[pos + i] = __ldg(src + i);
        pos += len;
        pstart = pos;
        return true;
    }
    __device__ __forceinline__ int lane() const { return ln; }
    __device__ __forceinline__ int lanes() const { return 32; }
    __device__ __forceinline__ void sync() { __syncwarp(); }
};

// a(x) * b(x) mod P, reflected representation (zlib crc32.c multmodp)
__host__ __device__ inline uint32_t crc_mul(uint32_t a, uint32_t b) {
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}

__global__ void __launch_bounds__(FUZ_INF_WARPS * 32) k_bgzf_inflate(
    const uint8_t *__restrict__ comp, int64_t comp_bytes, const int64_t *__restrict__ coff, const int32_t *__restrict__ csize,
    const int64_t *__restrict__ uoff, const uint32_t *__restrict__ crc, int64_t n_blk, uint8_t *out, int64_t out_bytes,
    const __grid_constant__ CrcTables ctab, fuz_status *st) {
    fuz_pdl_enter();
    __shared__ FuzInfTables tabs[FUZ_INF_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * FUZ_INF_WARPS + warp;
    if (b >= n_blk) return;
    const int64_t c0 = coff[b], u0 = uoff[b], u1 = uoff[b + 1];
    const int32_t cs = csize[b];
    if (c0 < 0 || cs < 0 || c0 + cs > comp_bytes || u0 < 0 || u1 < u0 || u1 > out_bytes || u1 - u0 > 65536) {
        if (lane == 0) fuz_raise(st, FUZ_E_ARG, (int)b);
        return;
    }
    if (u1 == u0) return;                          // empty block (the BGZF EOF marker)
    DevIO io;
    io.words = reinterpret_cast<const uint32_t *>(comp);
    io.n_words = (comp_bytes + 3) >> 2;
    const int omis = (int)(reinterpret_cast<uintptr_t>(out + u0) & 31);
    io.ob = out + u0 - omis;
    io.p0 = io.pos = io.pstart = omis; io.lim = omis + (int)(u1 - u0); io.pend = io.pendA = 0; io.a0 = -1; io.ln = lane;
    io.cbase = 0; io.wi = 0; io.cur = io.nxt = 0;
    FuzInflate<DevIO> inf(io, tabs[warp]);
    int rc = inf.run(c0, cs);
    io.flush();
    if (rc == FUZ_INF_OK && io.pos != io.lim) rc = FUZ_INF_SIZE;
    __syncwarp();
    if (rc == FUZ_INF_OK && crc) {
        // CRC-32 of the payload: 32 contiguous pieces, then log2(32) combine steps.  The decode tables of the
        // warp are free now: their 4 KB hold the slicing-by-4 tables (one dependent lookup per word).
        uint32_t (*s_crc)[256] = reinterpret_cast<uint32_t (*)[256]>(tabs[warp].lit);
        for (int i = lane; i < 256; i += 32) {
            uint32_t c = ctab.byte_tab[i];
            s_crc[0][i] = c;
            for (int k = 1; k < 4; k++) { c = (c >> 8) ^ ctab.byte_tab[c & 0xFFu]; s_crc[k][i] = c; }
        }
        __syncwarp();
        const int n = (int)(u1 - u0);
        const int piece = (n + 31) >> 5;
        const int lo = min(lane * piece, n), hi = min(lo + piece, n);
        uint32_t c = 0xFFFFFFFFu;
        const uint8_t *q = out + u0 + lo, *qe = out + u0 + hi;
        while (q < qe && (reinterpret_cast<uintptr_t>(q) & 3)) c = s_crc[0][(c ^ __ldcg(q++)) & 0xFFu] ^ (c >> 8);
#pragma unroll 4
        for (; q + 4 <= qe; q += 4) {                // aligned words
            c ^= __ldcg(reinterpret_cast<const uint32_t *>(q));
            c = s_crc[3][c & 0xFFu] ^ s_crc[2][(c >> 8) & 0xFFu] ^ s_crc[1][(c >> 16) & 0xFFu] ^ s_crc[0][c >> 24];
        }
        while (q < qe) c = s_crc[0][(c ^ __ldcg(q++)) & 0xFFu] ^ (c >> 8);
        c ^= 0xFFFFFFFFu;
        int len = hi - lo;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t c2 = __shfl_down_sync(0xffffffffu, c, d);
            const int len2 = __shfl_down_sync(0xffffffffu, len, d);
            // crc(A || B) = crc(A) * x^(8 |B|) + crc(B)
            uint32_t op = 1u << 31;
            for (int k = 3, m = len2; m; m >>= 1, k++)
                if (m & 1) op = crc_mul(ctab.x2n[k & 31], op);
            c = crc_mul(op, c) ^ c2;
            len += len2;
        }
        if (lane == 0 && c != crc[b]) rc = FUZ_INF_CRC;
        rc = __shfl_sync(0xffffffffu, rc, 0);
    }
    if (rc != FUZ_INF_OK && lane == 0) {
        fuz_raise(st, FUZ_E_FORMAT, (int)b);
        atomicMax((int *)&st->reserved[3], rc);     // which FUZ_INF_* (diagnostics)
    }
}

// ---------------------------------------------------------------- record index
__device__ __forceinline__ uint32_t ld_u32(const uint8_t *p) { return fuz_ld_u32_un(p); }

// weak test: enough to follow the chain (what fuz_host_index_records checks)
__device__ __forceinline__ bool rec_chain_ok(const uint8_t *rec, int64_t o, int64_t n, int64_t *next) {
    if (o + 4 > n) return false;
    const int32_t bs = (int32_t)ld_u32(rec + o);
    if (bs < 32 || o + 4 + (int64_t)bs > n) return false;
    *next = o + 4 + (int64_t)bs;
    return true;
}
// strong test: the fixed core of an alignment record is consistent (SAM spec 4.2)
__device__ __forceinline__ bool rec_header_ok(const uint8_t *rec, int64_t o, int64_t n, int n_ref) {
    if (o + 36 > n) return false;
    const int32_t bs = (int32_t)ld_u32(rec + o);
    if (bs < 32 || o + 4 + (int64_t)bs > n) return false;
    const int32_t ref = (int32_t)ld_u32(rec + o + 4), pos = (int32_t)ld_u32(rec + o + 8);
    if (ref < -1 || ref >= n_ref || pos < -1) return false;
    const uint32_t w12 = ld_u32(rec + o + 12), w16 = ld_u32(rec + o + 16);
    const int l_name = w12 & 0xFF, n_cig = w16 & 0xFFFF;
    const int32_t l_seq = (int32_t)ld_u32(rec + o + 20);
    const int32_t nref = (int32_t)ld_u32(rec + o + 24), npos = (int32_t)ld_u32(rec + o + 28);
    if (l_name < 1 || l_seq < 0 || nref < -1 || nref >= n_ref || npos < -1) return false;
    if (32 + (int64_t)l_name + 4 * (int64_t)n_cig + ((int64_t)l_seq + 1) / 2 + l_seq > (int64_t)bs) return false;
    if (rec[o + 36 + l_name - 1] != 0) return false;                // read_name is NUL terminated
    if (l_name > 1 && (rec[o + 36] < 33 || rec[o + 36] > 126)) return false;
    return true;
}

struct BamIndexScratch {
    int64_t *first, *land, *entry;   // per region: first header-like offset, where its chain leaves the region
    int32_t *cnt, *base;             // records starting in the region (guess), index of the region's first record
    int64_t *n_rec;                  // [0] record count, [1] needed capacity, [2] end of the indexed records (tail start)
};

__global__ void __launch_bounds__(256) k_bam_anchor(const uint8_t *__restrict__ rec, int64_t n, int n_ref, int64_t n_reg,
                                                    BamIndexScratch S, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_reg) return;
    const int64_t a = k * FUZ_BAM_REGION, e = min(a + FUZ_BAM_REGION, n);
    int64_t c = -1;
    if (k == 0) c = 0;
    for (int64_t base = a; c < 0 && base < e; base += 32) {
        const int64_t o = base + lane;
        int64_t nx;
        bool ok = o < e && rec_header_ok(rec, o, n, n_ref);
        if (ok) ok = rec_chain_ok(rec, o, n, &nx) && (nx == n || rec_header_ok(rec, nx, n, n_ref));
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (m) c = base + __ffs(m) - 1;
    }
    if (lane) return;
    int64_t o = c, cnt = 0;
    if (c >= 0) {
        while (o < e) {
            int64_t nx;
            if (!rec_chain_ok(rec, o, n, &nx)) { o = -2; break; }
            cnt++;
            o = nx;
        }
    }
    S.first[k] = c; S.land[k] = o; S.cnt[k] = (int32_t)cnt;
}

// One warp threads the regions together, 32 at a time: if every region of a group either is
// entered exactly where it guessed or holds no record start at all, the group is accepted with
// two warp scans; otherwise lane 0 walks it region by region.
// allow_tail (windows of a long stream, fuz_bam_index_window): a record cut by the end of the buffer ends the index there
// instead of breaking the chain; S.n_rec[2] says where it starts.
__global__ void __launch_bounds__(32) k_bam_resolve(const uint8_t *__restrict__ rec, int64_t n, int64_t n_reg, int64_t cap_rec,
                                                     BamIndexScratch S, int allow_tail, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;                         // e.g. a corrupt BGZF block: nothing to index
    const int lane = threadIdx.x;
    int64_t E = 0, N = 0;                          // where the true chain enters the next group; records so far
    bool bad = false;
    int64_t tail = -1;                             // start of a record cut by the end of the buffer (allow_tail)
    for (int64_t k0 = 0; k0 < n_reg && !bad; k0 += 32) {
        const int64_t k = k0 + lane;
        const bool in = k < n_reg;
        if (tail >= 0) {                           // everything behind the tail belongs to the cut record
            if (in) { S.entry[k] = -1; S.base[k] = (int32_t)N; }
            continue;
        }
        const int64_t first = in ? S.first[k] : -1, land = in ? S.land[k] : -1;
        const int cnt = in ? S.cnt[k] : 0;
        const int64_t end = min((k + 1) * FUZ_BAM_REGION, n);
        // speculation: every region with a guess is entered at its guess and left at its landing point
        const bool has = in && first >= 0 && land >= 0;
        int64_t lmax = has ? land : -1;            // inclusive max-scan of the landing points
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, lmax, d);
            if (lane >= d) lmax = max(lmax, t);
        }
        int64_t e_in = __shfl_up_sync(0xffffffffu, lmax, 1);
        if (lane == 0) e_in = -1;
        e_in = max(e_in, E);                       // entry point of this region under the speculation
        const bool ok = !in || (has ? first == e_in : e_in >= end);
        if (__all_sync(0xffffffffu, ok)) {
            const int c = has ? cnt : 0;
            const int incl = fuz_warp_incl_scan(c, lane);
            if (in) { S.entry[k] = has ? e_in : -1; S.base[k] = (int32_t)(N + incl - c); }
            N += __shfl_sync(0xffffffffu, incl, 31);
            E = max(E, __shfl_sync(0xffffffffu, lmax, 31));
            continue;
        }
        if (lane == 0) {
            const int m = (int)min((int64_t)32, n_reg - k0);
            for (int i = 0; i < m && !bad; i++) {
                const int64_t kk = k0 + i, e2 = min((kk + 1) * FUZ_BAM_REGION, n);
                if (E >= e2 || tail >= 0) { S.entry[kk] = -1; S.base[kk] = (int32_t)N; continue; }
                S.entry[kk] = E; S.base[kk] = (int32_t)N;
                if (S.first[kk] == E && S.land[kk] >= 0) { N += S.cnt[kk]; E = S.land[kk]; continue; }
                while (E < e2) {                   // the region guessed wrong (or its chain broke): walk it here
                    int64_t nx;
                    if (!rec_chain_ok(rec, E, n, &nx)) {
                        // the record runs past the end of the buffer (or its length field does): the tail of a window
                        if (allow_tail && (E + 4 > n || (int32_t)ld_u32(rec + E) >= 32)) tail = E; else bad = true;
                        break;
                    }
                    N++;
                    E = nx;
                }
            }
        }
        E = __shfl_sync(0xffffffffu, E, 0); N = __shfl_sync(0xffffffffu, N, 0);
        tail = __shfl_sync(0xffffffffu, tail, 0);
        bad = __shfl_sync(0xffffffffu, (int)bad, 0) != 0;
    }
    if (lane == 0) {
        S.n_rec[2] = tail >= 0 ? tail : n;
        if (bad || (tail < 0 && E != n)) { fuz_raise(st, FUZ_E_BADRECORD, (int)min(N, (int64_t)0x7fffffff)); S.n_rec[0] = 0; S.n_rec[1] = 0; }
        else {
            S.n_rec[1] = N;
            if (N > cap_rec || N > 0x7fffffff) { fuz_raise(st, FUZ_E_CAPACITY, 7); S.n_rec[0] = 0; }
            else S.n_rec[0] = N;
        }
    }
}

__global__ void __launch_bounds__(256) k_bam_fill(const uint8_t *__restrict__ rec, int64_t n, int64_t n_reg, BamIndexScratch S,
                                                  int64_t *__restrict__ rec_off, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t lim = S.n_rec[2];                  // end of the indexed records (n, or the start of a cut record)
    if (k == 0) rec_off[S.n_rec[0]] = lim;
    if (k >= n_reg) return;
    int64_t o = S.entry[k], i = S.base[k];
    const int64_t end = min(min((k + 1) * FUZ_BAM_REGION, n), lim);
    while (o >= 0 && o < end) {
        rec_off[i++] = o;
        o += 4 + (int64_t)(int32_t)ld_u32(rec + o);
    }
}

// ctg_rec_off[c] = first record with refID >= c (unmapped records, refID -1, sort last and count as n_ref)
__global__ void __launch_bounds__(256) k_bam_ctg_ranges(const uint8_t *__restrict__ rec, const int64_t *__restrict__ rec_off,
                                                        const int64_t *__restrict__ n_rec_p, int n_ref, int32_t *__restrict__ ctg_rec_off,
                                                        fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t n_rec = *n_rec_p;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n_rec == 0) {
        for (int64_t c = t0; c <= n_ref; c += (int64_t)gridDim.x * blockDim.x) ctg_rec_off[c] = 0;
        return;
    }
    for (int64_t r = t0; r < n_rec; r += (int64_t)gridDim.x * blockDim.x) {
        int ref = (int32_t)ld_u32(rec + rec_off[r] + 4);
        if (ref < -1 || ref >= n_ref) { fuz_raise(st, FUZ_E_BADRECORD, (int)r); continue; }
        if (ref < 0) ref = n_ref;
        int prev = -1;
        if (r > 0) {
            prev = (int32_t)ld_u32(rec + rec_off[r - 1] + 4);
            if (prev < 0 || prev > n_ref) prev = n_ref;
            if (prev > ref) { fuz_raise(st, FUZ_E_UNSORTED, (int)r); continue; }
        }
        for (int c = prev + 1; c <= ref; c++) ctg_rec_off[c] = (int32_t)r;
        if (r == n_rec - 1)
            for (int c = ref + 1; c <= n_ref; c++) ctg_rec_off[c] = (int32_t)n_rec;
    }
}

// ---------------------------------------------------------------- record index of SEVERAL files in one buffer
// fuz_bam_index_files: the reference leaves one sorted BAM per contig (unzip.py:90).  All files are inflated by ONE
// launch of k_bgzf_inflate into one buffer; file s ("segment") has its records at raw[seg_start[s], seg_end[s]).  The
// chain regions of all segments are anchored in one launch, one warp per segment threads its regions together, and the
// mapped records of all segments are written back to back (their unmapped tails and the headers between them dropped).
struct BamFilesScratch {
    const int64_t *seg_start, *seg_end, *seg_reg0;   // [n_seg], [n_seg], [n_seg + 1] first region of the segment
    const int32_t *seg_nref, *seg_ctg0, *reg_seg;    // [n_seg], [n_seg + 1] first contig of the segment, [n_reg] segment of a region
    int32_t *seg_cnt, *seg_base;                     // records of the segment; exclusive scan [n_seg + 1]
    int32_t *loc;                                    // [n_ctg + n_seg] per segment: n_ref + 1 LOCAL record offsets of its references
    int32_t *mbase;                                  // [n_seg + 1] first mapped record of the segment in the output
    int64_t *bbase, *mbytes;                         // [n_seg + 1] its first output byte; [n_seg] bytes of its mapped records
    int64_t *tmp_off;                                // [cap_rec + 1] offsets of ALL records in raw
    int64_t *totals;                                 // [0] mapped records, [1] all records (capacity needed), [2] mapped bytes
};

__global__ void __launch_bounds__(256) k_bamf_anchor(const uint8_t *__restrict__ raw, int64_t n_reg, BamIndexScratch S, BamFilesScratch F,
                                                     const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n_reg) return;
    const int s = F.reg_seg[k];
    const int64_t n = F.seg_end[s], lk = k - F.seg_reg0[s];
    const int n_ref = F.seg_nref[s];
    const int64_t a = F.seg_start[s] + lk * FUZ_BAM_REGION, e = min(a + FUZ_BAM_REGION, n);
    int64_t c = -1;
    if (lk == 0) c = a;
    for (int64_t base = a; c < 0 && base < e; base += 32) {
        const int64_t o = base + lane;
        int64_t nx;
        bool ok = o < e && rec_header_ok(raw, o, n, n_ref);
        if (ok) ok = rec_chain_ok(raw, o, n, &nx) && (nx == n || rec_header_ok(raw, nx, n, n_ref));
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (m) c = base + __ffs(m) - 1;
    }
    if (lane) return;
    int64_t o = c, cnt = 0;
    if (c >= 0) {
        while (o < e) {
            int64_t nx;
            if (!rec_chain_ok(raw, o, n, &nx)) { o = -2; break; }
            cnt++;
            o = nx;
        }
    }
    S.first[k] = c; S.land[k] = o; S.cnt[k] = (int32_t)cnt;
}

// one warp per segment; the algorithm of k_bam_resolve on the regions of that segment (S.base = LOCAL record index)
__global__ void __launch_bounds__(128) k_bamf_resolve(const uint8_t *__restrict__ raw, int n_seg, BamIndexScratch S, BamFilesScratch F,
                                                      fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x & 31;
    const int s = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (s >= n_seg) return;
    const int64_t s0 = F.seg_start[s], n = F.seg_end[s], r0 = F.seg_reg0[s], n_reg = F.seg_reg0[s + 1] - r0;
    int64_t E = s0, N = 0;
    bool bad = false;
    for (int64_t k0 = 0; k0 < n_reg && !bad; k0 += 32) {
        const int64_t k = k0 + lane;
        const bool in = k < n_reg;
        const int64_t first = in ? S.first[r0 + k] : -1, land = in ? S.land[r0 + k] : -1;
        const int cnt = in ? S.cnt[r0 + k] : 0;
        const int64_t end = min(s0 + (k + 1) * FUZ_BAM_REGION, n);
        const bool has = in && first >= 0 && land >= 0;
        int64_t lmax = has ? land : -1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, lmax, d);
            if (lane >= d) lmax = max(lmax, t);
        }
        int64_t e_in = __shfl_up_sync(0xffffffffu, lmax, 1);
        if (lane == 0) e_in = -1;
        e_in = max(e_in, E);
        const bool ok = !in || (has ? first == e_in : e_in >= end);
        if (__all_sync(0xffffffffu, ok)) {
            const int c = has ? cnt : 0;
            const int incl = fuz_warp_incl_scan(c, lane);
            if (in) { S.entry[r0 + k] = has ? e_in : -1; S.base[r0 + k] = (int32_t)(N + incl - c); }
            N += __shfl_sync(0xffffffffu, incl, 31);
            E = max(E, __shfl_sync(0xffffffffu, lmax, 31));
            continue;
        }
        if (lane == 0) {
            const int m = (int)min((int64_t)32, n_reg - k0);
            for (int i = 0; i < m && !bad; i++) {
                const int64_t kk = k0 + i, e2 = min(s0 + (kk + 1) * FUZ_BAM_REGION, n);
                if (E >= e2) { S.entry[r0 + kk] = -1; S.base[r0 + kk] = (int32_t)N; continue; }
                S.entry[r0 + kk] = E; S.base[r0 + kk] = (int32_t)N;
                if (S.first[r0 + kk] == E && S.land[r0 + kk] >= 0) { N += S.cnt[r0 + kk]; E = S.land[r0 + kk]; continue; }
                while (E < e2) {
                    int64_t nx;
                    if (!rec_chain_ok(raw, E, n, &nx)) { bad = true; break; }
                    N++;
                    E = nx;
                }
            }
        }
        E = __shfl_sync(0xffffffffu, E, 0); N = __shfl_sync(0xffffffffu, N, 0);
        bad = __shfl_sync(0xffffffffu, (int)bad, 0) != 0;
    }
    if (lane == 0) {
        if (bad || E != n || N > 0x7fffffff) { fuz_raise(st, FUZ_E_BADRECORD, s); F.seg_cnt[s] = 0; }
        else F.seg_cnt[s] = (int32_t)N;
    }
}

__global__ void __launch_bounds__(1024) k_bamf_scan_seg(int n_seg, int64_t cap_rec, BamFilesScratch F, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;                         // uniform: read before any thread of this kernel can raise
    const long long total = fuz_cta_scan_i32(F.seg_cnt, F.seg_base, n_seg);
    if (threadIdx.x == 0) {
        F.totals[1] = total;
        if (total > cap_rec || total > 0x7fffffff) fuz_raise(st, FUZ_E_CAPACITY, 7);
    }
}

__global__ void __launch_bounds__(256) k_bamf_fill(const uint8_t *__restrict__ raw, int64_t n_reg, BamIndexScratch S, BamFilesScratch F,
                                                   const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_reg) return;
    const int s = F.reg_seg[k];
    int64_t o = S.entry[k], i = (int64_t)F.seg_base[s] + S.base[k];
    const int64_t end = min(F.seg_start[s] + (k - F.seg_reg0[s] + 1) * FUZ_BAM_REGION, F.seg_end[s]);
    while (o >= 0 && o < end) {
        F.tmp_off[i++] = o;
        o += 4 + (int64_t)(int32_t)ld_u32(raw + o);
    }
}

// per segment what k_bam_ctg_ranges does for one file, with LOCAL record indices: loc[c] = first record of the segment with
// refID >= c; loc[n_ref] = its mapped records
__global__ void __launch_bounds__(256) k_bamf_ranges(const uint8_t *__restrict__ raw, int n_seg, BamFilesScratch F, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t n_all = F.seg_base[n_seg];
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_all; r += (int64_t)gridDim.x * blockDim.x) {
        const int s = fuz_upper_bound(F.seg_base, 0, n_seg + 1, (int)r) - 1;     // the last segment starting at or before r
        const int j = (int)(r - F.seg_base[s]), n_ref = F.seg_nref[s], cnt = F.seg_cnt[s];
        int32_t *loc = F.loc + F.seg_ctg0[s] + s;
        int ref = (int32_t)ld_u32(raw + F.tmp_off[r] + 4);
        if (ref < -1 || ref >= n_ref) { fuz_raise(st, FUZ_E_BADRECORD, (int)r); continue; }
        if (ref < 0) ref = n_ref;
        int prev = -1;
        if (j > 0) {
            prev = (int32_t)ld_u32(raw + F.tmp_off[r - 1] + 4);
            if (prev < 0 || prev > n_ref) prev = n_ref;
            if (prev > ref) { fuz_raise(st, FUZ_E_UNSORTED, (int)r); continue; }
        }
        for (int c = prev + 1; c <= ref; c++) loc[c] = j;
        if (j == cnt - 1)
            for (int c = ref + 1; c <= n_ref; c++) loc[c] = cnt;
    }
}

// one warp: where the mapped records of every segment go (records and bytes), totals, closing entries of the outputs
__global__ void __launch_bounds__(32) k_bamf_scan_mapped(int n_seg, int n_ctg, int64_t cap_bytes, BamFilesScratch F,
                                                         int64_t *__restrict__ rec_off, int32_t *__restrict__ ctg_rec_off, fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int lane = threadIdx.x;
    long long R = 0, B = 0;
    for (int s0 = 0; s0 < n_seg; s0 += 32) {
        const int s = s0 + lane;
        long long m = 0, b = 0;
        if (s < n_seg) {
            const int cnt = F.seg_cnt[s];
            m = F.loc[F.seg_ctg0[s] + s + F.seg_nref[s]];
            b = (m < cnt ? F.tmp_off[F.seg_base[s] + m] : F.seg_end[s]) - F.seg_start[s];
            F.mbytes[s] = b;
        }
        long long mi = m, bi = b;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const long long t = __shfl_up_sync(0xffffffffu, mi, d), u = __shfl_up_sync(0xffffffffu, bi, d);
            if (lane >= d) { mi += t; bi += u; }
        }
        if (s < n_seg) { F.mbase[s] = (int32_t)(R + mi - m); F.bbase[s] = B + bi - b; }
        R += __shfl_sync(0xffffffffu, mi, 31);
        B += __shfl_sync(0xffffffffu, bi, 31);
    }
    if (lane == 0) {
        F.mbase[n_seg] = (int32_t)R; F.bbase[n_seg] = B;
        F.totals[0] = R; F.totals[2] = B;
        if (B > cap_bytes) fuz_raise(st, FUZ_E_CAPACITY, 8);
        else { rec_off[R] = B; ctg_rec_off[n_ctg] = (int32_t)R; }
    }
}

__global__ void __launch_bounds__(256) k_bamf_finalize(int n_seg, int n_ctg, BamFilesScratch F, int64_t *__restrict__ rec_off,
                                                       int32_t *__restrict__ ctg_rec_off, const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t n_all = F.seg_base[n_seg], t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
    for (int64_t r = t0; r < n_all; r += nt) {
        const int s = fuz_upper_bound(F.seg_base, 0, n_seg + 1, (int)r) - 1;
        const int j = (int)(r - F.seg_base[s]);
        if (j < F.loc[F.seg_ctg0[s] + s + F.seg_nref[s]]) rec_off[F.mbase[s] + j] = F.bbase[s] + (F.tmp_off[r] - F.seg_start[s]);
    }
    for (int64_t g = t0; g < n_ctg; g += nt) {
        const int s = fuz_upper_bound(F.seg_ctg0, 0, n_seg + 1, (int)g) - 1;
        ctg_rec_off[g] = F.mbase[s] + F.loc[F.seg_ctg0[s] + s + ((int)g - F.seg_ctg0[s])];
    }
}

// n bytes src -> dst by nt cooperating threads (t = my index): 128-bit stores, the source realigned from 32-bit words;
// reads up to 7 bytes past src + n (callers keep slack behind their buffers)
__device__ __forceinline__ void copy_bytes(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t n, int t, int nt) {
    const int64_t head = min(n, (int64_t)((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15));
    const int64_t body = (n - head) >> 4;
    for (int64_t i = t; i < head; i += nt) dst[i] = src[i];
    for (int64_t q = t; q < body; q += nt) {
        const uintptr_t p = reinterpret_cast<uintptr_t>(src + head + 16 * q);
        const uint32_t *w = reinterpret_cast<const uint32_t *>(p & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(p & 3) * 8;
        const uint32_t x0 = __ldg(w), x1 = __ldg(w + 1), x2 = __ldg(w + 2), x3 = __ldg(w + 3), x4 = __ldg(w + 4);
        uint4 v;
        v.x = __funnelshift_r(x0, x1, sh); v.y = __funnelshift_r(x1, x2, sh);
        v.z = __funnelshift_r(x2, x3, sh); v.w = __funnelshift_r(x3, x4, sh);
        *reinterpret_cast<uint4 *>(dst + head + 16 * q) = v;
    }
    for (int64_t i = head + 16 * body + t; i < n; i += nt) dst[i] = src[i];
}

// one CTA per region: its share of the segment's mapped bytes to their place in the output
__global__ void __launch_bounds__(256) k_bamf_copy(const uint8_t *__restrict__ raw, uint8_t *__restrict__ out, BamFilesScratch F,
                                                   const fuz_status *st) {
    fuz_pdl_enter();
    if (st->error) return;
    const int64_t k = blockIdx.x;
    const int s = F.reg_seg[k];
    const int64_t s0 = F.seg_start[s];
    const int64_t a = s0 + (k - F.seg_reg0[s]) * FUZ_BAM_REGION, e = min(a + FUZ_BAM_REGION, s0 + F.mbytes[s]);
    if (a >= e) return;
    copy_bytes(out + F.bbase[s] + (a - s0), raw + a, e - a, threadIdx.x, blockDim.x);
}

// fuz_gather_records: output record i = source record sel[i], one warp per record (a stable selection of whole records by
// any key: the per-contig partition of reference falcon_unzip/select_reads_from_bam.py:70-86)
__global__ void __launch_bounds__(256) k_gather_records(const uint8_t *__restrict__ src, const int64_t *__restrict__ src_off, int64_t n_src,
                                                        int64_t src_bytes, const int64_t *__restrict__ sel, const int64_t *__restrict__ dst_off,
                                                        int64_t m, uint8_t *__restrict__ dst, int64_t dst_bytes, fuz_status *st) {
    fuz_pdl_enter();
    const int lane = threadIdx.x & 31;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < m; i += n_warps) {
        const int64_t r = sel[i];
        if (r < 0 || r >= n_src) { if (lane == 0) fuz_raise(st, FUZ_E_ARG, (int)min(i, (int64_t)0x7fffffff)); continue; }
        const int64_t a = src_off[r], n = src_off[r + 1] - a, d = dst_off[i];
        if (a < 0 || n < 0 || a + n > src_bytes || d < 0 || d + n > dst_bytes || dst_off[i + 1] - d != n) {
            if (lane == 0) fuz_raise(st, FUZ_E_ARG, (int)min(i, (int64_t)0x7fffffff));
            continue;
        }
        copy_bytes(dst + d, src + a, n, lane, 32);
    }
}

}  // namespace

// ---------------------------------------------------------------- host side
extern "C" int64_t fuz_host_bgzf_index(const uint8_t *h_file, int64_t n_bytes, int64_t cap, int64_t *h_coff, int32_t *h_csize,
                                       int64_t *h_uoff, uint32_t *h_crc) {
    if (!h_file || n_bytes < 0) return -1;
    int64_t o = 0, n = 0, u = 0;
    while (o < n_bytes) {
        if (o + 18 > n_bytes || h_file[o] != 0x1f || h_file[o + 1] != 0x8b || h_file[o + 2] != 8 || !(h_file[o + 3] & 4)) return -1;
        const int xlen = h_file[o + 10] | (h_file[o + 11] << 8);
        int64_t x = o + 12;
        const int64_t xend = x + xlen;
        if (xend > n_bytes) return -1;
        int64_t bsize = -1;
        while (x + 4 <= xend) {
            const int slen = h_file[x + 2] | (h_file[x + 3] << 8);
            if (h_file[x] == 66 && h_file[x + 1] == 67 && slen == 2 && x + 6 <= xend) bsize = (h_file[x + 4] | (h_file[x + 5] << 8)) + 1;
            x += 4 + slen;
        }
        if (bsize < 0 || o + bsize > n_bytes || xend + 8 > o + bsize) return -1;
        uint32_t crc, isize;
        memcpy(&crc, h_file + o + bsize - 8, 4);
        memcpy(&isize, h_file + o + bsize - 4, 4);
        if (isize > 65536) return -1;
        if (h_coff) {
            if (n >= cap) return -1;
            h_coff[n] = xend; h_csize[n] = (int32_t)(o + bsize - 8 - xend); h_uoff[n] = u;
            if (h_crc) h_crc[n] = crc;
        }
        u += isize;
        n++;
        o += bsize;
    }
    if (h_uoff && n <= cap) h_uoff[n] = u;
    return n;
}

static CrcTables make_crc_tables() {
    CrcTables t;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
        t.byte_tab[i] = c;
    }
    uint32_t p = 1u << 30;                         // x^1
    t.x2n[0] = p;
    for (int k = 1; k < 32; k++) t.x2n[k] = p = crc_mul(p, p);
    return t;
}

extern "C" int fuz_bgzf_inflate(fuz_ctx *ctx, const uint8_t *d_comp, int64_t comp_bytes, const int64_t *d_coff,
                                const int32_t *d_csize, const int64_t *d_uoff, const uint32_t *d_crc, int64_t n_blk,
                                uint8_t *d_out, int64_t out_bytes) {
    if (!ctx || !d_comp || !d_coff || !d_csize || !d_uoff || !d_out || n_blk < 0 || comp_bytes < 0 || out_bytes < 0)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_bgzf_inflate: bad argument");
    if (reinterpret_cast<uintptr_t>(d_comp) & 3) return fuz_fail(ctx, FUZ_E_ARG, "fuz_bgzf_inflate: d_comp must be 4-byte aligned");
    static const CrcTables tabs = make_crc_tables();
    FUZ_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(fuz_status), ctx->stream));
    ctx->ingest_pending = true;                    // fuz_bam_index_records reports an error raised here
    if (n_blk == 0) return FUZ_OK;
    const unsigned grid = (unsigned)((n_blk + FUZ_INF_WARPS - 1) / FUZ_INF_WARPS);
    fuz_launch(ctx, k_bgzf_inflate, grid, FUZ_INF_WARPS * 32, 0, ctx->stream, d_comp, comp_bytes, d_coff, d_csize, d_uoff, d_crc, n_blk,
               d_out, out_bytes, tabs, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bgzf_inflate");
    return FUZ_OK;
}

static int bam_index_impl(fuz_ctx *ctx, const uint8_t *d_rec, int64_t rec_bytes, int32_t n_ref, int64_t cap_rec,
                          int64_t *d_rec_off, int32_t *d_ctg_rec_off, int64_t *h_n_rec, int64_t *h_need_rec, int64_t *h_tail) {
    if (!ctx || !d_rec || rec_bytes < 0 || n_ref < 0 || cap_rec < 0 || !d_rec_off || !d_ctg_rec_off || !h_n_rec)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_bam_index_records: bad argument");
    cudaStream_t st = ctx->stream;
    const int64_t n_reg = (rec_bytes + FUZ_BAM_REGION - 1) / FUZ_BAM_REGION;
    FuzLayout L;
    const size_t o_first = L.add(8 * (size_t)(n_reg + 1)), o_land = L.add(8 * (size_t)(n_reg + 1)), o_entry = L.add(8 * (size_t)(n_reg + 1));
    const size_t o_cnt = L.add(4 * (size_t)(n_reg + 1)), o_base = L.add(4 * (size_t)(n_reg + 1)), o_n = L.add(32);
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    BamIndexScratch S;
    S.first = fuz_at<int64_t>(ctx, o_first); S.land = fuz_at<int64_t>(ctx, o_land); S.entry = fuz_at<int64_t>(ctx, o_entry);
    S.cnt = fuz_at<int32_t>(ctx, o_cnt); S.base = fuz_at<int32_t>(ctx, o_base); S.n_rec = fuz_at<int64_t>(ctx, o_n);
    if (!ctx->ingest_pending) FUZ_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(fuz_status), st));
    ctx->ingest_pending = false;
    FUZ_CUDA(ctx, cudaMemsetAsync(S.n_rec, 0, 32, st));
    if (n_reg > 0) {
        fuz_launch(ctx, k_bam_anchor, (unsigned)((n_reg * 32 + 255) / 256), 256, 0, st, d_rec, rec_bytes, (int)n_ref, n_reg, S, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_bam_anchor");
    }
    fuz_launch(ctx, k_bam_resolve, 1, 32, 0, st, d_rec, rec_bytes, n_reg, cap_rec, S, h_tail ? 1 : 0, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bam_resolve");
    fuz_launch(ctx, k_bam_fill, (unsigned)((n_reg + 256) / 256), 256, 0, st, d_rec, rec_bytes, n_reg, S, d_rec_off, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bam_fill");
    fuz_launch(ctx, k_bam_ctg_ranges, FUZ_GRID_BLOCKS, 256, 0, st, d_rec, d_rec_off, S.n_rec, (int)n_ref, d_ctg_rec_off, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bam_ctg_ranges");
    int64_t h[4] = {0, 0, 0, 0};
    FUZ_CUDA(ctx, cudaMemcpyAsync(h, S.n_rec, 32, cudaMemcpyDeviceToHost, st));
    fuz_status hs;
    rc = fuz_get_status(ctx, &hs);                 // synchronises
    *h_n_rec = h[0];
    if (h_need_rec) *h_need_rec = h[1];
    if (h_tail) *h_tail = h[2];
    return rc;
}

extern "C" int fuz_bam_index_records(fuz_ctx *ctx, const uint8_t *d_rec, int64_t rec_bytes, int32_t n_ref, int64_t cap_rec,
                                     int64_t *d_rec_off, int32_t *d_ctg_rec_off, int64_t *h_n_rec, int64_t *h_need_rec) {
    return bam_index_impl(ctx, d_rec, rec_bytes, n_ref, cap_rec, d_rec_off, d_ctg_rec_off, h_n_rec, h_need_rec, nullptr);
}

extern "C" int fuz_bam_index_window(fuz_ctx *ctx, const uint8_t *d_rec, int64_t rec_bytes, int32_t n_ref, int64_t cap_rec,
                                    int64_t *d_rec_off, int32_t *d_ctg_rec_off, int64_t *h_n_rec, int64_t *h_need_rec, int64_t *h_tail) {
    if (!h_tail) return fuz_fail(ctx, FUZ_E_ARG, "fuz_bam_index_window: h_tail is NULL");
    return bam_index_impl(ctx, d_rec, rec_bytes, n_ref, cap_rec, d_rec_off, d_ctg_rec_off, h_n_rec, h_need_rec, h_tail);
}

extern "C" int fuz_bam_index_files(fuz_ctx *ctx, const uint8_t *d_raw, int64_t raw_bytes, int32_t n_seg, const int64_t *h_seg_start,
                                   const int64_t *h_seg_end, const int32_t *h_seg_nref, int64_t cap_rec, uint8_t *d_rec_out,
                                   int64_t cap_bytes, int64_t *d_rec_off, int32_t *d_ctg_rec_off, int64_t *h_n_rec,
                                   int64_t *h_need_rec, int64_t *h_rec_bytes) {
    if (!ctx || !d_raw || raw_bytes < 0 || n_seg < 1 || !h_seg_start || !h_seg_end || !h_seg_nref || cap_rec < 0 || !d_rec_out ||
        cap_bytes < 0 || !d_rec_off || !d_ctg_rec_off || !h_n_rec)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_bam_index_files: bad argument");
    if (raw_bytes / 36 > 0x7fffffff) return fuz_fail(ctx, FUZ_E_ARG, "fuz_bam_index_files: more than 2^31 records possible; split the batch");
    // host tables: segments, their regions and contigs
    std::vector<int64_t> reg0((size_t)n_seg + 1, 0);
    std::vector<int32_t> ctg0((size_t)n_seg + 1, 0);
    int64_t prev_end = 0;
    for (int32_t s = 0; s < n_seg; s++) {
        if (h_seg_start[s] < prev_end || h_seg_end[s] < h_seg_start[s] || h_seg_end[s] > raw_bytes || h_seg_nref[s] < 0)
            return fuz_fail(ctx, FUZ_E_ARG, "fuz_bam_index_files: segment %d is not inside the buffer / not in ascending order", (int)s);
        prev_end = h_seg_end[s];
        reg0[s + 1] = reg0[s] + (h_seg_end[s] - h_seg_start[s] + FUZ_BAM_REGION - 1) / FUZ_BAM_REGION;
        const int64_t c = (int64_t)ctg0[s] + h_seg_nref[s];
        if (c > 0x7fffffff - n_seg) return fuz_fail(ctx, FUZ_E_ARG, "fuz_bam_index_files: too many references");
        ctg0[s + 1] = (int32_t)c;
    }
    const int64_t n_reg = reg0[n_seg];
    const int32_t n_ctg = ctg0[n_seg];
    std::vector<int32_t> reg_seg((size_t)n_reg);
    for (int32_t s = 0; s < n_seg; s++)
        for (int64_t k = reg0[s]; k < reg0[s + 1]; k++) reg_seg[(size_t)k] = s;

    cudaStream_t st = ctx->stream;
    FuzLayout L;
    const size_t ns = (size_t)n_seg, nr = (size_t)n_reg;
    // uploaded tables first (one copy)
    const size_t o_start = L.add(8 * ns), o_end = L.add(8 * ns), o_reg0 = L.add(8 * (ns + 1));
    const size_t o_nref = L.add(4 * ns), o_ctg0 = L.add(4 * (ns + 1)), o_regseg = L.add(4 * (nr + 1));
    const size_t tab_bytes = L.off;
    const size_t o_first = L.add(8 * (nr + 1)), o_land = L.add(8 * (nr + 1)), o_entry = L.add(8 * (nr + 1));
    const size_t o_cnt = L.add(4 * (nr + 1)), o_base = L.add(4 * (nr + 1));
    const size_t o_scnt = L.add(4 * (ns + 4)), o_sbase = L.add(4 * (ns + 4));
    const size_t o_loc = L.add(4 * ((size_t)n_ctg + ns)), o_mbase = L.add(4 * (ns + 1));
    const size_t o_bbase = L.add(8 * (ns + 1)), o_mbytes = L.add(8 * ns), o_tmp = L.add(8 * ((size_t)cap_rec + 1)), o_tot = L.add(32);
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    std::vector<uint8_t> tab(tab_bytes, 0);
    memcpy(tab.data() + o_start, h_seg_start, 8 * ns);
    memcpy(tab.data() + o_end, h_seg_end, 8 * ns);
    memcpy(tab.data() + o_reg0, reg0.data(), 8 * (ns + 1));
    memcpy(tab.data() + o_nref, h_seg_nref, 4 * ns);
    memcpy(tab.data() + o_ctg0, ctg0.data(), 4 * (ns + 1));
    if (nr) memcpy(tab.data() + o_regseg, reg_seg.data(), 4 * nr);
    FUZ_CUDA(ctx, cudaMemcpyAsync(ctx->arena, tab.data(), tab_bytes, cudaMemcpyHostToDevice, st));   // pageable source: staged before the call returns

    BamIndexScratch S;
    S.first = fuz_at<int64_t>(ctx, o_first); S.land = fuz_at<int64_t>(ctx, o_land); S.entry = fuz_at<int64_t>(ctx, o_entry);
    S.cnt = fuz_at<int32_t>(ctx, o_cnt); S.base = fuz_at<int32_t>(ctx, o_base); S.n_rec = fuz_at<int64_t>(ctx, o_tot);
    BamFilesScratch F;
    F.seg_start = fuz_at<int64_t>(ctx, o_start); F.seg_end = fuz_at<int64_t>(ctx, o_end); F.seg_reg0 = fuz_at<int64_t>(ctx, o_reg0);
    F.seg_nref = fuz_at<int32_t>(ctx, o_nref); F.seg_ctg0 = fuz_at<int32_t>(ctx, o_ctg0); F.reg_seg = fuz_at<int32_t>(ctx, o_regseg);
    F.seg_cnt = fuz_at<int32_t>(ctx, o_scnt); F.seg_base = fuz_at<int32_t>(ctx, o_sbase);
    F.loc = fuz_at<int32_t>(ctx, o_loc); F.mbase = fuz_at<int32_t>(ctx, o_mbase);
    F.bbase = fuz_at<int64_t>(ctx, o_bbase); F.mbytes = fuz_at<int64_t>(ctx, o_mbytes);
    F.tmp_off = fuz_at<int64_t>(ctx, o_tmp); F.totals = fuz_at<int64_t>(ctx, o_tot);
    if (!ctx->ingest_pending) FUZ_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(fuz_status), st));
    ctx->ingest_pending = false;
    FUZ_CUDA(ctx, cudaMemsetAsync(F.totals, 0, 32, st));
    FUZ_CUDA(ctx, cudaMemsetAsync(F.loc, 0, 4 * ((size_t)n_ctg + ns), st));
    FUZ_CUDA(ctx, cudaMemsetAsync(F.seg_cnt, 0, 4 * (ns + 4), st));
    if (n_reg > 0) {
        fuz_launch(ctx, k_bamf_anchor, (unsigned)((n_reg * 32 + 255) / 256), 256, 0, st, d_raw, n_reg, S, F, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_bamf_anchor");
    }
    fuz_launch(ctx, k_bamf_resolve, (unsigned)(((int64_t)n_seg * 32 + 127) / 128), 128, 0, st, d_raw, (int)n_seg, S, F, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bamf_resolve");
    fuz_launch(ctx, k_bamf_scan_seg, 1, 1024, 0, st, (int)n_seg, cap_rec, F, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bamf_scan_seg");
    if (n_reg > 0) {
        fuz_launch(ctx, k_bamf_fill, (unsigned)((n_reg + 255) / 256), 256, 0, st, d_raw, n_reg, S, F, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_bamf_fill");
    }
    fuz_launch(ctx, k_bamf_ranges, FUZ_GRID_BLOCKS, 256, 0, st, d_raw, (int)n_seg, F, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bamf_ranges");
    fuz_launch(ctx, k_bamf_scan_mapped, 1, 32, 0, st, (int)n_seg, (int)n_ctg, cap_bytes, F, d_rec_off, d_ctg_rec_off, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bamf_scan_mapped");
    fuz_launch(ctx, k_bamf_finalize, FUZ_GRID_BLOCKS, 256, 0, st, (int)n_seg, (int)n_ctg, F, d_rec_off, d_ctg_rec_off, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_bamf_finalize");
    if (n_reg > 0) {
        fuz_launch(ctx, k_bamf_copy, (unsigned)n_reg, 256, 0, st, d_raw, d_rec_out, F, ctx->d_status);
        FUZ_LAUNCH_CHECK(ctx, "k_bamf_copy");
    }
    int64_t h[4] = {0, 0, 0, 0};
    FUZ_CUDA(ctx, cudaMemcpyAsync(h, F.totals, 32, cudaMemcpyDeviceToHost, st));
    fuz_status hs;
    rc = fuz_get_status(ctx, &hs);                 // synchronises
    *h_n_rec = h[0];
    if (h_need_rec) *h_need_rec = h[1];
    if (h_rec_bytes) *h_rec_bytes = h[2];
    return rc;
}

extern "C" int fuz_gather_records(fuz_ctx *ctx, const uint8_t *d_src, const int64_t *d_src_off, int64_t n_src, int64_t src_bytes,
                                  const int64_t *d_sel, const int64_t *d_dst_off, int64_t m, uint8_t *d_dst, int64_t dst_bytes) {
    if (!ctx || !d_src || !d_src_off || n_src < 0 || src_bytes < 0 || m < 0 || dst_bytes < 0 || (m > 0 && (!d_sel || !d_dst_off || !d_dst)))
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_gather_records: bad argument");
    FUZ_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(fuz_status), ctx->stream));
    if (m == 0) return FUZ_OK;
    const unsigned grid = (unsigned)min((int64_t)FUZ_GRID_BLOCKS * 2, (m + 7) / 8);
    fuz_launch(ctx, k_gather_records, grid, 256, 0, ctx->stream, d_src, d_src_off, n_src, src_bytes, d_sel, d_dst_off, m, d_dst, dst_bytes,
               ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_gather_records");
    return FUZ_OK;
}
