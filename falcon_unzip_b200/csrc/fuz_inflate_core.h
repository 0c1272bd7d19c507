// RFC 1951 (DEFLATE) decoder used by the BGZF kernel of fuz_bgzf.cu: one warp per BGZF block,
// every lane runs the (uniform) decode loop redundantly, so no result has to be broadcast; only
// the table build, the literal staging and the match copies are split over the lanes.  Those
// parts live in the IO policy; everything here is lane agnostic, which lets oracle/inflate_model.cpp
// instantiate the same code with a scalar policy and check it against zlib on the CPU.
//
// Replaces the BGZF inflate inside `samtools view` (reference falcon_unzip/phasing.py:27); the
// format is pinned by RFC 1951/1952 and SAM spec section 4.1, the checker is zlib.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FUZ_HD __host__ __device__ __forceinline__
#else
#define FUZ_HD inline
#endif

#ifndef FUZ_INF_LBITS
#define FUZ_INF_LBITS 10           // literal/length codes up to this length decode with one table lookup
#endif
#define FUZ_INF_DBITS 7            // same for distance codes (and all code-length codes, <= 7 bits)
#define FUZ_INF_MAXL 288
#define FUZ_INF_MAXD 32

enum { FUZ_INF_OK = 0, FUZ_INF_BADBLOCK = 1, FUZ_INF_BADTABLE = 2, FUZ_INF_BADCODE = 3, FUZ_INF_BADDIST = 4,
       FUZ_INF_OVERRUN = 5, FUZ_INF_INPUT = 6, FUZ_INF_SIZE = 7, FUZ_INF_CRC = 8 };

// Table entry.  Literal: bit 31 clear, [3:0] code length (>= 1), [23:16] the byte.  Everything
// else has bit 31 set: [3:0] code length (0: not in the table, take the canonical search), [5:4]
// kind (1 match length, 2 end of block), [11:8] extra bits, [30:16] base value.  Distance and
// code-length tables use the same fields without bit 31.
struct FuzInfTables {
    uint32_t lit[1 << FUZ_INF_LBITS];
    uint32_t dist[1 << FUZ_INF_DBITS];
    uint16_t lsorted[FUZ_INF_MAXL];      // symbols ordered by (code length, symbol): canonical order
    uint16_t dsorted[FUZ_INF_MAXD];
    uint16_t lcount[16], dcount[16];     // codes per length
    uint8_t lens[FUZ_INF_MAXL + FUZ_INF_MAXD];
};
#define FUZ_INF_SPECIAL 0x80000000u

FUZ_HD uint32_t fuz_inf_lit_entry(int sym, int nbits) {
    if (sym < 256) return ((uint32_t)sym << 16) | (uint32_t)nbits;
    if (sym == 256) return FUZ_INF_SPECIAL | (2u << 4) | (uint32_t)nbits;
    if (sym > 285) return FUZ_INF_SPECIAL;                       // 286, 287: never valid in a stream (length 0 = reject)
    int base, eb;
    if (sym < 265) { base = sym - 254; eb = 0; }
    else if (sym == 285) { base = 258; eb = 0; }
    else { eb = (sym - 261) >> 2; base = 3 + ((4 + ((sym - 261) & 3)) << eb); }
    return FUZ_INF_SPECIAL | ((uint32_t)base << 16) | ((uint32_t)eb << 8) | (1u << 4) | (uint32_t)nbits;
}
FUZ_HD uint32_t fuz_inf_dist_entry(int sym, int nbits) {
    if (sym > 29) return 0;
    int base, eb;
    if (sym < 4) { base = sym + 1; eb = 0; }
    else { eb = (sym >> 1) - 1; base = 1 + ((2 + (sym & 1)) << eb); }
    return ((uint32_t)base << 16) | ((uint32_t)eb << 8) | (uint32_t)nbits;
}
FUZ_HD uint32_t fuz_inf_clen_entry(int sym, int nbits) { return ((uint32_t)sym << 16) | (uint32_t)nbits; }

FUZ_HD uint32_t fuz_inf_rev(uint32_t code, int n) {            // reverse the low n bits
    uint32_t r = 0;
    for (int i = 0; i < n; i++) { r = (r << 1) | (code & 1u); code >>= 1; }
    return r;
}

// IO policy (see DevIO in fuz_bgzf.cu, HostIO in oracle/inflate_model.cpp):
//   int seek(int64 byte)        position the reader on the aligned word holding `byte`, return byte & 3
//   uint32 next_word()          next 32 input bits
//   int64 word_pos()            words handed out so far, as an absolute word index
//   bool put(uint8)             append a literal (false: past the expected size; may be reported late,
//                               but never later than 32 literals and never after writing out of bounds)
//   bool copy(int len,int dist) append a match (false: distance too far back / past the expected size)
//   bool copy_in(int64 byte, int len)   append len input bytes starting at `byte` (stored block)
//   int lane(), int lanes(), void sync()
template <class IO>
struct FuzInflate {
    IO &io;
    FuzInfTables &T;
    // bit reader: the stream bits [32 W + bp, 32 W + 64) sit in (lo, hi), the following word in `ahead`,
    // W = io.word_pos() - 3.  refill() keeps bp < 32, so peek() always returns 32 valid bits with one
    // funnel shift; the word a refill fetches is not needed before the NEXT refill (its latency -- a
    // shuffle on the device -- stays off the decode chain).
    uint32_t lo = 0, hi = 0, ahead = 0;
    int bp = 0;
    int64_t end_byte = 0;

    FUZ_HD FuzInflate(IO &io_, FuzInfTables &t_) : io(io_), T(t_) {}

    FUZ_HD void start(int64_t first_byte, int64_t n_bytes) {
        const int mis = io.seek(first_byte);
        lo = io.next_word();
        hi = io.next_word();
        ahead = io.next_word();
        bp = 8 * mis;
        end_byte = first_byte + n_bytes;
    }
    FUZ_HD void refill() {
        if (bp >= 32) { lo = hi; hi = ahead; ahead = io.next_word(); bp -= 32; }
    }
    FUZ_HD uint32_t peek() const {                               // needs bp < 32
#ifdef __CUDA_ARCH__
        return __funnelshift_r(lo, hi, (uint32_t)bp);
#else
        return (uint32_t)((((uint64_t)hi << 32) | lo) >> bp);
#endif
    }
    FUZ_HD uint32_t take(int n) {                                // n <= 16
        refill();
        const uint32_t v = peek() & ((1u << n) - 1u);
        bp += n;
        return v;
    }
    // first input byte no bit of which has been consumed (call with bp on a byte boundary)
    FUZ_HD int64_t byte_pos() const { return (io.word_pos() - 3) * 4 + (bp >> 3); }
    FUZ_HD bool input_ok() const { return (io.word_pos() - 3) * 4 + ((bp + 7) >> 3) <= end_byte; }

    // Canonical Huffman decode tables from code lengths lens[0..n): the lookup table of 1 << tbits
    // entries for codes up to tbits long, plus count[] and sorted[] for the canonical search.
    // kind: 0 code lengths, 1 literal/length, 2 distance.  Same acceptance rules as zlib's
    // inflate_table: over-subscribed sets are rejected, incomplete ones too unless it is a single
    // one-bit literal/length or distance code.
    FUZ_HD int build(const uint8_t *lens, int n, uint32_t *table, int tbits, uint16_t *count, uint16_t *sorted, int kind) {
        int err = FUZ_INF_OK;
        if (io.lane() == 0) {
            uint16_t offs[16];
            for (int l = 0; l < 16; l++) count[l] = 0;
            for (int s = 0; s < n; s++) count[lens[s]]++;
            int left = 1, maxl = 0;
            for (int l = 1; l < 16; l++) {
                left <<= 1;
                left -= count[l];
                if (left < 0) err = FUZ_INF_BADTABLE;
                if (count[l]) maxl = l;
            }
            if (left > 0 && maxl > 0 && (kind == 0 || maxl != 1)) err = FUZ_INF_BADTABLE;
            offs[1] = 0;
            for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
            for (int s = 0; s < n; s++)
                if (lens[s]) sorted[offs[lens[s]]++] = (uint16_t)s;
            count[0] = (uint16_t)err;                          // slot 0 is unused by the decoder: carries the verdict
        }
        io.sync();
        err = count[0];
        for (int i = io.lane(); i < (1 << tbits); i += io.lanes()) table[i] = kind == 1 ? FUZ_INF_SPECIAL : 0u;
        io.sync();
        if (err) return err;
        int n_used = 0;
        for (int l = 1; l < 16; l++) n_used += count[l];
        for (int j = io.lane(); j < n_used; j += io.lanes()) {
            int l = 1, before = 0;
            uint32_t first = 0;                                 // first code of length l
            while (before + count[l] <= j) { before += count[l]; first = (first + count[l]) << 1; l++; }
            if (l > tbits) continue;
            const uint32_t code = first + (uint32_t)(j - before);
            const int sym = sorted[j];
            const uint32_t e = kind == 1 ? fuz_inf_lit_entry(sym, l) : kind == 2 ? fuz_inf_dist_entry(sym, l) : fuz_inf_clen_entry(sym, l);
            for (uint32_t i = fuz_inf_rev(code, l); i < (1u << tbits); i += 1u << l) table[i] = e;
        }
        io.sync();
        return FUZ_INF_OK;
    }

    // canonical search (codes longer than the table, and invalid codes): bit by bit like puff.c
    FUZ_HD uint32_t search(const uint16_t *count, const uint16_t *sorted, int kind) const {
        uint32_t code = 0, first = 0, index = 0;
        uint32_t b = peek();
        for (int l = 1; l < 16; l++) {
            code |= (uint32_t)b & 1u;
            b >>= 1;
            const uint32_t c = count[l];
            if (code < first + c) {
                const int sym = sorted[index + (code - first)];
                return kind == 1 ? fuz_inf_lit_entry(sym, l) : fuz_inf_dist_entry(sym, l);
            }
            index += c;
            first = (first + c) << 1;
            code <<= 1;
        }
        return 0;
    }

    FUZ_HD int fixed_tables() {
        for (int s = io.lane(); s < FUZ_INF_MAXL; s += io.lanes()) T.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
        for (int s = io.lane(); s < 32; s += io.lanes()) T.lens[FUZ_INF_MAXL + s] = 5;   // 30, 31: invalid entries
        io.sync();
        int e = build(T.lens, FUZ_INF_MAXL, T.lit, FUZ_INF_LBITS, T.lcount, T.lsorted, 1);
        if (e) return e;
        return build(T.lens + FUZ_INF_MAXL, 32, T.dist, FUZ_INF_DBITS, T.dcount, T.dsorted, 2);
    }

    FUZ_HD int dynamic_tables() {
        const int nlen = (int)take(5) + 257, ndist = (int)take(5) + 1, ncode = (int)take(4) + 4;
        if (nlen > 286 || ndist > 30) return FUZ_INF_BADTABLE;
        const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
        // every lane writes the same values to the same places: no hand-over needed before the build
        for (int i = 0; i < 19; i++) T.lens[i] = 0;
        for (int i = 0; i < ncode; i++) T.lens[order[i]] = (uint8_t)take(3);
        io.sync();
        int e = build(T.lens, 19, T.dist, FUZ_INF_DBITS, T.dcount, T.dsorted, 0);
        if (e) return e;
        // the code lengths of both alphabets; the table of the code-length code sits where the
        // distance table is built afterwards
        int i = 0, prev = 0;
        while (i < nlen + ndist) {
            refill();
            const uint32_t ent = T.dist[peek() & ((1u << FUZ_INF_DBITS) - 1u)];
            if ((ent & 15u) == 0) return FUZ_INF_BADCODE;
            bp += (int)(ent & 15u);
            const int sym = (int)(ent >> 16);
            int rep, val;
            if (sym < 16) { rep = 1; val = sym; prev = sym; }
            else if (sym == 16) { if (i == 0) return FUZ_INF_BADTABLE; rep = 3 + (int)take(2); val = prev; }
            else if (sym == 17) { rep = 3 + (int)take(3); val = 0; prev = 0; }
            else { rep = 11 + (int)take(7); val = 0; prev = 0; }
            if (i + rep > nlen + ndist) return FUZ_INF_BADTABLE;
            for (int k = 0; k < rep; k++, i++) T.lens[i < nlen ? i : FUZ_INF_MAXL + (i - nlen)] = (uint8_t)val;
        }
        if (T.lens[256] == 0) return FUZ_INF_BADTABLE;          // no end-of-block code
        io.sync();
        e = build(T.lens, nlen, T.lit, FUZ_INF_LBITS, T.lcount, T.lsorted, 1);
        if (e) return e;
        return build(T.lens + FUZ_INF_MAXL, ndist, T.dist, FUZ_INF_DBITS, T.dcount, T.dsorted, 2);
    }

    FUZ_HD int codes() {
        for (;;) {
            refill();
            uint32_t e = T.lit[peek() & ((1u << FUZ_INF_LBITS) - 1u)];
            if ((int32_t)e >= 0) {                               // literal straight from the table
                bp += (int)(e & 15u);
                if (!io.put((uint8_t)(e >> 16))) return FUZ_INF_OVERRUN;
                continue;
            }
            if ((e & 15u) == 0) {
                e = search(T.lcount, T.lsorted, 1);
                if ((e & 15u) == 0) return FUZ_INF_BADCODE;
                if ((int32_t)e >= 0) {
                    bp += (int)(e & 15u);
                    if (!io.put((uint8_t)(e >> 16))) return FUZ_INF_OVERRUN;
                    continue;
                }
            }
            // code + extra bits come out of one 32-bit peek (<= 15 + 5 and <= 15 + 13 bits)
            if (((e >> 4) & 3u) == 2) { bp += (int)(e & 15u); return FUZ_INF_OK; }
            uint32_t w = peek();
            int nc = (int)(e & 15u), ne = (int)((e >> 8) & 15u);
            const int len = (int)((e >> 16) & 0x7FFFu) + (int)((w >> nc) & ((1u << ne) - 1u));
            bp += nc + ne;
            refill();
            w = peek();
            uint32_t d = T.dist[w & ((1u << FUZ_INF_DBITS) - 1u)];
            if ((d & 15u) == 0) {
                d = search(T.dcount, T.dsorted, 2);
                if ((d & 15u) == 0) return FUZ_INF_BADCODE;
            }
            nc = (int)(d & 15u); ne = (int)((d >> 8) & 15u);
            const int dist = (int)(d >> 16) + (int)((w >> nc) & ((1u << ne) - 1u));
            bp += nc + ne;
            // (no input check here: every match produces output, which is bounded; the input
            // position is checked at the end of every deflate block)
            if (!io.copy(len, dist)) return FUZ_INF_BADDIST;
        }
    }

    // one raw deflate stream of n_bytes starting at byte_pos of the input
    FUZ_HD int run(int64_t first_byte, int64_t n_bytes) {
        start(first_byte, n_bytes);
        for (;;) {
            const uint32_t last = take(1), type = take(2);
            int e;
            if (type == 0) {
                bp = (bp + 7) & ~7;
                const uint32_t len = take(16), nlen = take(16);
                refill();
                if ((len ^ 0xFFFFu) != nlen) return FUZ_INF_BADBLOCK;
                const int64_t src = byte_pos();
                if (src + (int64_t)len > end_byte) return FUZ_INF_INPUT;
                if (!io.copy_in(src, (int)len)) return FUZ_INF_OVERRUN;
                start(src + len, end_byte - (src + len));
                e = FUZ_INF_OK;
            } else if (type == 1) {
                e = fixed_tables();
                if (!e) e = codes();
            } else if (type == 2) {
                e = dynamic_tables();
                if (!e) e = codes();
            } else {
                return FUZ_INF_BADBLOCK;
            }
            if (e) return e;
            if (!input_ok()) return FUZ_INF_INPUT;
            if (last) return FUZ_INF_OK;
        }
    }
};
