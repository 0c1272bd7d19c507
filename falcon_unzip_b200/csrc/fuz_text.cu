// LA4Falcon text -> columns on the device (reference rr_hctg_track.py:38-44, ovlp_filter_with_phase.py:60-62,
// 95-99: `l = l.strip().split()` and the int() / float() conversions).  The host parsers of fuz_host.cpp bound
// both overlap paths end to end (10-17 M lines/s on 16 threads against kernels that take a millisecond); here the
// text crosses PCIe once and is parsed where the filter runs.
//
//   k_txt_count   one thread per 128-byte chunk: non-blank lines that END in the chunk (a look-back to the
//                 previous newline tells whether the line entering the chunk already holds a token)
//   k_scan_wide   chunk counts -> first line index of every chunk
//   k_txt_parse   the same walk; the thread whose chunk holds the END of a line parses it from its start:
//                 tokens split on blanks, integers with sign, the identity column as an exact decimal test
//                 against 90, the last token as the overlap tag; 12 column arrays + line offset / length.
// Lines the kernel cannot settle exactly (identity in exponent / inf / nan notation or with more than 15
// significant digits) are flagged (bit 7) for the host, which re-evaluates just that column.
#include "fuz_internal.cuh"

namespace {

#define TXT_CHUNK 128

struct TxtCols {
    int32_t *q, *t, *len, *qs, *qe, *ql, *ts, *te, *tl;
    uint8_t *flags;
    int64_t *off;
    int32_t *llen;
};

__device__ __forceinline__ bool txt_blank(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\f' || c == '\v'; }

// position after the newline that precedes `a` (0 at the start of the text), and whether [that, a) holds a token
__device__ __forceinline__ int64_t txt_line_start(const uint8_t *__restrict__ text, int64_t a, bool *any) {
    bool seen = false;
    int64_t p = a;
    while (p > 0) {
        const uint8_t c = __ldg(text + p - 1);
        if (c == '\n') break;
        if (!txt_blank(c)) seen = true;
        p--;
    }
    *any = seen;
    return p;
}

__global__ void __launch_bounds__(256) k_txt_count(const uint8_t *__restrict__ text, int64_t n, int64_t n_chunks, int32_t *__restrict__ cnt) {
    fuz_pdl_enter();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    const int64_t a = c * TXT_CHUNK, b = min(a + TXT_CHUNK, n);
    bool any;
    (void)txt_line_start(text, a, &any);
    int lines = 0;
    for (int64_t p = a; p < b; p++) {
        const uint8_t ch = __ldg(text + p);
        if (ch == '\n') { lines += any ? 1 : 0; any = false; }
        else if (!txt_blank(ch)) any = true;
    }
    if (b == n && any) lines++;                    // last line without a newline
    cnt[c] = lines;
}

// reasons reported in status.reserved[3]
enum { TXT_OK = 0, TXT_FEW_COLUMNS = 1, TXT_BAD_INT = 2, TXT_BAD_ID = 3, TXT_BAD_FLOAT = 4 };

// one line [s, e): returns TXT_*; *host_idt = the identity column needs the host's float()
__device__ __forceinline__ int txt_parse_line(const uint8_t *__restrict__ text, int64_t s, int64_t e, int require_id9, long long *col,
                                              int *flags, bool *host_idt) {
    int nc = 0, tag = 0;
    bool idt_ok = false;
    *host_idt = false;
    int64_t p = s;
    while (p < e) {
        while (p < e && txt_blank(__ldg(text + p))) p++;
        if (p >= e) break;
        const int64_t t0 = p;
        if (nc == 3) {
            // float(l[3]) < 90 for [+-]digits[.digits]: decided by sign and integer part, exact as long as the
            // value cannot round up to 90.0 (at most 15 significant digits); anything else goes to the host
            bool neg = false, plain = true, digits = false;
            uint8_t ch = __ldg(text + p);
            if (ch == '-' || ch == '+') { neg = ch == '-'; p++; }
            long long ip = 0;
            int nsig = 0;
            while (p < e && !txt_blank(ch = __ldg(text + p)) && ch >= '0' && ch <= '9') {
                if (ip < 1000000) ip = ip * 10 + (ch - '0');
                if (nsig || ch != '0') nsig++;
                digits = true;
                p++;
            }
            if (p < e && __ldg(text + p) == '.') {
                p++;
                while (p < e && !txt_blank(ch = __ldg(text + p)) && ch >= '0' && ch <= '9') { if (nsig || ch != '0') nsig++; digits = true; p++; }
            }
            if (p < e && !txt_blank(__ldg(text + p))) plain = false;       // exponent, inf, nan, garbage
            while (p < e && !txt_blank(__ldg(text + p))) p++;
            if (!plain || !digits || nsig > 15) *host_idt = true;
            else idt_ok = !neg && ip >= 90;                                // -x < 90 always (and -0.0 < 90)
        } else if (nc < 12 && nc != 4 && nc != 8) {
            bool neg = false;
            uint8_t ch = __ldg(text + p);
            const bool sign = ch == '-' || ch == '+';
            if (sign) { neg = ch == '-'; p++; }
            if (p >= e || txt_blank(__ldg(text + p))) return TXT_BAD_INT;
            long long v = 0;
            while (p < e && !txt_blank(ch = __ldg(text + p))) {
                if (ch < '0' || ch > '9') return TXT_BAD_INT;
                v = v * 10 + (ch - '0');
                if (v > 0x7fffffffLL) return TXT_BAD_INT;
                p++;
            }
            if (require_id9 && (nc == 0 || nc == 1) && (p - t0 != 9 || sign)) return TXT_BAD_ID;
            col[nc] = neg ? -v : v;
        } else {
            while (p < e && !txt_blank(__ldg(text + p))) p++;
        }
        // the last token decides the tag (:106,:176-179)
        const int tl = (int)(p - t0);
        tag = 0;
        if (tl == 7 || tl == 8 || tl == 9) {
            const char *w = tl == 7 ? "overlap" : tl == 8 ? "contains" : "contained";
            bool eq = true;
            for (int k = 0; k < tl; k++) eq = eq && __ldg(text + t0 + k) == (uint8_t)w[k];
            if (eq) tag = tl - 6;
        }
        nc++;
    }
    if (nc < 12) return TXT_FEW_COLUMNS;
    *flags = (idt_ok ? 1 : 0) | (tag << 1);
    return TXT_OK;
}

__global__ void __launch_bounds__(256) k_txt_parse(const uint8_t *__restrict__ text, int64_t n, int64_t n_chunks, const int32_t *__restrict__ base,
                                                   int64_t cap, int require_id9, TxtCols C, fuz_status *st) {
    fuz_pdl_enter();
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0) {
        const int64_t total = base[n_chunks];
        st->reserved[0] = total;
        if (total > cap) fuz_raise(st, FUZ_E_CAPACITY, 12);
    }
    if (c >= n_chunks || base[n_chunks] > cap) return;
    const int64_t a = c * TXT_CHUNK, b = min(a + TXT_CHUNK, n);
    bool any;
    int64_t start = txt_line_start(text, a, &any);
    int64_t w = base[c];
    for (int64_t p = a; p <= b; p++) {
        bool end_here;
        if (p < b) {
            const uint8_t ch = __ldg(text + p);
            end_here = ch == '\n';
            if (!end_here && !txt_blank(ch)) any = true;
        } else {
            end_here = b == n && any;               // last line without a newline
            if (!end_here) break;
        }
        if (!end_here) continue;
        if (any) {
            long long col[12];
            int flags = 0;
            bool host_idt;
            const int rc = txt_parse_line(text, start, p, require_id9, col, &flags, &host_idt);
            if (rc != TXT_OK) {
                fuz_raise(st, FUZ_E_FORMAT, (int)min(w, (int64_t)0x7fffffff));
                atomicMax((int *)&st->reserved[3], rc);
            } else {
                C.q[w] = (int32_t)col[0]; C.t[w] = (int32_t)col[1]; C.len[w] = (int32_t)(-col[2]);
                C.qs[w] = (int32_t)col[5]; C.qe[w] = (int32_t)col[6]; C.ql[w] = (int32_t)col[7];
                C.ts[w] = (int32_t)col[9]; C.te[w] = (int32_t)col[10]; C.tl[w] = (int32_t)col[11];
                C.flags[w] = (uint8_t)(flags | (host_idt ? 0x80 : 0));
                if (host_idt) atomicAdd((unsigned long long *)&st->reserved[1], 1ull);
            }
            C.off[w] = start; C.llen[w] = (int32_t)(p - start);
            w++;
        }
        any = false;
        start = p + 1;
    }
}

}  // namespace

extern "C" int fuz_parse_la4falcon(fuz_ctx *ctx, const uint8_t *d_text, int64_t n_bytes, int64_t cap, int32_t require_id9,
                                   int32_t *d_q, int32_t *d_t, int32_t *d_len, int32_t *d_qs, int32_t *d_qe, int32_t *d_ql,
                                   int32_t *d_ts, int32_t *d_te, int32_t *d_tl, uint8_t *d_flags, int64_t *d_line_off,
                                   int32_t *d_line_len) {
    if (!ctx || (!d_text && n_bytes) || n_bytes < 0 || cap < 0 || !d_q || !d_t || !d_len || !d_qs || !d_qe || !d_ql || !d_ts || !d_te ||
        !d_tl || !d_flags || !d_line_off || !d_line_len)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_parse_la4falcon: bad argument");
    cudaStream_t st = ctx->stream;
    const int64_t n_chunks = (n_bytes + TXT_CHUNK - 1) / TXT_CHUNK;
    if (n_chunks > 0x7ffffff0LL) return fuz_fail(ctx, FUZ_E_ARG, "fuz_parse_la4falcon: text too large for one call");
    FuzLayout L;
    const size_t o_cnt = L.add(4 * (size_t)(n_chunks + 1)), o_base = L.add(4 * (size_t)(n_chunks + 2));
    int rc = fuz_arena_commit(ctx, L);
    if (rc) return rc;
    int32_t *cnt = fuz_at<int32_t>(ctx, o_cnt), *base = fuz_at<int32_t>(ctx, o_base);
    FUZ_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(fuz_status), st));
    if (n_chunks == 0) return FUZ_OK;
    const unsigned grid = (unsigned)((n_chunks + 255) / 256);
    fuz_launch(ctx, k_txt_count, grid, 256, 0, st, d_text, n_bytes, n_chunks, cnt);
    FUZ_LAUNCH_CHECK(ctx, "k_txt_count");
    if ((rc = fuz_scan_i32_wide(ctx, cnt, base, n_chunks))) return rc;
    TxtCols C{d_q, d_t, d_len, d_qs, d_qe, d_ql, d_ts, d_te, d_tl, d_flags, d_line_off, d_line_len};
    fuz_launch(ctx, k_txt_parse, grid, 256, 0, st, d_text, n_bytes, n_chunks, base, cap, (int)require_id9, C, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_txt_parse");
    return FUZ_OK;
}
