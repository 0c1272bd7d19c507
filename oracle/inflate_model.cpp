// TEST INFRASTRUCTURE ONLY (oracle/).  Scalar instantiation of the DEFLATE decoder the CUDA
// kernel runs (falcon_unzip_b200/csrc/fuz_inflate_core.h) so that the table construction, the
// canonical search and the block parsing can be checked against zlib on the CPU
// (tests/test_inflate_model.py).  Never loaded by falcon_unzip_b200/.
#include <stdint.h>
#include <string.h>

#include "../falcon_unzip_b200/csrc/fuz_inflate_core.h"

struct HostIO {
    const uint8_t *in; int64_t n_in; int64_t widx = 0;
    uint8_t *out; int64_t outpos = 0, limit;
    int seek(int64_t b) { widx = b >> 2; return (int)(b & 3); }
    uint32_t next_word() {
        uint32_t w = 0;
        const int64_t o = widx * 4;
        widx++;
        if (o + 4 <= n_in) memcpy(&w, in + o, 4);
        else if (o < n_in) memcpy(&w, in + o, (size_t)(n_in - o));
        return w;
    }
    int64_t word_pos() const { return widx; }
    bool put(uint8_t b) { if (outpos >= limit) return false; out[outpos++] = b; return true; }
    bool copy(int len, int dist) {
        if (dist > outpos || outpos + len > limit) return false;
        for (int i = 0; i < len; i++, outpos++) out[outpos] = out[outpos - dist];
        return true;
    }
    bool copy_in(int64_t src, int len) {
        if (outpos + len > limit || src + len > n_in) return false;
        memcpy(out + outpos, in + src, (size_t)len);
        outpos += len;
        return true;
    }
    int lane() const { return 0; }
    int lanes() const { return 1; }
    void sync() {}
};

// inflates in[off, off+n) into out (capacity cap); returns the FUZ_INF_* code, *n_out = bytes produced
extern "C" int inflate_model(const uint8_t *in, int64_t n_total, int64_t off, int64_t n, uint8_t *out, int64_t cap,
                             int64_t *n_out) {
    static FuzInfTables T;
    HostIO io{in, n_total, 0, out, 0, cap};
    FuzInflate<HostIO> inf(io, T);
    const int rc = inf.run(off, n);
    *n_out = io.outpos;
    return rc;
}
