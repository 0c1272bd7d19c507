// Synthetic workload generator of bench.py (libfuz_synth.so; test infrastructure, not part of the libfuz C ABI).
// One diploid contig with planted het SNPs and its truth-aligned reads as uncompressed BAM records, the same model as
// falcon_unzip_b200/synth.py (haplotype 0 iid uniform ACGT, het sites at het_rate, reads from a
// Bernoulli(1/2) haplotype at uniform starts with N(mu, 0.2 mu) lengths, iid errors split evenly between substitution /
// 1-base insertion / 1-base deletion, CIGAR with = / X / I / D as `blasr --bam` writes it, reference unzip.py:86-88), but
// with its own random stream (xoshiro256**, seeded by (seed, contig)), so that 15 G aligned bases take seconds instead of
// minutes.  The content depends on (seed, contig id) only.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {

struct Rng {
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    Rng(uint64_t a, uint64_t b) {
        uint64_t x = a * 0xD1342543DE82EF95ull + b + 1;
        for (int i = 0; i < 4; i++) s[i] = splitmix(x);
    }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
    double normal() {                                     // Box-Muller, one value per call
        double u1 = uniform(), u2 = uniform();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

struct Read { int64_t start, len; };

}  // namespace

// Upper bound of the record bytes / record count of one contig (sizes the caller's buffers).
extern "C" int64_t fuz_synth_bounds(int64_t contig_len, double coverage, int64_t mean_len, int64_t min_len, int64_t *n_reads) {
    const int64_t n = (int64_t)ceil(coverage * (double)contig_len / (double)mean_len);
    if (n_reads) *n_reads = n;
    // lengths are N(mu, 0.2 mu) clipped to [min_len, L]: mean below mu + 0.1 mu with room; 1.5 bytes per query base +
    // 4 bytes per CIGAR operation (3 per error at most) + 80 bytes of core and name
    const double per = 1.2 * (double)std::max(mean_len, min_len);
    return (int64_t)((double)n * (per * 1.75 + 160.0)) + (1 << 20);
}

// Generates contig `ci` (refID `refid` in the records).  ref_seq receives contig_len ASCII bases (haplotype 0 = the
// reference).  Returns the bytes written (records), -1 when a capacity is too small.
extern "C" int64_t fuz_synth_contig(uint64_t seed, int64_t ci, int32_t refid, int64_t contig_len, double coverage, int64_t mean_len,
                                    int64_t min_len, double het_rate, double error_rate, uint8_t *records, int64_t cap_bytes,
                                    int64_t *rec_off, int64_t cap_rec, int64_t *n_rec_out, char *ref_seq, int64_t *aligned_out) {
    Rng rng(seed, (uint64_t)ci);
    const int64_t L = contig_len;
    std::vector<uint8_t> h0((size_t)L), h1;
    for (int64_t i = 0; i < L; i += 32) {
        uint64_t r = rng.next();
        for (int64_t k = i; k < std::min(L, i + 32); k++, r >>= 2) h0[(size_t)k] = (uint8_t)(r & 3);
    }
    h1 = h0;
    const int64_t n_het = (int64_t)llround((double)L * het_rate);
    {   // het positions without replacement: rejection on a bitmap (n_het << L)
        std::vector<uint8_t> used((size_t)L, 0);
        for (int64_t k = 0; k < n_het;) {
            const int64_t p = (int64_t)rng.below((uint64_t)L);
            if (used[(size_t)p]) continue;
            used[(size_t)p] = 1;
            h1[(size_t)p] = (uint8_t)((h0[(size_t)p] + 1 + rng.below(3)) & 3);
            k++;
        }
    }
    if (ref_seq) for (int64_t i = 0; i < L; i++) ref_seq[i] = "ACGT"[h0[(size_t)i]];
    const int64_t n_reads = (int64_t)ceil(coverage * (double)L / (double)mean_len);
    if (n_reads > cap_rec) return -1;
    std::vector<Read> reads((size_t)n_reads);
    for (auto &r : reads) {
        double len = (double)mean_len + 0.2 * (double)mean_len * rng.normal();
        int64_t l = (int64_t)len;
        l = std::max(l, min_len); l = std::min(l, L);
        r.len = l;
        r.start = (int64_t)(rng.uniform() * (double)(L - l + 1));
    }
    std::stable_sort(reads.begin(), reads.end(), [](const Read &a, const Read &b) { return a.start < b.start; });
    static const uint8_t kNib[4] = {1, 2, 4, 8};
    const double p_event = error_rate;                   // per template base: sub / ins / del with p/3 each
    const double inv_log = p_event > 0 ? 1.0 / log(1.0 - p_event) : 0.0;
    int64_t w = 0, aligned = 0;
    std::vector<uint32_t> cig;
    std::vector<uint8_t> q;                               // query base codes 0..3
    int64_t serial = ci * 1000000;
    for (int64_t ri = 0; ri < n_reads; ri++, serial++) {
        const Read &rd = reads[(size_t)ri];
        const uint8_t *hap = (rng.next() & 1) ? h1.data() : h0.data();
        const uint16_t flag = (rng.next() & 1) ? 16 : 0;
        cig.clear(); q.clear();
        q.reserve((size_t)rd.len + 64);
        auto push_op = [&](uint32_t op, uint32_t len) {
            if (!len) return;
            if (!cig.empty() && (cig.back() & 15u) == op) cig.back() += len << 4; else cig.push_back(len << 4 | op);
        };
        // events at geometric gaps; the first and the last template base carry no event
        int64_t t = 0;                                     // template offset inside the read
        while (t < rd.len) {
            int64_t gap = rd.len - t;                      // bases copied unchanged before the next event
            if (p_event > 0) {
                double u = rng.uniform();
                if (u < 1e-300) u = 1e-300;
                const int64_t g = (int64_t)(log(u) * inv_log);
                if (g < gap) gap = g;
            }
            if (t == 0 && gap == 0) gap = 1;               // no event on the first base
            for (int64_t k = 0; k < gap; k++) q.push_back(hap[(size_t)(rd.start + t + k)]);
            push_op(7, (uint32_t)gap);                     // =
            aligned += gap;
            t += gap;
            if (t >= rd.len) break;
            if (t == rd.len - 1) {                         // no event on the last base
                q.push_back(hap[(size_t)(rd.start + t)]);
                push_op(7, 1); aligned += 1; t += 1;
                break;
            }
            const uint8_t tb = hap[(size_t)(rd.start + t)];
            switch (rng.below(3)) {
            case 0: q.push_back((uint8_t)((tb + 1 + rng.below(3)) & 3)); push_op(8, 1); aligned += 1; t += 1; break;   // X
            case 1: q.push_back((uint8_t)rng.below(4)); push_op(1, 1);                                                    // I, then the base itself
                    q.push_back(tb); push_op(7, 1); aligned += 1; t += 1; break;
            default: push_op(2, 1); t += 1; break;                                                                         // D
            }
        }
        if (cig.size() > 65535) return -2;
        char name[64];
        const int l_name = snprintf(name, sizeof(name), "m%08lld/%lld/0_%lld", (long long)serial, (long long)serial, (long long)rd.len) + 1;
        const int64_t l_seq = (int64_t)q.size(), seq_bytes = (l_seq + 1) / 2;
        const int64_t body = 32 + l_name + 4 * (int64_t)cig.size() + seq_bytes + l_seq;
        if (w + body + 4 > cap_bytes) return -1;
        rec_off[ri] = w;
        uint8_t *p = records + w;
        auto put32 = [&](int64_t o, uint32_t v) { memcpy(p + o, &v, 4); };
        put32(0, (uint32_t)body); put32(4, (uint32_t)refid); put32(8, (uint32_t)rd.start);
        p[12] = (uint8_t)l_name; p[13] = 254;
        const uint16_t bin = (uint16_t)reg2bin(rd.start, rd.start + rd.len), ncig = (uint16_t)cig.size();
        memcpy(p + 14, &bin, 2); memcpy(p + 16, &ncig, 2); memcpy(p + 18, &flag, 2);
        put32(20, (uint32_t)l_seq); put32(24, 0xFFFFFFFFu); put32(28, 0xFFFFFFFFu); put32(32, 0);
        memcpy(p + 36, name, (size_t)l_name);
        uint8_t *c = p + 36 + l_name;
        memcpy(c, cig.data(), 4 * cig.size());
        uint8_t *sq = c + 4 * cig.size();
        for (int64_t k = 0; k + 1 < l_seq; k += 2) sq[k >> 1] = (uint8_t)(kNib[q[(size_t)k]] << 4 | kNib[q[(size_t)k + 1]]);
        if (l_seq & 1) sq[l_seq >> 1] = (uint8_t)(kNib[q[(size_t)l_seq - 1]] << 4);
        memset(sq + seq_bytes, 0xFF, (size_t)l_seq);
        w += body + 4;
    }
    rec_off[n_reads] = w;
    if (n_rec_out) *n_rec_out = n_reads;
    if (aligned_out) *aligned_out = aligned;
    return w;
}
