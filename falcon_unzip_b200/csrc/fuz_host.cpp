// Host-side helpers of the phasing path (no CUDA): record index, QNAME -> q_id, CPython-2
// dict order.  They feed the device kernels and format their results; none of them
// computes any pileup / phasing result.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "fuz.h"

extern "C" int fuz_host_index_records(const uint8_t *h_rec_buf, int64_t rec_bytes, int64_t *h_rec_off, int64_t cap_rec,
                                      int64_t *n_rec) {
    if (!h_rec_buf || !n_rec || rec_bytes < 0) return FUZ_E_ARG;
    int64_t o = 0, n = 0;
    while (o < rec_bytes) {
        if (o + 4 > rec_bytes) return FUZ_E_BADRECORD;
        int32_t bs;
        memcpy(&bs, h_rec_buf + o, 4);
        if (bs < 32 || o + 4 + (int64_t)bs > rec_bytes) return FUZ_E_BADRECORD;
        if (h_rec_off) {
            if (n >= cap_rec) return FUZ_E_CAPACITY;
            h_rec_off[n] = o;
        }
        o += 4 + (int64_t)bs;
        n++;
    }
    if (h_rec_off) {
        if (n > cap_rec) return FUZ_E_CAPACITY;
        h_rec_off[n] = o;
    }
    *n_rec = n;
    return FUZ_OK;
}

// first-seen QNAME -> q_id, assigned before any filtering (reference phasing.py:47-54)
extern "C" int fuz_host_assign_qids(const uint8_t *h_rec_buf, const int64_t *h_rec_off, int64_t n_rec,
                                    const int32_t *h_ctg_rec_off, int32_t n_ctg, int32_t *h_rec_qid, int32_t *h_ctg_nq,
                                    int64_t *h_name_first) {
    if (!h_rec_buf || !h_rec_off || !h_ctg_rec_off || !h_rec_qid || !h_ctg_nq || n_ctg < 1) return FUZ_E_ARG;
    int64_t q_base = 0;
    for (int32_t c = 0; c < n_ctg; c++) {
        const int64_t r0 = h_ctg_rec_off[c], r1 = h_ctg_rec_off[c + 1];
        if (r0 > r1 || r1 > n_rec) return FUZ_E_ARG;
        std::unordered_map<std::string_view, int32_t> table;
        table.reserve((size_t)(r1 - r0) * 2 + 16);
        for (int64_t r = r0; r < r1; r++) {
            const uint8_t *rec = h_rec_buf + h_rec_off[r];
            const int l_name = rec[12];
            if (l_name < 1) return FUZ_E_BADRECORD;
            std::string_view name(reinterpret_cast<const char *>(rec + 36), (size_t)l_name - 1);
            auto it = table.find(name);
            int32_t q;
            if (it == table.end()) {
                q = (int32_t)table.size();
                table.emplace(name, q);
                if (h_name_first) h_name_first[q_base + q] = r;
            } else {
                q = it->second;
            }
            h_rec_qid[r] = q;
        }
        h_ctg_nq[c] = (int32_t)table.size();
        q_base += (int64_t)table.size();
    }
    return FUZ_OK;
}

// CPython 2.7 Objects/dictobject.c, insert-only, int keys (hash(i) = i, -1 -> -2):
// iteration order = slot order.
extern "C" int fuz_host_py27_int_dict_order(const int64_t *keys, int64_t n, int64_t *out) {
    if ((!keys || !out) && n > 0) return FUZ_E_ARG;
    struct Entry { int64_t hash; int64_t key; bool used; };
    std::vector<Entry> slots(8, Entry{0, 0, false});
    uint64_t mask = 7;
    int64_t used = 0;
    auto find_slot = [](std::vector<Entry> &tab, uint64_t m, int64_t hash, int64_t key, bool match) -> size_t {
        uint64_t i = (uint64_t)hash & m;
        uint64_t perturb = (uint64_t)hash;
        for (;;) {
            Entry &e = tab[i & m];
            if (!e.used || (match && e.hash == hash && e.key == key)) return (size_t)(i & m);
            i = i * 5 + perturb + 1;
            perturb >>= 5;
        }
    };
    for (int64_t k = 0; k < n; k++) {
        const int64_t key = keys[k];
        const int64_t hash = key == -1 ? -2 : key;
        size_t s = find_slot(slots, mask, hash, key, true);
        if (slots[s].used) continue;
        slots[s] = Entry{hash, key, true};
        used++;
        if ((uint64_t)used * 3 >= (mask + 1) * 2) {
            const uint64_t minused = (uint64_t)(used > 50000 ? 2 : 4) * (uint64_t)used;
            uint64_t newsize = 8;
            while (newsize <= minused) newsize <<= 1;
            std::vector<Entry> fresh(newsize, Entry{0, 0, false});
            for (const Entry &e : slots)
                if (e.used) fresh[find_slot(fresh, newsize - 1, e.hash, e.key, false)] = e;
            slots.swap(fresh);
            mask = newsize - 1;
        }
    }
    int64_t w = 0;
    for (const Entry &e : slots)
        if (e.used) out[w++] = e.key;
    return FUZ_OK;
}

// LA4Falcon -m lines: "q t -len idt qstrand qs qe ql tstrand ts te tl tag" (rr_hctg_track.py:38-44)
static int64_t parse_la4falcon_range(const char *text, int64_t n_bytes, int64_t cap, int32_t *q, int32_t *t, int32_t *len,
                                     int32_t *tlen) {
    int64_t n = 0, i = 0;
    while (i < n_bytes && n < cap) {
        // one line
        long long col[12];
        int c = 0;
        bool any = false;
        while (i < n_bytes && text[i] != '\n') {
            while (i < n_bytes && (text[i] == ' ' || text[i] == '\t' || text[i] == '\r')) i++;
            if (i >= n_bytes || text[i] == '\n') break;
            any = true;
            // token
            int64_t s = i;
            while (i < n_bytes && text[i] != ' ' && text[i] != '\t' && text[i] != '\n' && text[i] != '\r') i++;
            if (c < 12) {
                if (c == 3) { col[c] = 0; }            // idt: float, unused (:42)
                else {
                    bool neg = false; int64_t k = s; long long v = 0;
                    if (k < i && (text[k] == '-' || text[k] == '+')) { neg = text[k] == '-'; k++; }
                    if (k == i) { if (c == 0 || c == 1 || c == 2 || c == 11) return -1; }
                    for (; k < i; k++) {
                        if (text[k] < '0' || text[k] > '9') { if (c == 0 || c == 1 || c == 2 || c == 11) return -1; v = 0; break; }
                        v = v * 10 + (text[k] - '0');
                    }
                    col[c] = neg ? -v : v;
                }
            }
            c++;
        }
        if (i < n_bytes) i++;                          // newline
        if (!any) continue;                            // blank line
        if (c < 12) return -1;
        q[n] = (int32_t)col[0]; t[n] = (int32_t)col[1]; len[n] = (int32_t)(-col[2]); tlen[n] = (int32_t)col[11];
        n++;
    }
    return n;
}

// Large inputs are cut at line ends into one piece per host thread; every piece is parsed
// into its own vectors, which are then copied to their place in the output (line order kept).
extern "C" int64_t fuz_host_parse_la4falcon(const char *text, int64_t n_bytes, int64_t cap, int32_t *q, int32_t *t,
                                            int32_t *len, int32_t *tlen) {
    if (!text || n_bytes < 0 || !q || !t || !len || !tlen) return -1;
    unsigned hw = std::thread::hardware_concurrency();
    int n_thr = (int)std::min<int64_t>(std::min<unsigned>(hw ? hw : 1, 32), n_bytes / (1 << 20));
    if (n_thr <= 1) return parse_la4falcon_range(text, n_bytes, cap, q, t, len, tlen);
    std::vector<int64_t> cut(n_thr + 1, n_bytes);
    cut[0] = 0;
    for (int k = 1; k < n_thr; k++) {
        int64_t p = std::max(cut[k - 1], n_bytes / n_thr * k);
        const void *nl = p < n_bytes ? memchr(text + p, '\n', (size_t)(n_bytes - p)) : nullptr;
        cut[k] = nl ? (const char *)nl - text + 1 : n_bytes;
    }
    struct Piece { std::vector<int32_t> q, t, len, tlen; int64_t n = 0; };
    std::vector<Piece> pieces(n_thr);
    std::vector<std::thread> pool;
    for (int k = 0; k < n_thr; k++)
        pool.emplace_back([&, k] {
            Piece &pc = pieces[k];
            const int64_t bytes = cut[k + 1] - cut[k];
            int64_t lines = 1;
            for (const char *s = text + cut[k], *e = s + bytes; s < e;) {
                const void *nl = memchr(s, '\n', (size_t)(e - s));
                if (!nl) break;
                lines++;
                s = (const char *)nl + 1;
            }
            pc.q.resize(lines); pc.t.resize(lines); pc.len.resize(lines); pc.tlen.resize(lines);
            pc.n = parse_la4falcon_range(text + cut[k], bytes, lines, pc.q.data(), pc.t.data(), pc.len.data(), pc.tlen.data());
        });
    for (auto &th : pool) th.join();
    int64_t total = 0;
    for (auto &pc : pieces) {
        if (pc.n < 0) return -1;
        total += pc.n;
    }
    std::vector<int64_t> at(n_thr + 1, 0);
    for (int k = 0; k < n_thr; k++) at[k + 1] = at[k] + pieces[k].n;
    pool.clear();
    for (int k = 0; k < n_thr; k++)
        pool.emplace_back([&, k] {
            const Piece &pc = pieces[k];
            const int64_t room = std::max<int64_t>(0, std::min(pc.n, cap - at[k]));     // stops at cap like the serial parser
            memcpy(q + at[k], pc.q.data(), (size_t)room * 4); memcpy(t + at[k], pc.t.data(), (size_t)room * 4);
            memcpy(len + at[k], pc.len.data(), (size_t)room * 4); memcpy(tlen + at[k], pc.tlen.data(), (size_t)room * 4);
        });
    for (auto &th : pool) th.join();
    return std::min(total, cap);
}

// ---------------------------------------------------------------- LA4Falcon -mo (overlap filter)
// "q t -len idt qstrand qs qe ql tstrand ts te tl tag" (reference ovlp_filter_with_phase.py:60-62,
// 95-99): every column the three filter stages read, plus where the line sits in the text (the
// selected lines are printed again, :352).  flags: bit 0 = float(col 3) >= 90 (i.e. not `idt < 90`),
// bits 1-2 = last token: 1 "overlap", 2 "contains", 3 "contained", 0 anything else.
// q / t must be %09d ids (9 digits): the reference compares and sorts them as strings.
#include <charconv>
#include <stdlib.h>

struct MoCols {
    int32_t *q, *t, *len, *qs, *qe, *ql, *ts, *te, *tl;
    uint8_t *flags;
    int64_t *off;
    int32_t *llen;
};

static int64_t parse_mo_range(const char *text, int64_t base, int64_t n_bytes, int64_t cap, const MoCols &c, int64_t at) {
    int64_t n = 0, i = 0;
    while (i < n_bytes && n < cap) {
        const int64_t line0 = i;
        long long col[12];
        int nc = 0;
        bool idt_ok = false, any = false;
        int64_t last_s = 0, last_e = 0;
        while (i < n_bytes && text[i] != '\n') {
            while (i < n_bytes && (text[i] == ' ' || text[i] == '\t' || text[i] == '\r' || text[i] == '\f' || text[i] == '\v')) i++;
            if (i >= n_bytes || text[i] == '\n') break;
            any = true;
            const int64_t s = i;
            while (i < n_bytes && text[i] != ' ' && text[i] != '\t' && text[i] != '\n' && text[i] != '\r' && text[i] != '\f' && text[i] != '\v') i++;
            last_s = s; last_e = i;
            if (nc < 12) {
                if (nc == 3) {                          // idt: float(l[3]) < 90 (:97,:100)
                    double v = 0;
                    auto r = std::from_chars(text + s + (text[s] == '+' ? 1 : 0), text + i, v);
                    if (r.ec != std::errc() || r.ptr != text + i) return -1;
                    idt_ok = !(v < 90.0);
                    col[nc] = 0;
                } else if (nc == 4 || nc == 8) {
                    col[nc] = 0;                        // strands: never read
                } else {
                    bool neg = false;
                    int64_t k = s;
                    long long v = 0;
                    if (k < i && (text[k] == '-' || text[k] == '+')) { neg = text[k] == '-'; k++; }
                    if (k == i) return -1;
                    for (; k < i; k++) {
                        if (text[k] < '0' || text[k] > '9') return -1;
                        v = v * 10 + (text[k] - '0');
                        if (v > 0x7fffffffLL) return -1;
                    }
                    if ((nc == 0 || nc == 1) && (i - s != 9 || neg || text[s] == '+')) return -2;   // not a %09d id
                    col[nc] = neg ? -v : v;
                }
            }
            nc++;
        }
        const int64_t line1 = i;
        if (i < n_bytes) i++;
        if (!any) continue;
        if (nc < 12) return -1;
        int tag = 0;
        const int64_t tl_ = last_e - last_s;
        if (tl_ == 7 && !memcmp(text + last_s, "overlap", 7)) tag = 1;
        else if (tl_ == 8 && !memcmp(text + last_s, "contains", 8)) tag = 2;
        else if (tl_ == 9 && !memcmp(text + last_s, "contained", 9)) tag = 3;
        const int64_t w = at + n;
        c.q[w] = (int32_t)col[0]; c.t[w] = (int32_t)col[1]; c.len[w] = (int32_t)(-col[2]);
        c.qs[w] = (int32_t)col[5]; c.qe[w] = (int32_t)col[6]; c.ql[w] = (int32_t)col[7];
        c.ts[w] = (int32_t)col[9]; c.te[w] = (int32_t)col[10]; c.tl[w] = (int32_t)col[11];
        c.flags[w] = (uint8_t)((idt_ok ? 1 : 0) | (tag << 1));
        c.off[w] = base + line0; c.llen[w] = (int32_t)(line1 - line0);
        n++;
    }
    return n;
}

// Returns the number of lines parsed, -1 on a malformed line (the reference would raise), -2 when a
// read id is not a 9-digit %09d id.  Lines are split over the host threads at line ends.
extern "C" int64_t fuz_host_parse_la4falcon_mo(const char *text, int64_t n_bytes, int64_t cap, int32_t *q, int32_t *t, int32_t *len,
                                               int32_t *qs, int32_t *qe, int32_t *ql, int32_t *ts, int32_t *te, int32_t *tl,
                                               uint8_t *flags, int64_t *line_off, int32_t *line_len) {
    if (!text || n_bytes < 0 || !q || !t || !len || !qs || !qe || !ql || !ts || !te || !tl || !flags || !line_off || !line_len) return -1;
    const MoCols c{q, t, len, qs, qe, ql, ts, te, tl, flags, line_off, line_len};
    unsigned hw = std::thread::hardware_concurrency();
    int n_thr = (int)std::min<int64_t>(std::min<unsigned>(hw ? hw : 1, 32), n_bytes / (1 << 20));
    if (n_thr <= 1) return parse_mo_range(text, 0, n_bytes, cap, c, 0);
    std::vector<int64_t> cut(n_thr + 1, n_bytes), lines(n_thr, 0), at(n_thr + 1, 0), got(n_thr, 0);
    cut[0] = 0;
    for (int k = 1; k < n_thr; k++) {
        int64_t p = std::max(cut[k - 1], n_bytes / n_thr * k);
        const void *nl = p < n_bytes ? memchr(text + p, '\n', (size_t)(n_bytes - p)) : nullptr;
        cut[k] = nl ? (const char *)nl - text + 1 : n_bytes;
    }
    // pass 1: non-blank lines per piece (so that every piece knows where its rows go)
    auto count_lines = [&](int k) {
        int64_t n = 0;
        bool any = false;
        for (int64_t i = cut[k]; i < cut[k + 1]; i++) {
            const char ch = text[i];
            if (ch == '\n') { n += any; any = false; }
            else if (ch != ' ' && ch != '\t' && ch != '\r' && ch != '\f' && ch != '\v') any = true;
        }
        lines[k] = n + (any ? 1 : 0);
    };
    std::vector<std::thread> pool;
    for (int k = 0; k < n_thr; k++) pool.emplace_back(count_lines, k);
    for (auto &th : pool) th.join();
    for (int k = 0; k < n_thr; k++) at[k + 1] = at[k] + lines[k];
    if (at[n_thr] > cap) return -1;
    pool.clear();
    for (int k = 0; k < n_thr; k++)
        pool.emplace_back([&, k] { got[k] = parse_mo_range(text + cut[k], cut[k], cut[k + 1] - cut[k], lines[k], c, at[k]); });
    for (auto &th : pool) th.join();
    for (int k = 0; k < n_thr; k++) {
        if (got[k] < 0) return got[k];
        if (got[k] != lines[k]) return -1;
    }
    return at[n_thr];
}

// Output text of the overlap filter (reference ovlp_filter_with_phase.py:266-275, 352): for every
// selected line its whitespace-split tokens joined by single blanks, then the phase strings
// ("ctg.block.phase") of its q and t read.  phase_text / phase_off: one string per read id.
// Call with out = NULL to get the size.
extern "C" int64_t fuz_host_format_ovlp(const char *text, const int64_t *line_off, const int32_t *line_len, const int32_t *q,
                                        const int32_t *t, const int64_t *sel, int64_t n_sel, const char *phase_text,
                                        const int64_t *phase_off, char *out, int64_t cap) {
    if (!text || !line_off || !line_len || !q || !t || (!sel && n_sel) || !phase_text || !phase_off) return -1;
    unsigned hw = std::thread::hardware_concurrency();
    const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<unsigned>(hw ? hw : 1, 32), n_sel / 4096));
    std::vector<int64_t> sizes(n_thr, 0), start(n_thr + 1, 0);
    auto emit = [&](int k, char *dst) -> int64_t {
        int64_t w = 0;
        const int64_t lo = n_sel * k / n_thr, hi = n_sel * (k + 1) / n_thr;
        for (int64_t s = lo; s < hi; s++) {
            const int64_t li = sel[s];
            const char *p = text + line_off[li], *e = p + line_len[li];
            bool first = true;
            while (p < e) {
                while (p < e && (*p == ' ' || *p == '\t' || *p == '\r' || *p == '\f' || *p == '\v')) p++;
                if (p >= e) break;
                const char *s0 = p;
                while (p < e && *p != ' ' && *p != '\t' && *p != '\r' && *p != '\f' && *p != '\v') p++;
                if (!first) { if (dst) dst[w] = ' '; w++; }
                if (dst) memcpy(dst + w, s0, (size_t)(p - s0));
                w += p - s0;
                first = false;
            }
            for (int side = 0; side < 2; side++) {
                const int32_t r = side ? t[li] : q[li];
                const int64_t a = phase_off[r], b = phase_off[r + 1];
                if (dst) dst[w] = ' ';
                w++;
                if (dst) memcpy(dst + w, phase_text + a, (size_t)(b - a));
                w += b - a;
            }
            if (dst) dst[w] = '\n';
            w++;
        }
        return w;
    };
    std::vector<std::thread> pool;
    for (int k = 0; k < n_thr; k++) pool.emplace_back([&, k] { sizes[k] = emit(k, nullptr); });
    for (auto &th : pool) th.join();
    for (int k = 0; k < n_thr; k++) start[k + 1] = start[k] + sizes[k];
    if (!out) return start[n_thr];
    if (start[n_thr] > cap) return -1;
    pool.clear();
    for (int k = 0; k < n_thr; k++) pool.emplace_back([&, k] { emit(k, out + start[k]); });
    for (auto &th : pool) th.join();
    return start[n_thr];
}

// ---------------------------------------------------------------- CPython-2.7 orders for the tracking rows
// rawread_to_contigs is printed in the iteration order of CPython-2 dicts (rr_hctg_track.py:97-100,113,126;
// SURVEY.md B.4).  The emulation of Objects/stringobject.c:string_hash and of the insert-only table of
// dictobject.c below is the C++ twin of falcon_unzip_b200/py2compat.py (which tests compare it with): the
// Python version costs 0.6 s per 1.5 M overlap lines, more than everything on the device together.
#include <stdio.h>

static int64_t py27_str_hash(const char *s, size_t n) {
    if (n == 0) return 0;
    uint64_t x = (uint64_t)(unsigned char)s[0] << 7;
    for (size_t i = 0; i < n; i++) x = (1000003ULL * x) ^ (unsigned char)s[i];
    x ^= (uint64_t)n;
    int64_t h = (int64_t)x;
    return h == -1 ? -2 : h;
}

struct Py27Table {
    struct Entry { int64_t hash; int64_t key; };
    std::vector<Entry> slots;
    std::vector<uint8_t> used;
    uint64_t mask = 7;
    int64_t n_used = 0;
    Py27Table() : slots(8), used(8, 0) {}
    template <class Eq>
    bool insert(int64_t hash, int64_t key, Eq eq) {                 // false: an equal key is there already
        uint64_t i = (uint64_t)hash & mask, perturb = (uint64_t)hash;
        for (;;) {
            const uint64_t s = i & mask;
            if (!used[s]) { slots[s] = Entry{hash, key}; used[s] = 1; break; }
            if (slots[s].hash == hash && eq(slots[s].key, key)) return false;
            i = i * 5 + perturb + 1;
            perturb >>= 5;
        }
        n_used++;
        if ((uint64_t)n_used * 3 >= (mask + 1) * 2) {
            const uint64_t minused = (uint64_t)(n_used > 50000 ? 2 : 4) * (uint64_t)n_used;
            uint64_t newsize = 8;
            while (newsize <= minused) newsize <<= 1;
            std::vector<Entry> ns(newsize);
            std::vector<uint8_t> nu(newsize, 0);
            const uint64_t nm = newsize - 1;
            for (uint64_t k = 0; k <= mask; k++) {
                if (!used[k]) continue;
                uint64_t j = (uint64_t)slots[k].hash & nm, p = (uint64_t)slots[k].hash;
                while (nu[j & nm]) { j = j * 5 + p + 1; p >>= 5; }
                ns[j & nm] = slots[k]; nu[j & nm] = 1;
            }
            slots.swap(ns); used.swap(nu); mask = nm;
        }
        return true;
    }
    template <class F>
    void each(F f) const {
        for (uint64_t k = 0; k <= mask; k++)
            if (used[k]) f(slots[k].key);
    }
};

// keys[i] = blob[off[i], off[i+1]) inserted in order (duplicates ignored); out = index of every distinct key in
// dict iteration order.  Returns the number of distinct keys.
extern "C" int64_t fuz_host_py27_str_dict_order(const char *blob, const int64_t *off, int64_t n, int64_t *out) {
    if ((!blob || !off || !out) && n > 0) return -1;
    Py27Table tab;
    auto eq = [&](int64_t a, int64_t b) {
        const int64_t la = off[a + 1] - off[a], lb = off[b + 1] - off[b];
        return la == lb && !memcmp(blob + off[a], blob + off[b], (size_t)la);
    };
    for (int64_t i = 0; i < n; i++) tab.insert(py27_str_hash(blob + off[i], (size_t)(off[i + 1] - off[i])), i, eq);
    int64_t w = 0;
    tab.each([&](int64_t k) { out[w++] = k; });
    return w;
}

static int64_t id9_hash(int32_t id) {                  // hash of the string "%09d" % id
    char buf[16];
    if (id < 0 || id > 999999999) return py27_str_hash(buf, (size_t)snprintf(buf, sizeof(buf), "%09d", id));
    uint32_t v = (uint32_t)id;
    for (int k = 8; k >= 0; k--) { buf[k] = (char)('0' + v % 10); v /= 10; }
    return py27_str_hash(buf, 9);
}

// b-reads in the iteration order of the reference's bread_to_areads dict (rr_hctg_track.py:97-100,113): per LAS
// file the targets of its kept lines are inserted at their first kept line into the file's dict, whose iteration
// order feeds the merged dict (first sight wins); the result is the iteration order of the merged dict.  Keys are
// the "%09d" strings of the ids.  t_kept / file_kept: target and file of every KEPT line in (file, line) order.
extern "C" int64_t fuz_host_rr_bread_order(const int32_t *t_kept, const int32_t *file_kept, int64_t n, int32_t *out) {
    if ((!t_kept || !file_kept || !out) && n > 0) return -1;
    auto eq = [](int64_t a, int64_t b) { return a == b; };
    Py27Table merged;
    int32_t t_max = -1;
    for (int64_t i = 0; i < n; i++) {
        if (t_kept[i] < 0) return -1;
        t_max = std::max(t_max, t_kept[i]);
    }
    std::vector<int64_t> stamp((size_t)t_max + 1, -1);     // last segment in which the target was inserted
    for (int64_t i = 0; i < n;) {
        int64_t j = i;
        Py27Table per_file;
        while (j < n && file_kept[j] == file_kept[i]) {
            const int32_t t = t_kept[j];
            if (stamp[t] != i) { stamp[t] = i; per_file.insert(id9_hash(t), t, eq); }
            j++;
        }
        per_file.each([&](int64_t k) { merged.insert(id9_hash((int32_t)k), k, eq); });
        i = j;
    }
    int64_t w = 0;
    merged.each([&](int64_t k) { out[w++] = (int32_t)k; });
    return w;
}

// Rows of rawread_to_contigs (rr_hctg_track.py:126-138) for the given b-reads, in the given order: the contigs a
// b-read voted for in dict order of their names, stably sorted by score, "bread ctg count rank score in_ctg".
// Returns the size of the text (call with out = NULL first), -1 if cap is too small.
extern "C" int64_t fuz_host_rr_format_rows(const int32_t *breads, int64_t n_breads, const int32_t *vt_off, const int32_t *vt_ctg,
                                           const int32_t *vt_count, const int64_t *vt_score, const char *ctg_blob,
                                           const int64_t *ctg_off, const uint8_t *in_map, const int32_t *rc_off,
                                           const int32_t *rc_ctg, char *out, int64_t cap) {
    if ((!breads && n_breads) || !vt_off || !vt_ctg || !vt_count || !vt_score || !ctg_blob || !ctg_off || !in_map || !rc_off || !rc_ctg)
        return -1;
    int64_t w = 0;
    std::vector<std::pair<int64_t, int>> items;        // (score, vote row), in dict order first
    char line[512];
    for (int64_t b = 0; b < n_breads; b++) {
        const int32_t tid = breads[b];
        const int lo = vt_off[tid], hi = vt_off[tid + 1];
        if (lo == hi) continue;
        Py27Table tab;
        auto eq = [&](int64_t x, int64_t y) { return vt_ctg[x] == vt_ctg[y]; };
        for (int r = lo; r < hi; r++) {
            const int c = vt_ctg[r];
            tab.insert(py27_str_hash(ctg_blob + ctg_off[c], (size_t)(ctg_off[c + 1] - ctg_off[c])), r, eq);
        }
        items.clear();
        tab.each([&](int64_t r) { items.emplace_back(vt_score[r], (int)r); });
        std::stable_sort(items.begin(), items.end(), [](const std::pair<int64_t, int> &x, const std::pair<int64_t, int> &y) { return x.first < y.first; });
        int rank = 0;
        for (const auto &it : items) {
            const int r = it.second, c = vt_ctg[r];
            int in_ctg = 0;
            if (in_map[tid])
                for (int k = rc_off[tid]; k < rc_off[tid + 1]; k++) in_ctg |= rc_ctg[k] == c;
            const int nn = snprintf(line, sizeof(line), "%09d %.*s %d %d %lld %d\n", tid, (int)(ctg_off[c + 1] - ctg_off[c]),
                                    ctg_blob + ctg_off[c], vt_count[r], rank, (long long)it.first, in_ctg);
            if (nn < 0 || nn >= (int)sizeof(line)) return -1;
            if (out) {
                if (w + nn > cap) return -1;
                memcpy(out + w, line, (size_t)nn);
            }
            w += nn;
            rank++;
        }
    }
    return w;
}

// ---------------------------------------------------------------- text of the two large per-contig files
// het_call/variant_map ("pos ref allele q_id", phasing.py:126,128) and g_atable/atable ("pos1 b11 b12 pos2 b21 b22
// c11 c12 c21 c22", phasing.py:199) hold ~depth x sites and ~10 x sites rows; formatting them row by row in Python
// took two thirds of the file-level call.  Both return the size of the text, -1 if cap is too small, -2 if a
// position lies outside ref_seq (the reference raises IndexError, phasing.py:123).
static inline char *put_int(char *p, long long v) {
    char tmp[24];
    int n = 0;
    unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (v < 0) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}

extern "C" int64_t fuz_host_format_variant_map(const int32_t *site_pos, const int32_t *vm_site, const uint8_t *vm_base,
                                               const int32_t *vm_qid, int64_t v0, int64_t v1, const char *ref_seq, int64_t ref_len,
                                               char *out, int64_t cap) {
    if (!site_pos || !vm_site || !vm_base || !vm_qid || !ref_seq || !out || v0 > v1) return -1;
    if ((v1 - v0) * 32 > cap) return -1;
    static const char B[] = "ACGT";
    char *p = out;
    for (int64_t i = v0; i < v1; i++) {
        const int32_t pos = site_pos[vm_site[i]];
        if (pos < 1 || pos > ref_len || vm_base[i] > 3) return -2;
        p = put_int(p, pos); *p++ = ' '; *p++ = ref_seq[pos - 1]; *p++ = ' '; *p++ = B[vm_base[i]]; *p++ = ' ';
        p = put_int(p, vm_qid[i]); *p++ = '\n';
    }
    return p - out;
}

extern "C" int64_t fuz_host_format_atable(const int32_t *site_pos, const uint8_t *site_al, const int32_t *at_s1, const int32_t *at_s2,
                                          const int32_t *at_ct, int64_t a0, int64_t a1, char *out, int64_t cap) {
    if (!site_pos || !site_al || !at_s1 || !at_s2 || !at_ct || !out || a0 > a1) return -1;
    if ((a1 - a0) * 96 > cap) return -1;
    static const char B[] = "ACGT";
    char *p = out;
    for (int64_t i = a0; i < a1; i++) {
        const int32_t s1 = at_s1[i], s2 = at_s2[i];
        if (site_al[2 * s1] > 3 || site_al[2 * s1 + 1] > 3 || site_al[2 * s2] > 3 || site_al[2 * s2 + 1] > 3) return -2;
        p = put_int(p, site_pos[s1]); *p++ = ' '; *p++ = B[site_al[2 * s1]]; *p++ = ' '; *p++ = B[site_al[2 * s1 + 1]]; *p++ = ' ';
        p = put_int(p, site_pos[s2]); *p++ = ' '; *p++ = B[site_al[2 * s2]]; *p++ = ' '; *p++ = B[site_al[2 * s2 + 1]];
        for (int k = 0; k < 4; k++) { *p++ = ' '; p = put_int(p, at_ct[4 * i + k]); }
        *p++ = '\n';
    }
    return p - out;
}

// phased_reads of one contig ("q_id ctg block phase n0 n1 qname", phasing.py:478,480): reads in the iteration
// order of the reference's read_to_variants dict -- a CPython-2 dict with int keys inserted in order of first
// appearance in variant_map.  vm_qid: the contig's variant_map rows; pr_*: the contig's
// phased_reads rows, sorted by q_id; names: blob / offsets of the QNAME of every q_id.  Returns the size of the
// text (call with out = NULL first), -1 if cap is too small or a q_id has no name.
// name_of(q, &len) -> pointer to the QNAME of q_id q
template <class NameOf>
static int64_t format_phased_reads_core(const int32_t *vm_qid, int64_t n_vm, const int32_t *pr_qid, const int32_t *pr_block,
                                        const int32_t *pr_phase, const int32_t *pr_n0, const int32_t *pr_n1, int64_t n_pr,
                                        const char *ctg_id, int64_t n_names, NameOf name_of, char *out, int64_t cap) {
    Py27Table tab;
    auto eq = [](int64_t a, int64_t b) { return a == b; };
    for (int64_t i = 0; i < n_vm; i++) {
        const int64_t k = vm_qid[i];
        tab.insert(k == -1 ? -2 : k, k, eq);                // hash(int) = the int (-1 -> -2); duplicates are ignored
    }
    const size_t ctg_len = strlen(ctg_id);
    int64_t w = 0;
    bool bad = false;
    tab.each([&](int64_t q) {
        const int32_t *lo = std::lower_bound(pr_qid, pr_qid + n_pr, (int32_t)q);
        for (const int32_t *p = lo; p < pr_qid + n_pr && *p == (int32_t)q; p++) {
            const int64_t i = p - pr_qid;
            if (q < 0 || q >= n_names) { bad = true; return; }
            int64_t nl = 0;
            const char *nm = name_of(q, &nl);
            const int64_t need = 5 * 12 + (int64_t)ctg_len + nl + 8;
            if (out) {
                if (w + need > cap) { bad = true; return; }
                char *s = out + w;
                s = put_int(s, q); *s++ = ' ';
                memcpy(s, ctg_id, ctg_len); s += ctg_len; *s++ = ' ';
                s = put_int(s, pr_block[i]); *s++ = ' '; s = put_int(s, pr_phase[i]); *s++ = ' ';
                s = put_int(s, pr_n0[i]); *s++ = ' '; s = put_int(s, pr_n1[i]); *s++ = ' ';
                memcpy(s, nm, (size_t)nl); s += nl; *s++ = '\n';
                w = s - out;
            } else {
                w += need;                                   // upper bound for the size query
            }
        }
    });
    return bad ? -1 : w;
}

extern "C" int64_t fuz_host_format_phased_reads(const int32_t *vm_qid, int64_t n_vm, const int32_t *pr_qid, const int32_t *pr_block,
                                                const int32_t *pr_phase, const int32_t *pr_n0, const int32_t *pr_n1, int64_t n_pr,
                                                const char *ctg_id, const char *name_blob, const int64_t *name_off, int64_t n_names,
                                                char *out, int64_t cap) {
    if ((!vm_qid && n_vm) || (n_pr && (!pr_qid || !pr_block || !pr_phase || !pr_n0 || !pr_n1)) || !ctg_id || !name_blob || !name_off) return -1;
    return format_phased_reads_core(vm_qid, n_vm, pr_qid, pr_block, pr_phase, pr_n0, pr_n1, n_pr, ctg_id, n_names,
                                    [&](int64_t q, int64_t *nl) { *nl = name_off[q + 1] - name_off[q]; return name_blob + name_off[q]; }, out, cap);
}

// The same with the QNAMEs as fixed-width rows (NUL padded, `width` bytes each: the layout the device gathers them in).
extern "C" int64_t fuz_host_format_phased_reads_rows(const int32_t *vm_qid, int64_t n_vm, const int32_t *pr_qid, const int32_t *pr_block,
                                                     const int32_t *pr_phase, const int32_t *pr_n0, const int32_t *pr_n1, int64_t n_pr,
                                                     const char *ctg_id, const char *name_rows, int64_t width, int64_t n_names,
                                                     char *out, int64_t cap) {
    if ((!vm_qid && n_vm) || (n_pr && (!pr_qid || !pr_block || !pr_phase || !pr_n0 || !pr_n1)) || !ctg_id || (!name_rows && n_names) || width < 1) return -1;
    return format_phased_reads_core(vm_qid, n_vm, pr_qid, pr_block, pr_phase, pr_n0, pr_n1, n_pr, ctg_id, n_names,
                                    [&](int64_t q, int64_t *nl) { const char *r = name_rows + q * width; *nl = (int64_t)strnlen(r, (size_t)width); return r; },
                                    out, cap);
}

// het_call/q_id_map (phasing.py:132-134): "q_id qname" for q_id = 0 .. n-1 from fixed-width QNAME rows.  cap >= n * (width + 13).
extern "C" int64_t fuz_host_format_q_id_map_rows(const char *name_rows, int64_t width, int64_t n, char *out, int64_t cap) {
    if ((!name_rows && n) || !out || width < 1 || n < 0 || n * (width + 13) > cap) return -1;
    char *p = out;
    for (int64_t q = 0; q < n; q++) {
        const char *r = name_rows + q * width;
        const size_t nl = strnlen(r, (size_t)width);
        p = put_int(p, q); *p++ = ' ';
        memcpy(p, r, nl); p += nl; *p++ = '\n';
    }
    return p - out;
}

// het_call/variant_pos (phasing.py:116-124): "pos ref total b0 c0 b1 c1 b2 c2 b3 c3", bases by descending (count, base).
// site_cnt: 4 counts per site in A, C, G, T order.  cap >= 80 bytes per row.  -2: position outside ref_seq.
extern "C" int64_t fuz_host_format_variant_pos(const int32_t *site_pos, const int32_t *site_cnt, int64_t s0, int64_t s1,
                                               const char *ref_seq, int64_t ref_len, char *out, int64_t cap) {
    if (!site_pos || !site_cnt || !ref_seq || !out || s0 > s1 || (s1 - s0) * 80 > cap) return -1;
    static const char B[] = "ACGT";
    char *p = out;
    for (int64_t i = s0; i < s1; i++) {
        const int32_t pos = site_pos[i];
        if (pos < 1 || pos > ref_len) return -2;
        const int32_t *c = site_cnt + 4 * i;
        int ord[4] = {0, 1, 2, 3};
        // descending (count, base): keys 4 * count + base are distinct
        std::sort(ord, ord + 4, [&](int a, int b) { return 4LL * c[a] + a > 4LL * c[b] + b; });
        p = put_int(p, pos); *p++ = ' '; *p++ = ref_seq[pos - 1]; *p++ = ' ';
        p = put_int(p, (long long)c[0] + c[1] + c[2] + c[3]);
        for (int k = 0; k < 4; k++) { *p++ = ' '; *p++ = B[ord[k]]; *p++ = ' '; p = put_int(p, c[ord[k]]); }
        *p++ = '\n';
    }
    return p - out;
}

// Python 2 str(float): '%.12g', plus '.0' when the text has no '.', 'e', 'inf' or 'nan' 
static char *put_py27_float(char *p, double v) {
    char tmp[40];
    int n = snprintf(tmp, sizeof(tmp), "%.12g", v);
    bool plain = true;
    for (int i = 0; i < n; i++) if (tmp[i] == '.' || tmp[i] == 'e' || tmp[i] == 'n') plain = false;
    memcpy(p, tmp, (size_t)n); p += n;
    if (plain) { *p++ = '.'; *p++ = '0'; }
    return p;
}

// get_phased_blocks/phased_variants (phasing.py:411-421) for the sites [s0, s1) of one contig: per block id 1 .. max (in
// ascending order, blocks without sites skipped) a P row "P pid min max span n span/n" and its sites in position order as
// "V pid pos pos_ref_b0 pos_ref_b1 lext rext lscore rscore" (b0 = the allele of the site's phase).  cap >= 64 + 160 bytes per site.
extern "C" int64_t fuz_host_format_phased_variants(const int32_t *site_pos, const uint8_t *site_al, const int32_t *ph_block,
                                                   const uint8_t *ph_state, const int32_t *ph_lext, const int32_t *ph_rext,
                                                   const int32_t *ph_lscore, const int32_t *ph_rscore, int64_t s0, int64_t s1,
                                                   const char *ref_seq, int64_t ref_len, char *out, int64_t cap) {
    if (!site_pos || !site_al || !ph_block || !ph_state || !ph_lext || !ph_rext || !ph_lscore || !ph_rscore || !ref_seq || !out || s0 > s1) return -1;
    if ((s1 - s0) * 224 + 64 > cap) return -1;
    static const char B[] = "ACGT";
    const int64_t n = s1 - s0;
    int32_t n_blocks = 0;
    for (int64_t i = s0; i < s1; i++) n_blocks = std::max(n_blocks, ph_block[i]);
    // sites grouped by block id, file order inside a block (counting sort = numpy's stable argsort)
    std::vector<int64_t> start((size_t)n_blocks + 2, 0);
    for (int64_t i = s0; i < s1; i++) if (ph_block[i] >= 1) start[(size_t)ph_block[i] + 1]++;
    for (int32_t b = 1; b <= n_blocks; b++) start[(size_t)b + 1] += start[(size_t)b];
    std::vector<int64_t> idx((size_t)n);
    {
        std::vector<int64_t> cur(start.begin(), start.end());
        for (int64_t i = s0; i < s1; i++) if (ph_block[i] >= 1) idx[(size_t)cur[(size_t)ph_block[i]]++] = i;
    }
    char *p = out;
    for (int32_t pid = 1; pid <= n_blocks; pid++) {
        const int64_t a = start[(size_t)pid], b = start[(size_t)pid + 1];
        if (a == b) continue;
        int32_t mn = site_pos[idx[(size_t)a]], mx = mn;
        for (int64_t k = a; k < b; k++) { const int32_t v = site_pos[idx[(size_t)k]]; mn = std::min(mn, v); mx = std::max(mx, v); }
        *p++ = 'P'; *p++ = ' '; p = put_int(p, pid); *p++ = ' '; p = put_int(p, mn); *p++ = ' '; p = put_int(p, mx); *p++ = ' ';
        p = put_int(p, (long long)mx - mn); *p++ = ' '; p = put_int(p, b - a); *p++ = ' ';
        p = put_py27_float(p, 1.0 * (double)((long long)mx - mn) / (double)(b - a)); *p++ = '\n';
        for (int64_t k = a; k < b; k++) {
            const int64_t i = idx[(size_t)k];
            const int32_t pos = site_pos[i];
            const int st = ph_state[i];
            if (pos < 1 || pos > ref_len || st > 1 || site_al[2 * i] > 3 || site_al[2 * i + 1] > 3) return -2;
            const char rb = ref_seq[pos - 1];
            *p++ = 'V'; *p++ = ' '; p = put_int(p, pid); *p++ = ' '; p = put_int(p, pos); *p++ = ' ';
            p = put_int(p, pos); *p++ = '_'; *p++ = rb; *p++ = '_'; *p++ = B[site_al[2 * i + st]]; *p++ = ' ';
            p = put_int(p, pos); *p++ = '_'; *p++ = rb; *p++ = '_'; *p++ = B[site_al[2 * i + 1 - st]]; *p++ = ' ';
            p = put_int(p, ph_lext[i]); *p++ = ' '; p = put_int(p, ph_rext[i]); *p++ = ' ';
            p = put_int(p, ph_lscore[i]); *p++ = ' '; p = put_int(p, ph_rscore[i]); *p++ = '\n';
        }
    }
    return p - out;
}
