/*
 * TEST INFRASTRUCTURE ONLY (oracle/).  Scalar C restatement of the reference's phasing
 * hot path, used as the parity checker for the CUDA kernels and as the "port" CPU
 * baseline in bench.py.  Nothing under falcon_unzip_b200/ may link or call this.
 *
 * Restates (reference = PacificBiosciences/FALCON_unzip, falcon_unzip/phasing.py):
 *   fo_make_het_call          phasing.py:14-134   (filter :63-75, pileup :77-96,
 *                                                  streaming flush :98-129)
 *   fo_association_table      phasing.py:137-206
 *   fo_phased_blocks          phasing.py:208-214 (get_score), :216-421
 *   fo_phased_reads           phasing.py:423-480
 * with the Python-2 semantics of SURVEY.md Appendix B (allele order "ACTG" B.1, the
 * float clip/het tests in IEEE double in the reference's operation order B.2).
 *
 * Pinned against the reference's own source executed under Python 3 (oracle/ref_exec.py)
 * by tests/test_oracle_vs_reference.py and the fixtures under tests/golden/; the
 * reference ships no golden vectors of its own.
 *
 * Input of fo_make_het_call is the concatenated uncompressed BAM alignment records (the
 * same bytes the CUDA path reads); the SAM text the reference parses is a rendering of
 * exactly these fields (QNAME->q_id is done by the caller, phasing.py:47-54).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define FO_OK 0
#define FO_E_NOMEM 1
#define FO_E_BADRECORD 2   /* reference would raise (ZeroDivisionError / IndexError) */
#define FO_E_BADINPUT 3

/* ------------------------------------------------------------------ growable vectors */
typedef struct { int32_t *p; int64_t n, cap; } ivec;
typedef struct { uint8_t *p; int64_t n, cap; } bvec;

static int ivec_push(ivec *v, int32_t x) {
    if (v->n == v->cap) {
        int64_t nc = v->cap ? v->cap * 2 : 16;
        int32_t *np_ = (int32_t *)realloc(v->p, (size_t)nc * sizeof(int32_t));
        if (!np_) return FO_E_NOMEM;
        v->p = np_; v->cap = nc;
    }
    v->p[v->n++] = x;
    return FO_OK;
}
static int bvec_push(bvec *v, uint8_t x) {
    if (v->n == v->cap) {
        int64_t nc = v->cap ? v->cap * 2 : 16;
        uint8_t *np_ = (uint8_t *)realloc(v->p, (size_t)nc);
        if (!np_) return FO_E_NOMEM;
        v->p = np_; v->cap = nc;
    }
    v->p[v->n++] = x;
    return FO_OK;
}

static inline int32_t rd_i32(const uint8_t *p) { int32_t v; memcpy(&v, p, 4); return v; }
static inline uint32_t rd_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t rd_u16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

/* ================================================================== make_het_call */
typedef struct {
    int64_t n_sites;
    int32_t *site_pos;     /* 0-based position                                    */
    int32_t *site_total;   /* A+C+G+T depth (phasing.py:111)                      */
    uint8_t *site_base;    /* 4 per site: letters sorted as phasing.py:116-117    */
    int32_t *site_count;   /* 4 per site: counts in the same order                */
    int64_t n_vmap;
    int32_t *vm_pos;       /* 0-based                                             */
    uint8_t *vm_allele;    /* letter                                              */
    int32_t *vm_qid;
    int64_t n_accepted;    /* records passing the filter (phasing.py:72-75)       */
    int64_t aligned_bases; /* sum of M/=/X lengths over accepted records          */
    int32_t pos_last;      /* start of the last accepted record, -1 if none       */
} fo_hetcall_result;

/* pileup[pos][symbol] -> list of q_ids (phasing.py:88-90); 16 BAM symbols */
typedef struct { ivec sym[16]; uint16_t mask; } cell_t;



static void cell_clear(cell_t *c) {
    for (int s = 0; s < 16; s++) { free(c->sym[s].p); c->sym[s].p = NULL; c->sym[s].n = c->sym[s].cap = 0; }
    c->mask = 0;
}

typedef struct { ivec spos, stot, scnt, vpos, vqid; bvec sbase, vall; } het_out;

/* phasing.py:103-129 for one live position */
static int evaluate_pos(cell_t *c, int32_t pos, het_out *o) {
    int distinct = __builtin_popcount(c->mask);
    if (distinct < 2) return FO_OK;                            /* :103-105 */
    static const int acgt_sym[4] = {1, 2, 4, 8};               /* A C G T  */
    static const char acgt_chr[4] = {'A', 'C', 'G', 'T'};
    int64_t cnt[4]; char base[4]; int64_t total = 0;
    for (int b = 0; b < 4; b++) { cnt[b] = c->sym[acgt_sym[b]].n; base[b] = acgt_chr[b]; total += cnt[b]; }
    if (total < 10) return FO_OK;                              /* :112-114 */
    /* :116-117 sort() then reverse() on (count, base) tuples: descending count,
       ties by descending base letter */
    for (int i = 0; i < 4; i++)
        for (int j = i + 1; j < 4; j++)
            if (cnt[j] > cnt[i] || (cnt[j] == cnt[i] && base[j] > base[i])) {
                int64_t t = cnt[i]; cnt[i] = cnt[j]; cnt[j] = t;
                char tb = base[i]; base[i] = base[j]; base[j] = tb;
            }
    double p0 = 1.0 * (double)cnt[0] / (double)total;          /* :118-119 */
    double p1 = 1.0 * (double)cnt[1] / (double)total;
    double th = 0.25;
    if (!(p0 < 1.0 - th && p1 > th)) return FO_OK;             /* :120 */
    int rc = ivec_push(&o->spos, pos);
    rc |= ivec_push(&o->stot, (int32_t)total);
    for (int i = 0; i < 4; i++) { rc |= bvec_push(&o->sbase, (uint8_t)base[i]); rc |= ivec_push(&o->scnt, (int32_t)cnt[i]); }
    for (int k = 0; k < 2; k++) {                              /* :125-128 */
        int sym = base[k] == 'A' ? 1 : base[k] == 'C' ? 2 : base[k] == 'G' ? 4 : 8;
        ivec *l = &c->sym[sym];
        for (int64_t i = 0; i < l->n; i++) {
            rc |= ivec_push(&o->vpos, pos);
            rc |= bvec_push(&o->vall, (uint8_t)base[k]);
            rc |= ivec_push(&o->vqid, l->p[i]);
        }
    }
    return rc ? FO_E_NOMEM : FO_OK;
}

int fo_make_het_call(const uint8_t *recs, const int64_t *rec_off, int64_t n_rec,
                     const int32_t *qid, fo_hetcall_result *res) {
    memset(res, 0, sizeof(*res));
    res->pos_last = -1;
    /* position-indexed pileup table; grown on demand */
    int64_t ncell = 0; cell_t *cells = NULL;
    int64_t min_live = -1, max_live = -1;                      /* live window (superset) */
    het_out o; memset(&o, 0, sizeof(o));
    int rc = FO_OK;
    for (int64_t r = 0; r < n_rec && rc == FO_OK; r++) {
        const uint8_t *rec = recs + rec_off[r];
        int32_t pos = rd_i32(rec + 8);
        int l_name = rec[12];
        int n_cig = rd_u16(rec + 16);
        int32_t l_seq = rd_i32(rec + 20);
        const uint8_t *cig = rec + 36 + l_name;
        const uint8_t *seq = cig + 4 * (int64_t)n_cig;
        /* :63-75 */
        int64_t skip = 0, total = 0;
        for (int k = 0; k < n_cig; k++) {
            uint32_t c = rd_u32(cig + 4 * k); int64_t adv = c >> 4; int op = c & 15;
            if (op > 8) { rc = FO_E_BADRECORD; break; }
            total += adv;
            if (op == 4) skip += adv;
        }
        if (rc) break;
        if (total == 0) { rc = FO_E_BADRECORD; break; }        /* ZeroDivisionError :72 */
        if (1.0 - 1.0 * (double)skip / (double)total < 0.1) continue;
        if (total < 2000) continue;
        /* :77-96 */
        int64_t rp = pos, qp = 0;
        for (int k = 0; k < n_cig && rc == FO_OK; k++) {
            uint32_t c = rd_u32(cig + 4 * k); int64_t adv = c >> 4; int op = c & 15;
            if (op == 4) qp += adv;                            /* S */
            if (op == 0 || op == 7 || op == 8) {               /* M = X */
                if (qp + adv > l_seq || rp < 0) { rc = FO_E_BADRECORD; break; } /* IndexError :84 */
                if (rp + adv > ncell) {
                    int64_t nn = ncell ? ncell : 1 << 16;
                    while (nn < rp + adv) nn *= 2;
                    cell_t *nc = (cell_t *)realloc(cells, (size_t)nn * sizeof(cell_t));
                    if (!nc) { rc = FO_E_NOMEM; break; }
                    memset(nc + ncell, 0, (size_t)(nn - ncell) * sizeof(cell_t));
                    cells = nc; ncell = nn;
                }
                if (adv > 0) {
                    if (min_live < 0 || rp < min_live) min_live = rp;
                    if (rp + adv - 1 > max_live) max_live = rp + adv - 1;
                }
                for (int64_t i = 0; i < adv; i++) {
                    int sym = (seq[qp >> 1] >> ((qp & 1) ? 0 : 4)) & 15;
                    cell_t *ce = &cells[rp];
                    ce->mask |= (uint16_t)(1u << sym);
                    if (ivec_push(&ce->sym[sym], qid[r])) { rc = FO_E_NOMEM; break; }
                    rp++; qp++;
                }
                res->aligned_bases += adv;
            } else if (op == 1) {                              /* I */
                qp += adv;
            } else if (op == 2) {                              /* D */
                rp += adv;
            }                                                  /* N H P: nothing (quirk) */
        }
        if (rc) break;
        res->n_accepted++;
        res->pos_last = pos;
        /* :98-129 flush every live position < POS in ascending order */
        if (min_live >= 0) {
            int64_t hi = pos < max_live + 1 ? pos : max_live + 1;
            for (int64_t p = min_live; p < hi && rc == FO_OK; p++) {
                cell_t *ce = &cells[p];
                if (ce->mask) { rc = evaluate_pos(ce, (int32_t)p, &o); cell_clear(ce); }
            }
            if (hi > min_live) min_live = hi;
            if (min_live > max_live) { min_live = -1; max_live = -1; }
        }
    }
    if (cells) {
        for (int64_t p = 0; p < ncell; p++) if (cells[p].mask) cell_clear(&cells[p]);
        free(cells);
    }
    if (rc) {
        free(o.spos.p); free(o.stot.p); free(o.scnt.p); free(o.vpos.p); free(o.vqid.p);
        free(o.sbase.p); free(o.vall.p);
        return rc;
    }
    res->n_sites = o.spos.n; res->site_pos = o.spos.p; res->site_total = o.stot.p;
    res->site_base = o.sbase.p; res->site_count = o.scnt.p;
    res->n_vmap = o.vpos.n; res->vm_pos = o.vpos.p; res->vm_allele = o.vall.p; res->vm_qid = o.vqid.p;
    return FO_OK;
}

void fo_free_hetcall(fo_hetcall_result *r) {
    free(r->site_pos); free(r->site_total); free(r->site_base); free(r->site_count);
    free(r->vm_pos); free(r->vm_allele); free(r->vm_qid);
    memset(r, 0, sizeof(*r));
}

/* ================================================================== association table */
typedef struct {
    int64_t n_rows;
    int32_t *pos1, *pos2;      /* as in the file (1-based)                        */
    uint8_t *b;                /* 4 per row: b11 b12 b21 b22                      */
    int32_t *ct;               /* 4 per row: c11 c12 c21 c22                      */
} fo_atable_result;

static int actg_index(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'T' ? 2 : c == 'G' ? 3 : -1; }

static int cmp_i32(const void *a, const void *b) {
    int32_t x = *(const int32_t *)a, y = *(const int32_t *)b; return (x > y) - (x < y);
}
/* set(list): sort + unique in place, returns new length */
static int64_t make_set(int32_t *a, int64_t n) {
    if (n == 0) return 0;
    qsort(a, (size_t)n, sizeof(int32_t), cmp_i32);
    int64_t m = 1;
    for (int64_t i = 1; i < n; i++) if (a[i] != a[m - 1]) a[m++] = a[i];
    return m;
}
static int32_t set_intersect(const int32_t *a, int64_t na, const int32_t *b, int64_t nb) {
    int64_t i = 0, j = 0; int32_t s = 0;
    while (i < na && j < nb) { if (a[i] < b[j]) i++; else if (a[i] > b[j]) j++; else { s++; i++; j++; } }
    return s;
}

typedef struct { int32_t pos; uint8_t allele[2]; int64_t off[2], n[2]; } site_t;

/* Groups vmap rows by site in file order (phasing.py:147-158).  Each site must carry
 * exactly two alleles (always true for make_het_call output); the two alleles are
 * ordered by the string "ACTG" = CPython 2.7 dict order of 1-char keys (B.1). */
static int group_sites(const int32_t *vm_pos, const uint8_t *vm_allele, const int32_t *vm_qid,
                       int64_t n, site_t **sites_out, int64_t *n_sites_out, int32_t **sets_out) {
    site_t *sites = NULL; int64_t ns = 0, cap = 0;
    int32_t *sets = (int32_t *)malloc((size_t)(n ? n : 1) * sizeof(int32_t));
    if (!sets) return FO_E_NOMEM;
    int64_t i = 0, w = 0; int32_t max_pos = 0;
    while (i < n) {
        int64_t j = i;
        while (j < n && vm_pos[j] == vm_pos[i]) j++;
        /* a position appearing in two separate runs would be one dict key upstream */
        if (ns > 0 && vm_pos[i] <= max_pos)
            for (int64_t s = 0; s < ns; s++) if (sites[s].pos == vm_pos[i]) { free(sites); free(sets); return FO_E_BADINPUT; }
        if (ns == 0 || vm_pos[i] > max_pos) max_pos = vm_pos[i];
        uint8_t al[2]; int na = 0;
        for (int64_t k = i; k < j; k++) {
            int found = 0;
            for (int a = 0; a < na; a++) if (al[a] == vm_allele[k]) found = 1;
            if (!found) { if (na == 2 || actg_index(vm_allele[k]) < 0) { free(sites); free(sets); return FO_E_BADINPUT; } al[na++] = vm_allele[k]; }
        }
        if (na != 2) { free(sites); free(sets); return FO_E_BADINPUT; }
        if (actg_index(al[0]) > actg_index(al[1])) { uint8_t t = al[0]; al[0] = al[1]; al[1] = t; }
        if (ns == cap) {
            cap = cap ? cap * 2 : 256;
            site_t *np_ = (site_t *)realloc(sites, (size_t)cap * sizeof(site_t));
            if (!np_) { free(sites); free(sets); return FO_E_NOMEM; }
            sites = np_;
        }
        site_t *st = &sites[ns++];
        st->pos = vm_pos[i]; st->allele[0] = al[0]; st->allele[1] = al[1];
        for (int a = 0; a < 2; a++) {
            st->off[a] = w;
            for (int64_t k = i; k < j; k++) if (vm_allele[k] == al[a]) sets[w++] = vm_qid[k];
            st->n[a] = make_set(sets + st->off[a], w - st->off[a]);   /* set(qids) :189 */
            w = st->off[a] + st->n[a];
        }
        i = j;
    }
    *sites_out = sites; *n_sites_out = ns; *sets_out = sets;
    return FO_OK;
}

int fo_association_table(const int32_t *vm_pos, const uint8_t *vm_allele, const int32_t *vm_qid,
                         int64_t n, fo_atable_result *res) {
    memset(res, 0, sizeof(*res));
    site_t *sites; int64_t ns; int32_t *sets;
    int rc = group_sites(vm_pos, vm_allele, vm_qid, n, &sites, &ns, &sets);
    if (rc) return rc;
    ivec p1 = {0}, p2 = {0}, ct = {0}; bvec bb = {0};
    int sorted = 1;
    for (int64_t i = 1; i < ns; i++) if (sites[i].pos < sites[i - 1].pos) sorted = 0;
    for (int64_t i1 = 0; i1 < ns && !rc; i1++) {               /* :164 */
        int link_count = 0;
        for (int64_t i2 = i1 + 1; i2 < ns; i2++) {             /* :166 */
            if ((int64_t)sites[i2].pos - sites[i1].pos > (1 << 16)) {   /* :169-170 */
                if (sorted) break;                             /* same rows as `continue` */
                continue;
            }
            int32_t c[4]; int32_t total_s = 0;
            for (int a = 0; a < 2; a++)
                for (int b = 0; b < 2; b++) {                  /* :187-191 */
                    c[a * 2 + b] = set_intersect(sets + sites[i1].off[a], sites[i1].n[a],
                                                 sets + sites[i2].off[b], sites[i2].n[b]);
                    total_s += c[a * 2 + b];
                }
            if (total_s < 6) continue;                         /* :192-193 */
            rc |= ivec_push(&p1, sites[i1].pos); rc |= ivec_push(&p2, sites[i2].pos);
            rc |= bvec_push(&bb, sites[i1].allele[0]); rc |= bvec_push(&bb, sites[i1].allele[1]);
            rc |= bvec_push(&bb, sites[i2].allele[0]); rc |= bvec_push(&bb, sites[i2].allele[1]);
            for (int k = 0; k < 4; k++) rc |= ivec_push(&ct, c[k]);
            link_count++;
            if (link_count > 500) break;                       /* :204-206 */
        }
    }
    free(sites); free(sets);
    if (rc) { free(p1.p); free(p2.p); free(ct.p); free(bb.p); return FO_E_NOMEM; }
    res->n_rows = p1.n; res->pos1 = p1.p; res->pos2 = p2.p; res->b = bb.p; res->ct = ct.p;
    return FO_OK;
}

void fo_free_atable(fo_atable_result *r) {
    free(r->pos1); free(r->pos2); free(r->b); free(r->ct); memset(r, 0, sizeof(*r));
}

/* ================================================================== phased blocks */
typedef struct {
    int64_t n_v;               /* V rows, in output order (block id, then position) */
    int32_t *pid, *pos;
    uint8_t *h;                /* 2 per row: hap-0 allele, hap-1 allele             */
    int32_t *lext, *rext, *lscore, *rscore;
    int32_t n_blocks;
} fo_blocks_result;

typedef struct {
    int32_t pos; int stated; uint8_t st[2];        /* states[pos] = (hap0, hap1)   */
    ivec left, right;                              /* edge ids, append order        */
    int32_t lext, rext, lscore, rscore;
} node_t;
typedef struct { int32_t p1, p2; int32_t n1, n2; uint8_t b11, b12, b21, b22; int32_t cis, trans; } edge_t;

/* get_score (phasing.py:208-214): c_score[(p1,p2)][(s1[0]+s2[0], s1[1]+s2[1])] with
 * the four keys built at :255-256; s1 is the state of the lower position. */
static int32_t edge_score(const edge_t *e, const uint8_t s1[2], const uint8_t s2[2]) {
    if ((s1[0] == e->b11 && s2[0] == e->b21 && s1[1] == e->b12 && s2[1] == e->b22) ||
        (s1[0] == e->b12 && s2[0] == e->b22 && s1[1] == e->b11 && s2[1] == e->b21)) return e->cis;
    return e->trans;  /* (b12+b21, b11+b22) / (b11+b22, b12+b21) */
}

static int cmp_node_pos(const void *a, const void *b) {
    int32_t x = ((const node_t *)a)->pos, y = ((const node_t *)b)->pos; return (x > y) - (x < y);
}
static int64_t find_node(const node_t *nodes, int64_t n, int32_t pos) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t m = (lo + hi) / 2; if (nodes[m].pos < pos) lo = m + 1; else hi = m; }
    return (lo < n && nodes[lo].pos == pos) ? lo : -1;
}

/* greedy first-touch state of node x (phasing.py:259-283 / :285-309) */
static void init_state(node_t *nodes, const edge_t *edges, int64_t x, uint8_t b1, uint8_t b2) {
    node_t *nx = &nodes[x];
    uint8_t st1[2] = {b1, b2}, st2[2] = {b2, b1};
    int64_t score1 = 0, score2 = 0;
    for (int64_t k = 0; k < nx->left.n; k++) {
        const edge_t *e = &edges[nx->left.p[k]]; const node_t *pp = &nodes[e->n1];
        if (!pp->stated) continue;
        score1 += edge_score(e, pp->st, st1); score2 += edge_score(e, pp->st, st2);
    }
    for (int64_t k = 0; k < nx->right.n; k++) {
        const edge_t *e = &edges[nx->right.p[k]]; const node_t *pp = &nodes[e->n2];
        if (!pp->stated) continue;
        score1 += edge_score(e, st1, pp->st); score2 += edge_score(e, st2, pp->st);
    }
    if (score1 >= score2) { nx->st[0] = b1; nx->st[1] = b2; } else { nx->st[0] = b2; nx->st[1] = b1; }
    nx->stated = 1;
}

int fo_phased_blocks(const int32_t *pos1, const int32_t *pos2, const uint8_t *b, const int32_t *ct,
                     int64_t n_rows, fo_blocks_result *res) {
    memset(res, 0, sizeof(*res));
    /* nodes = positions appearing in accepted rows (:245-250), sorted (:311-312) */
    node_t *nodes = (node_t *)calloc((size_t)(2 * n_rows + 1), sizeof(node_t));
    edge_t *edges = (edge_t *)calloc((size_t)(n_rows + 1), sizeof(edge_t));
    if (!nodes || !edges) { free(nodes); free(edges); return FO_E_NOMEM; }
    int64_t nn = 0, ne = 0;
    for (int64_t r = 0; r < n_rows; r++) {
        int32_t cis = ct[4 * r] + ct[4 * r + 3], trans = ct[4 * r + 1] + ct[4 * r + 2];
        if (abs(cis - trans) < 6) continue;                    /* :245 */
        nodes[nn++].pos = pos1[r]; nodes[nn++].pos = pos2[r];
    }
    qsort(nodes, (size_t)nn, sizeof(node_t), cmp_node_pos);
    { int64_t m = 0; for (int64_t i = 0; i < nn; i++) if (m == 0 || nodes[i].pos != nodes[m - 1].pos) nodes[m++] = nodes[i]; nn = m; }
    int rc = FO_OK;
    /* pass 1 in file order (:240-309) */
    for (int64_t r = 0; r < n_rows && !rc; r++) {
        int32_t cis = ct[4 * r] + ct[4 * r + 3], trans = ct[4 * r + 1] + ct[4 * r + 2];
        if (abs(cis - trans) < 6) continue;
        edge_t *e = &edges[ne];
        e->p1 = pos1[r]; e->p2 = pos2[r];
        e->n1 = (int32_t)find_node(nodes, nn, pos1[r]); e->n2 = (int32_t)find_node(nodes, nn, pos2[r]);
        e->b11 = b[4 * r]; e->b12 = b[4 * r + 1]; e->b21 = b[4 * r + 2]; e->b22 = b[4 * r + 3];
        e->cis = cis; e->trans = trans;
        /* a repeated (pos1,pos2) key overwrites c_score upstream (:255); the lists keep
           both entries.  Restating that would need key lookups; reject instead. */
        for (int64_t k = 0; k < nodes[e->n1].right.n; k++)
            if (edges[nodes[e->n1].right.p[k]].p2 == e->p2) rc = FO_E_BADINPUT;
        if (e->p1 >= e->p2) rc = FO_E_BADINPUT;                /* get_score swaps on pos1 > pos2 */
        if (rc) break;
        rc |= ivec_push(&nodes[e->n1].right, (int32_t)ne);     /* :251-254 */
        rc |= ivec_push(&nodes[e->n2].left, (int32_t)ne);
        if (rc) { rc = FO_E_NOMEM; break; }
        ne++;
        if (!nodes[e->n1].stated) init_state(nodes, edges, e->n1, e->b11, e->b12);
        if (!nodes[e->n2].stated) init_state(nodes, edges, e->n2, e->b21, e->b22);
    }
    /* pass 2 (:315-344): up to 10 sweeps, left neighbours only */
    for (int iter = 1; iter <= 10 && !rc; iter++) {
        int64_t update = 0;
        for (int64_t x = 0; x < nn; x++) {
            node_t *nx = &nodes[x];
            uint8_t st1[2] = {nx->st[0], nx->st[1]}, st2[2] = {nx->st[1], nx->st[0]};
            int64_t score1 = 0, score2 = 0;
            for (int64_t k = 0; k < nx->left.n; k++) {
                const edge_t *e = &edges[nx->left.p[k]]; const node_t *pp = &nodes[e->n1];
                score1 += edge_score(e, pp->st, st1); score2 += edge_score(e, pp->st, st2);
            }
            if (!(score1 >= score2)) { nx->st[0] = st2[0]; nx->st[1] = st2[1]; update++; }
        }
        if (update == 0) break;
    }
    /* pass 3 (:353-383) */
    for (int64_t x = 0; x < nn && !rc; x++) {
        node_t *nx = &nodes[x];
        uint8_t st0[2] = {nx->st[0], nx->st[1]}, st0_[2] = {nx->st[1], nx->st[0]};
        nx->lext = nx->pos; nx->lscore = 0; nx->rext = nx->pos; nx->rscore = 0;
        for (int64_t k = 0; k < nx->left.n; k++) {
            const edge_t *e = &edges[nx->left.p[k]]; const node_t *pp = &nodes[e->n1];
            int32_t s = edge_score(e, pp->st, st0), s_ = edge_score(e, pp->st, st0_);
            nx->lscore += s - s_;
            if (s - s_ > 0 && pp->pos < nx->lext) nx->lext = pp->pos;
        }
        for (int64_t k = 0; k < nx->right.n; k++) {
            const edge_t *e = &edges[nx->right.p[k]]; const node_t *pp = &nodes[e->n2];
            int32_t s = edge_score(e, st0, pp->st), s_ = edge_score(e, st0_, pp->st);
            nx->rscore += s - s_;
            if (s - s_ > 0 && pp->pos > nx->rext) nx->rext = pp->pos;
        }
    }
    /* pass 4 (:388-408) + emission order (:411-421) */
    ivec o_pid = {0}, o_node = {0};
    if (!rc) {
        int32_t block_id = 1; int64_t max_right_ext = 0;
        int64_t pb_start = 0;                                  /* start of current pb in o_node */
        for (int64_t x = 0; x < nn && !rc; x++) {
            node_t *nx = &nodes[x];
            if (nx->rscore < 10 || nx->lscore < 10) continue;  /* :394 */
            if (max_right_ext < nx->lext) {                    /* :397-401 */
                if (o_node.n - pb_start > 3) block_id++; else { o_node.n = pb_start; o_pid.n = pb_start; }
                pb_start = o_node.n;
            }
            rc |= ivec_push(&o_node, (int32_t)x); rc |= ivec_push(&o_pid, block_id);
            if (nx->rext > max_right_ext) max_right_ext = nx->rext;
        }
        if (o_node.n - pb_start > 3) res->n_blocks = block_id;
        else { o_node.n = pb_start; o_pid.n = pb_start; res->n_blocks = block_id - 1; }
        if (rc) rc = FO_E_NOMEM;
    }
    if (!rc) {
        int64_t nv = o_node.n;
        res->n_v = nv;
        res->pid = (int32_t *)malloc((size_t)(nv + 1) * 4); res->pos = (int32_t *)malloc((size_t)(nv + 1) * 4);
        res->h = (uint8_t *)malloc((size_t)(2 * nv + 1));
        res->lext = (int32_t *)malloc((size_t)(nv + 1) * 4); res->rext = (int32_t *)malloc((size_t)(nv + 1) * 4);
        res->lscore = (int32_t *)malloc((size_t)(nv + 1) * 4); res->rscore = (int32_t *)malloc((size_t)(nv + 1) * 4);
        if (!res->pid || !res->pos || !res->h || !res->lext || !res->rext || !res->lscore || !res->rscore) rc = FO_E_NOMEM;
        for (int64_t i = 0; i < nv && !rc; i++) {
            const node_t *nx = &nodes[o_node.p[i]];
            res->pid[i] = o_pid.p[i]; res->pos[i] = nx->pos;
            res->h[2 * i] = nx->st[0]; res->h[2 * i + 1] = nx->st[1];
            res->lext[i] = nx->lext; res->rext[i] = nx->rext; res->lscore[i] = nx->lscore; res->rscore[i] = nx->rscore;
        }
    }
    for (int64_t x = 0; x < nn; x++) { free(nodes[x].left.p); free(nodes[x].right.p); }
    free(nodes); free(edges); free(o_pid.p); free(o_node.p);
    return rc;
}

void fo_free_blocks(fo_blocks_result *r) {
    free(r->pid); free(r->pos); free(r->h); free(r->lext); free(r->rext); free(r->lscore); free(r->rscore);
    memset(r, 0, sizeof(*r));
}

/* ================================================================== phased reads */
typedef struct {
    int64_t n_rows;            /* grouped by q_id in first-appearance order (the caller
                                  applies the py2 dict order, B.3), block id ascending */
    int32_t *qid, *pid, *phase, *n0, *n1;
} fo_reads_result;

typedef struct { int32_t qid; int32_t pos; uint8_t allele; int64_t first; } rv_t;
static int cmp_rv(const void *a, const void *b) {
    const rv_t *x = (const rv_t *)a, *y = (const rv_t *)b;
    if (x->qid != y->qid) return (x->qid > y->qid) - (x->qid < y->qid);
    if (x->pos != y->pos) return (x->pos > y->pos) - (x->pos < y->pos);
    return (x->allele > y->allele) - (x->allele < y->allele);
}
typedef struct { int32_t qid; int64_t first; int64_t lo, hi; } rq_t;
static int cmp_rq_first(const void *a, const void *b) {
    int64_t x = ((const rq_t *)a)->first, y = ((const rq_t *)b)->first; return (x > y) - (x < y);
}
typedef struct { int32_t pos; uint8_t allele; int32_t pid; int32_t phase; } vp_t;
static int cmp_vp(const void *a, const void *b) {
    const vp_t *x = (const vp_t *)a, *y = (const vp_t *)b;
    if (x->pos != y->pos) return (x->pos > y->pos) - (x->pos < y->pos);
    return (x->allele > y->allele) - (x->allele < y->allele);
}

/* vmap rows (pos 1-based as in the file) + V rows -> phased_reads rows.
 * variant identity "pos_ref_allele" (:446) == (pos, allele) because ref is a function
 * of pos.  read_to_variants[q] is a *set* (:448-449); variant_to_phase: later V rows
 * overwrite earlier ones (:462-463). */
int fo_phased_reads(const int32_t *vm_pos, const uint8_t *vm_allele, const int32_t *vm_qid, int64_t n,
                    const int32_t *v_pid, const int32_t *v_pos, const uint8_t *v_h, int64_t n_v,
                    fo_reads_result *res) {
    memset(res, 0, sizeof(*res));
    rv_t *rv = (rv_t *)malloc((size_t)(n + 1) * sizeof(rv_t));
    vp_t *vp = (vp_t *)malloc((size_t)(2 * n_v + 1) * sizeof(vp_t));
    rq_t *rq = (rq_t *)malloc((size_t)(n + 1) * sizeof(rq_t));
    if (!rv || !vp || !rq) { free(rv); free(vp); free(rq); return FO_E_NOMEM; }
    for (int64_t i = 0; i < n; i++) { rv[i].qid = vm_qid[i]; rv[i].pos = vm_pos[i]; rv[i].allele = vm_allele[i]; rv[i].first = i; }
    qsort(rv, (size_t)n, sizeof(rv_t), cmp_rv);
    int64_t nvp = 0;
    for (int64_t i = 0; i < n_v; i++) {
        /* dict assignment order: [l[3]] = (pid,0) then [l[4]] = (pid,1); emulate
           overwrite by keeping the LAST entry per key after a stable pass */
        vp[nvp].pos = v_pos[i]; vp[nvp].allele = v_h[2 * i]; vp[nvp].pid = v_pid[i]; vp[nvp].phase = 0; nvp++;
        vp[nvp].pos = v_pos[i]; vp[nvp].allele = v_h[2 * i + 1]; vp[nvp].pid = v_pid[i]; vp[nvp].phase = 1; nvp++;
    }
    /* stable insertion of "last wins": mark superseded entries */
    /* (V rows from get_phased_blocks never repeat a key; handle generally) */
    {
        /* mergesort-free approach: add sequence number into phase's high bits */
        for (int64_t i = 0; i < nvp; i++) vp[i].phase |= (int32_t)((i & 0x3FFFFFFF) << 1);
        qsort(vp, (size_t)nvp, sizeof(vp_t), cmp_vp);
        int64_t m = 0;
        for (int64_t i = 0; i < nvp; ) {
            int64_t j = i, best = i;
            while (j < nvp && vp[j].pos == vp[i].pos && vp[j].allele == vp[i].allele) { if ((vp[j].phase >> 1) > (vp[best].phase >> 1)) best = j; j++; }
            vp[m] = vp[best]; vp[m].phase &= 1; m++;
            i = j;
        }
        nvp = m;
    }
    /* per q_id groups, then order by first appearance */
    int64_t nq = 0;
    for (int64_t i = 0; i < n; ) {
        int64_t j = i; int64_t first = rv[i].first;
        while (j < n && rv[j].qid == rv[i].qid) { if (rv[j].first < first) first = rv[j].first; j++; }
        rq[nq].qid = rv[i].qid; rq[nq].first = first; rq[nq].lo = i; rq[nq].hi = j; nq++;
        i = j;
    }
    qsort(rq, (size_t)nq, sizeof(rq_t), cmp_rq_first);
    ivec oq = {0}, op = {0}, oph = {0}, o0 = {0}, o1 = {0};
    int rc = 0;
    for (int64_t g = 0; g < nq && !rc; g++) {
        /* distinct variants of this read, sorted by (pos, allele); collect (pid, phase) */
        int64_t lo = rq[g].lo, hi = rq[g].hi;
        int32_t *pp = (int32_t *)malloc((size_t)(hi - lo) * 2 * sizeof(int32_t)); int64_t np_ = 0;
        if (!pp) { rc = 1; break; }
        for (int64_t i = lo; i < hi; i++) {
            if (i > lo && rv[i].pos == rv[i - 1].pos && rv[i].allele == rv[i - 1].allele) continue; /* set */
            vp_t key; key.pos = rv[i].pos; key.allele = rv[i].allele;
            vp_t *f = (vp_t *)bsearch(&key, vp, (size_t)nvp, sizeof(vp_t), cmp_vp);
            if (!f) continue;                                  /* :470 */
            pp[2 * np_] = f->pid; pp[2 * np_ + 1] = f->phase; np_++;
        }
        /* pl sorted (:474-475): iterate distinct pids ascending */
        int32_t last_pid = 0; int have_last = 0;
        for (;;) {
            int32_t cur = 0; int found = 0;
            for (int64_t k = 0; k < np_; k++) {
                int32_t pid = pp[2 * k];
                if (have_last && pid <= last_pid) continue;
                if (!found || pid < cur) { cur = pid; found = 1; }
            }
            if (!found) break;
            int32_t c0 = 0, c1 = 0;
            for (int64_t k = 0; k < np_; k++) if (pp[2 * k] == cur) { if (pp[2 * k + 1] == 0) c0++; else c1++; }
            int ph = -1;
            if (c0 - c1 > 1) ph = 0; else if (c1 - c0 > 1) ph = 1;         /* :477-480 */
            if (ph >= 0) {
                rc |= ivec_push(&oq, rq[g].qid); rc |= ivec_push(&op, cur); rc |= ivec_push(&oph, ph);
                rc |= ivec_push(&o0, c0); rc |= ivec_push(&o1, c1);
            }
            last_pid = cur; have_last = 1;
        }
        free(pp);
    }
    free(rv); free(vp); free(rq);
    if (rc) { free(oq.p); free(op.p); free(oph.p); free(o0.p); free(o1.p); return FO_E_NOMEM; }
    res->n_rows = oq.n; res->qid = oq.p; res->pid = op.p; res->phase = oph.p; res->n0 = o0.p; res->n1 = o1.p;
    return FO_OK;
}

void fo_free_reads(fo_reads_result *r) {
    free(r->qid); free(r->pid); free(r->phase); free(r->n0); free(r->n1); memset(r, 0, sizeof(*r));
}

/* ================================================================== full pileup (tests)
 * Depth of A,C,G,T at every position [0, ctg_len) over the accepted records: the
 * pileup of phasing.py:77-96 without the streaming flush.  counts[4*pos + b]. */
int fo_pileup_counts(const uint8_t *recs, const int64_t *rec_off, int64_t n_rec, int64_t ctg_len,
                     uint32_t *counts) {
    memset(counts, 0, (size_t)ctg_len * 4 * sizeof(uint32_t));
    for (int64_t r = 0; r < n_rec; r++) {
        const uint8_t *rec = recs + rec_off[r];
        int32_t pos = rd_i32(rec + 8);
        int l_name = rec[12];
        int n_cig = rd_u16(rec + 16);
        const uint8_t *cig = rec + 36 + l_name;
        const uint8_t *seq = cig + 4 * (int64_t)n_cig;
        int64_t skip = 0, total = 0;
        for (int k = 0; k < n_cig; k++) {
            uint32_t c = rd_u32(cig + 4 * k);
            total += c >> 4;
            if ((c & 15) == 4) skip += c >> 4;
        }
        if (total == 0) return FO_E_BADRECORD;
        if (1.0 - 1.0 * (double)skip / (double)total < 0.1) continue;
        if (total < 2000) continue;
        int64_t rp = pos, qp = 0;
        for (int k = 0; k < n_cig; k++) {
            uint32_t c = rd_u32(cig + 4 * k); int64_t adv = c >> 4; int op = c & 15;
            if (op == 4 || op == 1) qp += adv;
            else if (op == 2) rp += adv;
            else if (op == 0 || op == 7 || op == 8) {
                for (int64_t i = 0; i < adv; i++, rp++, qp++) {
                    int sym = (seq[qp >> 1] >> ((qp & 1) ? 0 : 4)) & 15;
                    int b = sym == 1 ? 0 : sym == 2 ? 1 : sym == 4 ? 2 : sym == 8 ? 3 : -1;
                    if (b >= 0 && rp >= 0 && rp < ctg_len) counts[4 * rp + b]++;
                }
            }
        }
    }
    return FO_OK;
}
