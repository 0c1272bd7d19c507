// QNAME -> q_id on the device (reference falcon_unzip/phasing.py:47-54): q_ids are handed out
// per contig in order of first appearance of the read name, before any filtering.
//
//   k_qid_insert  one thread per record: hash (contig, name) into an open-addressing table whose
//                 slots hold a record index; records with the same key meet in one slot, which
//                 keeps the SMALLEST record index (atomicCAS to claim, atomicMin to lower);
//   k_qid_first   one thread per record: find the slot again; rep = its record; first = (rep == r);
//   scan          exclusive scan of the first flags: rank of every first occurrence in file order;
//   k_qid_assign  q_id = rank(rep) - rank(first record of the contig); name_first, ctg_nq.
//
// Slots never empty again and a key always probes the same sequence, so a key lives in exactly
// one slot; which record represents it while the table is filled is irrelevant (all equal names).
#include "fuz_internal.cuh"

namespace {

struct QidScratch {
    int32_t *slots;      // [cap] record index or -1
    uint32_t mask;       // cap - 1 (cap is a power of two >= 2 * n_rec)
    int32_t *rep;        // [n_rec] first record with the same (contig, name)
    int32_t *first;      // [n_rec + 1] 1 if the record is the first of its name
    int32_t *rank;       // [n_rec + 1] exclusive scan of first
};

// the name of a record that is safe to read: offset inside the buffer, name inside the record
struct RecName {
    const uint8_t *p;
    int len;             // without the trailing NUL; -1: unusable record (reported by k_project)
};

__device__ __forceinline__ RecName rec_name(const uint8_t *__restrict__ rec_buf, const int64_t *__restrict__ rec_off, int r,
                                            int64_t rec_bytes) {
    RecName n;
    n.p = nullptr; n.len = -1;
    const int64_t o = rec_off[r], e = rec_off[r + 1];
    if (o < 0 || e > rec_bytes || e - o < 37) return n;
    const int l_name = rec_buf[o + 12];
    if (l_name < 1 || 36 + (int64_t)l_name > e - o) return n;
    n.p = rec_buf + o + 36;
    n.len = l_name - 1;
    return n;
}

__device__ __forceinline__ uint32_t name_hash(const RecName &n, int c) {
    uint32_t h = 2166136261u ^ (uint32_t)c * 0x9E3779B1u;       // FNV-1a over the name, seeded with the contig
    for (int i = 0; i < n.len; i++) h = (h ^ n.p[i]) * 16777619u;
    h ^= h >> 15;
    return h;
}

__device__ __forceinline__ bool same_name(const RecName &a, const RecName &b) {
    if (a.len != b.len) return false;
    for (int i = 0; i < a.len; i++)
        if (a.p[i] != b.p[i]) return false;
    return true;
}

__global__ void k_qid_insert(const uint8_t *__restrict__ rec_buf, const int64_t *__restrict__ rec_off, int n_rec,
                             int64_t rec_bytes, const int32_t *__restrict__ ctg_rec_off, int n_ctg, QidScratch Q) {
    fuz_pdl_enter();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rec; r += gridDim.x * blockDim.x) {
        const RecName me = rec_name(rec_buf, rec_off, r, rec_bytes);
        if (me.len < 0) continue;
        const int c = fuz_upper_bound(ctg_rec_off, 0, n_ctg + 1, r) - 1;
        if (c < 0 || c >= n_ctg) continue;
        const int c0 = ctg_rec_off[c], c1 = ctg_rec_off[c + 1];
        for (uint32_t s = name_hash(me, c) & Q.mask;; s = (s + 1) & Q.mask) {
            const int cur = atomicCAS(&Q.slots[s], -1, r);
            if (cur == -1) break;                                        // claimed an empty slot
            if (cur >= c0 && cur < c1 && same_name(me, rec_name(rec_buf, rec_off, cur, rec_bytes))) {
                atomicMin(&Q.slots[s], r);
                break;
            }
        }
    }
}

__global__ void k_qid_first(const uint8_t *__restrict__ rec_buf, const int64_t *__restrict__ rec_off, int n_rec,
                            int64_t rec_bytes, const int32_t *__restrict__ ctg_rec_off, int n_ctg, QidScratch Q) {
    fuz_pdl_enter();
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_rec; r += gridDim.x * blockDim.x) {
        if (r == n_rec) { Q.first[r] = 0; continue; }
        int rep = r;                                                     // unusable records stand for themselves
        const RecName me = rec_name(rec_buf, rec_off, r, rec_bytes);
        const int c = fuz_upper_bound(ctg_rec_off, 0, n_ctg + 1, r) - 1;
        if (me.len >= 0 && c >= 0 && c < n_ctg) {
            const int c0 = ctg_rec_off[c], c1 = ctg_rec_off[c + 1];
            for (uint32_t s = name_hash(me, c) & Q.mask;; s = (s + 1) & Q.mask) {
                const int cur = Q.slots[s];
                if (cur == -1) break;                                    // cannot happen after k_qid_insert
                if (cur >= c0 && cur < c1 && same_name(me, rec_name(rec_buf, rec_off, cur, rec_bytes))) { rep = cur; break; }
            }
        }
        Q.rep[r] = rep;
        Q.first[r] = rep == r ? 1 : 0;
    }
}

__global__ void k_qid_assign(int n_rec, const int32_t *__restrict__ ctg_rec_off, int n_ctg, QidScratch Q,
                             int32_t *__restrict__ rec_qid, int32_t *__restrict__ ctg_nq, int64_t *__restrict__ name_first,
                             int32_t *__restrict__ ctg_slots) {
    fuz_pdl_enter();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int r = tid; r < n_rec; r += nt) {
        const int c = fuz_upper_bound(ctg_rec_off, 0, n_ctg + 1, r) - 1;
        const int base = (c >= 0 && c < n_ctg) ? Q.rank[ctg_rec_off[c]] : 0;
        rec_qid[r] = Q.rank[Q.rep[r]] - base;
        if (Q.first[r] && name_first) name_first[Q.rank[r]] = r;
    }
    for (int c = tid; c < n_ctg; c += nt) {
        ctg_nq[c] = Q.rank[ctg_rec_off[c + 1]] - Q.rank[ctg_rec_off[c]];
        if (ctg_slots) ctg_slots[c] = ctg_rec_off[c + 1] - ctg_rec_off[c];   // one read slot per record (see fuz_phase_batch)
    }
}

}  // namespace

// scratch bytes of fuz_assign_qids_impl for n_rec records
size_t fuz_qid_scratch_bytes(int32_t n_rec) {
    uint32_t cap = 64;
    while (cap < 2u * (uint32_t)n_rec) cap <<= 1;
    return 4 * (size_t)cap + 3 * (((size_t)(n_rec + 2) * 4 + 255) & ~(size_t)255) + 256;
}

// scratch = nullptr: the context arena (stand-alone call); otherwise a caller-owned buffer of
// fuz_qid_scratch_bytes(n_rec), which lets the kernels run next to a stage that uses the arena
int fuz_assign_qids_impl(fuz_ctx *ctx, const uint8_t *d_rec_buf, const int64_t *d_rec_off, int32_t n_rec, int64_t rec_bytes,
                         const int32_t *d_ctg_rec_off, int32_t n_ctg, int32_t *d_rec_qid, int32_t *d_ctg_nq,
                         int64_t *d_name_first, int32_t *d_ctg_slots, uint8_t *scratch) {
    cudaStream_t st = ctx->stream;
    uint32_t cap = 64;
    while (cap < 2u * (uint32_t)n_rec) cap <<= 1;
    const size_t sz = ((size_t)(n_rec + 2) * 4 + 255) & ~(size_t)255;
    if (!scratch) {
        FuzLayout L;
        size_t o = L.add(fuz_qid_scratch_bytes(n_rec));
        int rc = fuz_arena_commit(ctx, L);
        if (rc) return rc;
        scratch = fuz_at<uint8_t>(ctx, o);
    }
    QidScratch Q;
    Q.slots = reinterpret_cast<int32_t *>(scratch); Q.mask = cap - 1;
    Q.rep = reinterpret_cast<int32_t *>(scratch + 4 * (size_t)cap);
    Q.first = reinterpret_cast<int32_t *>(scratch + 4 * (size_t)cap + sz);
    Q.rank = reinterpret_cast<int32_t *>(scratch + 4 * (size_t)cap + 2 * sz);
    int rc;
    FUZ_CUDA(ctx, cudaMemsetAsync(Q.slots, 0xFF, 4 * (size_t)cap, st));
    fuz_launch(ctx, k_qid_insert, FUZ_GRID_BLOCKS, 256, 0, st, d_rec_buf, d_rec_off, (int)n_rec, rec_bytes, d_ctg_rec_off, (int)n_ctg, Q);
    FUZ_LAUNCH_CHECK(ctx, "k_qid_insert");
    fuz_launch(ctx, k_qid_first, FUZ_GRID_BLOCKS, 256, 0, st, d_rec_buf, d_rec_off, (int)n_rec, rec_bytes, d_ctg_rec_off, (int)n_ctg, Q);
    FUZ_LAUNCH_CHECK(ctx, "k_qid_first");
    if ((rc = fuz_scan_i32(ctx, Q.first, Q.rank, n_rec, nullptr, FUZ_FIN_NONE, 0))) return rc;
    fuz_launch(ctx, k_qid_assign, FUZ_GRID_BLOCKS, 256, 0, st, (int)n_rec, d_ctg_rec_off, (int)n_ctg, Q, d_rec_qid, d_ctg_nq, d_name_first,
               d_ctg_slots);
    FUZ_LAUNCH_CHECK(ctx, "k_qid_assign");
    return FUZ_OK;
}

extern "C" int fuz_assign_qids(fuz_ctx *ctx, const uint8_t *d_rec_buf, const int64_t *d_rec_off, int32_t n_rec,
                               int64_t rec_bytes, const int32_t *d_ctg_rec_off, int32_t n_ctg, int32_t *d_rec_qid,
                               int32_t *d_ctg_nq, int64_t *d_name_first) {
    if (!ctx || !d_rec_buf || !d_rec_off || !d_ctg_rec_off || !d_rec_qid || !d_ctg_nq || n_rec < 0 || n_ctg < 1 || rec_bytes < 0)
        return fuz_fail(ctx, FUZ_E_ARG, "fuz_assign_qids: bad arguments");
    if (n_rec > 0x3fffffff) return fuz_fail(ctx, FUZ_E_ARG, "fuz_assign_qids: too many records");
    return fuz_assign_qids_impl(ctx, d_rec_buf, d_rec_off, n_rec, rec_bytes, d_ctg_rec_off, n_ctg, d_rec_qid, d_ctg_nq,
                                d_name_first, nullptr, nullptr);
}
