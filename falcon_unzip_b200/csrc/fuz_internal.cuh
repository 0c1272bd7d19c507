// Internal declarations shared by the translation units of libfuz.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <string>
#include <vector>

#include <utility>

#include "fuz.h"

#define FUZ_TILE 8192              // contigs are padded to whole tiles of this size (= the tile of the segment pileup)
#define FUZ_PTILE 2048             // tile of the projection pileup and of the cross-check het test (divides FUZ_TILE)
#define FUZ_PTILE_THREADS 256      // 8 positions (one 32-bit word of 4-bit codes) per thread
#define FUZ_NW (FUZ_PTILE_THREADS / 32)
#define FUZ_GRID_BLOCKS (148 * 4)  // grid-stride kernels: 4 CTAs of 256 threads per SM

struct fuz_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    // scratch arena (device), grown on demand, reused across calls
    uint8_t *arena = nullptr;
    size_t arena_cap = 0;
    // small device buffer that survives from one stage to the next inside fuz_phase_batch
    // (row range of every site, duplicate flags of the variant_map rows)
    uint8_t *keep = nullptr;
    size_t keep_cap = 0;
    // q_ids assigned by the library (fuz_phase_batch with d_rec_qid == NULL): live for the whole call
    uint8_t *qid_buf = nullptr;
    size_t qid_cap = 0;
    // the q_id kernels run on a side stream next to the projection; joined before k_signature
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool join_pending = false;
    // status block
    fuz_status *d_status = nullptr;
    fuz_status *h_status = nullptr;   // pinned
    int64_t launches = 0;
    int pileup_impl = 0;
    int host_fetch = 1;                // host entry: fetch only header/name/CIGAR/SEQ from page-locked records
    int pdl = 1;                       // programmatic dependent launch between the kernels of a call
    int fetch_ctas = FUZ_GRID_BLOCKS * 2;   // CTAs of k_fetch_records (fewer leave SM room for a second context that computes meanwhile)
    int rr_filter_only = 0;            // fuz_rr_track stops after the overlap filter
    int project_ctas = 148 * 5;        // CTAs of k_project (persistent warps, records from a global cursor)
    int phase_staging = 0;             // 0 auto, 1 at most the sweep tier, 2 global memory only (tests)
    int sweep_passes = 64;             // parallel fixed-point passes of the pass-2 sweep before the sequential walk (0 = none)
    int64_t max_pairs_per_site = 96;
    uint8_t *reads_buf = nullptr; size_t reads_cap = 0;           // scratch of the read stage (runs beside the block stage)
    uint8_t *scan_state = nullptr; size_t scan_state_cap = 0;   // tile states of the multi-CTA scan
    bool ingest_pending = false;       // fuz_bgzf_inflate ran; the next fuz_bam_index_records keeps its status
    bool phase_attr_set = false;
    bool pileup_attr_set = false;
    bool gather_attr_set = false;
    int grid_rr = 4;                   // tracking path: 0 = every kernel on 4 CTAs per SM; n > 0 = k_rr_replay on n CTAs per SM (more chains in flight
                                       // thrash the caches: 6 is 25 % slower than 4), k_rr_fill on 8, k_rr_vote on 16 (latency bound)
    int grid_sig = 8, grid_assoc = 6, grid_reads = 8;   // CTAs per SM of the latency-bound grid-stride kernels (k_signature / association / read stage): measured against 4
    int gather_tma = 1;                // pileup_impl 0: TMA-fed persistent gather (1) or the plain tile-per-CTA kernel (0)
    int pileup_debug = 0;
    uint32_t *trace = nullptr;         // host-mapped progress markers (debugging)
    int64_t seg_cap_min = 0, ent_cap_min = 0;   // reservations of the segment pileup beyond the heuristics (capacity retry)
    // per-launch profile (diagnostics): an event after every kernel launch
    bool profile = false;
    cudaEvent_t prof_start = nullptr;
    std::vector<std::pair<const char *, cudaEvent_t>> prof_marks;
    size_t prof_used = 0;
    // timing of the dominant kernel
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    size_t timing_used = 0;
    // device + pinned staging for the *_host entry points
    uint8_t *stage_dev = nullptr; size_t stage_dev_cap = 0;
    uint8_t *stage_pin = nullptr; size_t stage_pin_cap = 0;
};

int fuz_fail(fuz_ctx *ctx, int code, const char *fmt, ...);

#define FUZ_CUDA(ctx, expr)                                                              \
    do {                                                                                 \
        cudaError_t e_ = (expr);                                                         \
        if (e_ != cudaSuccess)                                                           \
            return fuz_fail((ctx), FUZ_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define FUZ_LAUNCH_CHECK(ctx, name)                                                      \
    do {                                                                                 \
        (ctx)->launches++;                                                               \
        cudaError_t e_ = cudaGetLastError();                                             \
        if (e_ != cudaSuccess)                                                           \
            return fuz_fail((ctx), FUZ_E_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
        if ((ctx)->profile) fuz_profile_mark((ctx), name);                               \
    } while (0)

// Programmatic dependent launch: every kernel of the library starts with fuz_pdl_enter()
// (let the next kernel of the stream start launching, then wait until everything before this
// kernel has completed and is visible), and is launched through fuz_launch(), which allows the
// overlap of its launch with the tail of its predecessor.  Profile mode launches serialised.
#ifdef __CUDACC__
__device__ __forceinline__ void fuz_pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline void fuz_launch(fuz_ctx *ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (ctx->profile || !ctx->pdl) ? 0 : 1;
    (void)cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);   // errors: FUZ_LAUNCH_CHECK
}
#endif

void fuz_profile_mark(fuz_ctx *ctx, const char *name);

// Bump layout over the context arena: add() all buffers, then commit() (grows the arena
// if needed; growing synchronises the stream) and resolve pointers with at().
struct FuzLayout {
    size_t off = 0;
    size_t add(size_t bytes) {
        size_t o = off;
        off += (bytes + 255) & ~(size_t)255;
        return o;
    }
};
int fuz_arena_commit(fuz_ctx *ctx, const FuzLayout &l);
// inter-stage buffer: row_off [cap_sites + 2] int32, at_off [cap_sites + 2] int32, dup [cap_vmap + 1] uint8
int fuz_keep_commit(fuz_ctx *ctx, int64_t cap_sites, int64_t cap_vmap, int32_t **row_off, uint8_t **dup,
                    int32_t **at_off = nullptr);
template <typename T>
static inline T *fuz_at(fuz_ctx *ctx, size_t off) { return reinterpret_cast<T *>(ctx->arena + off); }

#ifdef __CUDACC__
// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ void fuz_raise(fuz_status *st, int code, int idx) {
    if (atomicCAS(&st->error, 0, code) == 0) st->error_index = idx;
}

// 32-bit little-endian load from an arbitrarily aligned address.  Always touches the
// aligned word after it: callers guarantee >= 8 readable bytes past any loaded field
// (rec_buf carries 16 bytes of slack).
__device__ __forceinline__ uint32_t fuz_ld_u32_un(const uint8_t *p) {
    uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(a & ~(uintptr_t)3);
    uint32_t sh = (uint32_t)(a & 3) * 8;
    return __funnelshift_r(__ldg(w), __ldg(w + 1), sh);
}

__device__ __forceinline__ int fuz_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ int fuz_warp_sum(int v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ long long fuz_warp_sum64(long long v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
// first index in [lo, hi) with a[i] >= key
__device__ __forceinline__ int fuz_lower_bound(const int32_t *a, int lo, int hi, int key) {
    while (lo < hi) {
        int m = (lo + hi) >> 1;
        if (a[m] < key) lo = m + 1; else hi = m;
    }
    return lo;
}
// first index in [lo, hi) with a[i] > key
__device__ __forceinline__ int fuz_upper_bound(const int32_t *a, int lo, int hi, int key) {
    while (lo < hi) {
        int m = (lo + hi) >> 1;
        if (a[m] <= key) lo = m + 1; else hi = m;
    }
    return lo;
}
// Exclusive scan of n int32 values by ONE CTA of 1024 threads (every thread of the CTA must
// call it); out[n] = total, which is also returned (in every thread).  A pass = 4096 elements,
// 4 consecutive ones per thread (one coalesced 128-bit load); the loads of 4 passes are issued
// together so that arrays up to 16 K entries cost one memory round trip and one barrier per pass.
__device__ __forceinline__ long long fuz_cta_scan_i32(const int32_t *__restrict__ in, int32_t *__restrict__ out, int64_t n) {
    __shared__ int fuz_scan_tot[2][32];            // double buffered: one barrier per pass
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    long long carry_s = 0;                         // replicated in every thread
    int buf = 0;
    const bool vec_in = (reinterpret_cast<uintptr_t>(in) & 15) == 0, vec_out = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    for (int64_t base = 0; base < n; base += 4096 * 4) {
        int4 v[4];
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const int64_t i0 = base + p * 4096 + (int64_t)tid * 4;
            v[p] = make_int4(0, 0, 0, 0);
            if (i0 < n) {
                if (vec_in && i0 + 4 <= n) v[p] = *reinterpret_cast<const int4 *>(in + i0);
                else {
                    v[p].x = in[i0];
                    if (i0 + 1 < n) v[p].y = in[i0 + 1];
                    if (i0 + 2 < n) v[p].z = in[i0 + 2];
                    if (i0 + 3 < n) v[p].w = in[i0 + 3];
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 4; p++, buf ^= 1) {
            const int64_t i0 = base + p * 4096 + (int64_t)tid * 4;
            if (base + p * 4096 >= n) break;       // uniform
            const int s = v[p].x + v[p].y + v[p].z + v[p].w;
            const int incl = fuz_warp_incl_scan(s, lane);
            if (lane == 31) fuz_scan_tot[buf][warp] = incl;
            __syncthreads();
            const int t = fuz_scan_tot[buf][lane];
            const int ti = fuz_warp_incl_scan(t, lane);
            const int wexcl = __shfl_sync(0xffffffffu, ti - t, warp);
            const int chunk_total = __shfl_sync(0xffffffffu, ti, 31);
            const long long excl = carry_s + wexcl + (incl - s);
            if (i0 < n) {
                int4 o;
                o.x = (int)excl; o.y = o.x + v[p].x; o.z = o.y + v[p].y; o.w = o.z + v[p].z;
                if (vec_out && i0 + 4 <= n) *reinterpret_cast<int4 *>(out + i0) = o;
                else {
                    out[i0] = o.x;
                    if (i0 + 1 < n) out[i0 + 1] = o.y;
                    if (i0 + 2 < n) out[i0 + 2] = o.z;
                    if (i0 + 3 < n) out[i0 + 3] = o.w;
                }
            }
            carry_s += chunk_total;
        }
    }
    if (tid == 0) out[n] = (int32_t)carry_s;
    return carry_s;
}

// publish the total of a scan into the status block as a row count with a capacity check
__device__ __forceinline__ void fuz_scan_publish(fuz_status *st, int fin_op, int64_t fin_cap, long long total) {
    if (!st || st->error) return;
    switch (fin_op) {
    case 1: st->need_sites = total; if (total > fin_cap) fuz_raise(st, FUZ_E_CAPACITY, 0); else st->n_sites = total; break;
    case 2: st->need_vmap = total; if (total > fin_cap) fuz_raise(st, FUZ_E_CAPACITY, 1); else st->n_vmap = total; break;
    case 3: st->need_atable = total; if (total > fin_cap) fuz_raise(st, FUZ_E_CAPACITY, 2); else st->n_atable = total; break;
    case 4: st->need_reads = total; if (total > fin_cap) fuz_raise(st, FUZ_E_CAPACITY, 3); else st->n_reads = total; break;
    case 5: st->need_pairs = total; if (total > fin_cap) fuz_raise(st, FUZ_E_CAPACITY, 4); break;
    default: break;
    }
}
#endif

// exclusive scan of n int32 values (n read from *d_n if d_n != nullptr, clamped to cap);
// out has n+1 entries (out[n] = total).  fin_op publishes the total into the status block
// as a row count with a capacity check (FUZ_FIN_NONE: nothing).
enum { FUZ_FIN_NONE = 0, FUZ_FIN_SITES = 1, FUZ_FIN_VMAP = 2, FUZ_FIN_ATABLE = 3, FUZ_FIN_READS = 4, FUZ_FIN_PAIRS = 5 };
int fuz_scan_i32(fuz_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t n_cap,
                 const int64_t *d_n, int fin_op, int64_t fin_cap);
// the same for a host-known n of millions of entries: multi-CTA single-pass scan (decoupled
// look-back); main stream only
int fuz_scan_i32_wide(fuz_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t n, int fin_op = FUZ_FIN_NONE, int64_t fin_cap = 0,
                      const int64_t *d_n = nullptr);   // d_n: device-side length <= n (n = capacity)
