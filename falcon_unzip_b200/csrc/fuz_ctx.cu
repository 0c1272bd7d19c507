// Context, scratch arena, status block, scan primitive.
#include <stdlib.h>
#include <string.h>

#include "fuz_internal.cuh"

static std::string g_create_err;

int fuz_fail(fuz_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_err = buf;
    return code;
}

extern "C" int fuz_version(void) { return FUZ_VERSION; }
extern "C" int fuz_tile_size(void) { return FUZ_TILE; }

extern "C" const char *fuz_last_error(fuz_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_create_err.c_str();
}

extern "C" int fuz_ctx_create(int device, fuz_ctx **out) {
    if (!out) return fuz_fail(nullptr, FUZ_E_ARG, "fuz_ctx_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fuz_fail(nullptr, FUZ_E_CUDA, "no CUDA device available (%s); libfuz has no CPU fallback",
                        e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fuz_fail(nullptr, FUZ_E_ARG, "device %d out of range [0,%d)", device, n);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return fuz_fail(nullptr, FUZ_E_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fuz_fail(nullptr, FUZ_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fuz_fail(nullptr, FUZ_E_CUDA, "device %d is sm_%d%d; libfuz is built for sm_100a (B200) only",
                        device, prop.major, prop.minor);
    fuz_ctx *ctx = new fuz_ctx();
    ctx->device = device;
    if (const char *v = getenv("FUZ_PDL")) ctx->pdl = atoi(v) != 0;      // experiments: FUZ_PDL=0 -> plain launches
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        delete ctx;
        return fuz_fail(nullptr, FUZ_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->own_stream = true;
    if ((e = cudaMalloc(&ctx->d_status, sizeof(fuz_status))) != cudaSuccess ||
        (e = cudaMallocHost(&ctx->h_status, sizeof(fuz_status))) != cudaSuccess) {
        fuz_ctx_destroy(ctx);
        return fuz_fail(nullptr, FUZ_E_CUDA, "status allocation: %s", cudaGetErrorString(e));
    }
    cudaMemset(ctx->d_status, 0, sizeof(fuz_status));
    *out = ctx;
    return FUZ_OK;
}

extern "C" int fuz_ctx_destroy(fuz_ctx *ctx) {
    if (!ctx) return FUZ_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &p : ctx->timing_events) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto &p : ctx->prof_marks) cudaEventDestroy(p.second);
    if (ctx->prof_start) cudaEventDestroy(ctx->prof_start);
    if (ctx->arena) cudaFree(ctx->arena);
    if (ctx->keep) cudaFree(ctx->keep);
    if (ctx->qid_buf) cudaFree(ctx->qid_buf);
    if (ctx->scan_state) cudaFree(ctx->scan_state);
    if (ctx->reads_buf) cudaFree(ctx->reads_buf);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->stage_dev) cudaFree(ctx->stage_dev);
    if (ctx->stage_pin) cudaFreeHost(ctx->stage_pin);
    if (ctx->d_status) cudaFree(ctx->d_status);
    if (ctx->h_status) cudaFreeHost(ctx->h_status);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return FUZ_OK;
}

extern "C" int fuz_set_stream(fuz_ctx *ctx, void *cuda_stream) {
    if (!ctx) return FUZ_E_ARG;
    if (ctx->own_stream && ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return FUZ_OK;
}

extern "C" int fuz_set_option(fuz_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return FUZ_E_ARG;
    if (!strcmp(key, "pileup_impl")) {
        if (value < 0 || value > 3) return fuz_fail(ctx, FUZ_E_ARG, "pileup_impl must be 0, 1, 2 or 3");
        ctx->pileup_impl = (int)value;
    } else if (!strcmp(key, "host_fetch")) {
        if (value != 0 && value != 1) return fuz_fail(ctx, FUZ_E_ARG, "host_fetch must be 0 or 1");
        ctx->host_fetch = (int)value;
    } else if (!strcmp(key, "pdl")) {
        if (value != 0 && value != 1) return fuz_fail(ctx, FUZ_E_ARG, "pdl must be 0 or 1");
        ctx->pdl = (int)value;
    } else if (!strcmp(key, "rr_filter_only")) {
        if (value != 0 && value != 1) return fuz_fail(ctx, FUZ_E_ARG, "rr_filter_only must be 0 or 1");
        ctx->rr_filter_only = (int)value;
    } else if (!strcmp(key, "phase_staging")) {
        if (value < 0 || value > 2) return fuz_fail(ctx, FUZ_E_ARG, "phase_staging must be 0, 1 or 2");
        ctx->phase_staging = (int)value;
    } else if (!strcmp(key, "grid_sig") || !strcmp(key, "grid_assoc") || !strcmp(key, "grid_reads")) {
        if (value < 1 || value > 8) return fuz_fail(ctx, FUZ_E_ARG, "%s must be in 1 .. 8 (CTAs per SM)", key);
        (key[5] == 's' ? ctx->grid_sig : key[5] == 'a' ? ctx->grid_assoc : ctx->grid_reads) = (int)value;
    } else if (!strcmp(key, "grid_rr")) {
        if (value < 0 || value > 8) return fuz_fail(ctx, FUZ_E_ARG, "grid_rr must be in 0 .. 8");
        ctx->grid_rr = (int)value;
    } else if (!strcmp(key, "gather_tma")) {
        if (value != 0 && value != 1) return fuz_fail(ctx, FUZ_E_ARG, "gather_tma must be 0 or 1");
        ctx->gather_tma = (int)value;
    } else if (!strcmp(key, "fetch_ctas")) {
        if (value < 1 || value > 148 * 16) return fuz_fail(ctx, FUZ_E_ARG, "fetch_ctas must be in 1 .. 2368");
        ctx->fetch_ctas = (int)value;
    } else if (!strcmp(key, "sweep_passes")) {
        if (value < 0 || value > 1 << 20) return fuz_fail(ctx, FUZ_E_ARG, "sweep_passes must be in 0 .. 2^20");
        ctx->sweep_passes = (int)value;
    } else if (!strcmp(key, "project_ctas")) {
        if (value < 1 || value > 148 * 6) return fuz_fail(ctx, FUZ_E_ARG, "project_ctas must be in 1 .. 888");
        ctx->project_ctas = (int)value;
    } else if (!strcmp(key, "trace_ptr")) {
        ctx->trace = reinterpret_cast<uint32_t *>(static_cast<uintptr_t>(value));
    } else if (!strcmp(key, "pileup_debug")) {
        ctx->pileup_debug = (int)value;
    } else if (!strcmp(key, "seg_cap")) {
        if (value < 0) return fuz_fail(ctx, FUZ_E_ARG, "seg_cap must be >= 0");
        ctx->seg_cap_min = value;
    } else if (!strcmp(key, "ent_cap")) {
        if (value < 0) return fuz_fail(ctx, FUZ_E_ARG, "ent_cap must be >= 0");
        ctx->ent_cap_min = value;
    } else if (!strcmp(key, "max_pairs_per_site")) {
        if (value < 1) return fuz_fail(ctx, FUZ_E_ARG, "max_pairs_per_site must be >= 1");
        ctx->max_pairs_per_site = value;
    } else {
        return fuz_fail(ctx, FUZ_E_ARG, "unknown option '%s'", key);
    }
    return FUZ_OK;
}

extern "C" int fuz_sync(fuz_ctx *ctx) {
    if (!ctx) return FUZ_E_ARG;
    FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FUZ_OK;
}

extern "C" int fuz_get_status(fuz_ctx *ctx, fuz_status *h_status) {
    if (!ctx || !h_status) return FUZ_E_ARG;
    FUZ_CUDA(ctx, cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(fuz_status), cudaMemcpyDeviceToHost, ctx->stream));
    FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *h_status = *ctx->h_status;
    if (h_status->error != FUZ_OK) {
        static const char *names[] = {"ok", "cuda", "arg", "capacity", "bad record", "unsorted records",
                                      "pileup depth > 65535", "internal inconsistency", "format"};
        int e = h_status->error;
        return fuz_fail(ctx, e, "device reported error %d (%s) at index %d", e,
                        (e >= 0 && e <= 8) ? names[e] : "?", h_status->error_index);
    }
    return FUZ_OK;
}

extern "C" int64_t fuz_launch_count(fuz_ctx *ctx) { return ctx ? ctx->launches : -1; }

extern "C" int fuz_kernel_timing(fuz_ctx *ctx, int enable) {
    if (!ctx) return FUZ_E_ARG;
    ctx->timing = enable != 0;
    ctx->timing_used = 0;
    return FUZ_OK;
}

extern "C" int fuz_get_kernel_timing(fuz_ctx *ctx, double *h_ms_total, int64_t *h_launches) {
    if (!ctx || !h_ms_total || !h_launches) return FUZ_E_ARG;
    FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double total = 0;
    for (size_t i = 0; i < ctx->timing_used; i++) {
        float ms = 0;
        FUZ_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->timing_events[i].first, ctx->timing_events[i].second));
        total += ms;
    }
    *h_ms_total = total;
    *h_launches = (int64_t)ctx->timing_used;
    return FUZ_OK;
}

int fuz_keep_commit(fuz_ctx *ctx, int64_t cap_sites, int64_t cap_vmap, int32_t **row_off, uint8_t **dup, int32_t **at_off) {
    const size_t sz_off = ((size_t)(cap_sites + 2) * 4 + 255) & ~(size_t)255;
    const size_t need = 2 * sz_off + (size_t)cap_vmap + 256;
    if (need > ctx->keep_cap) {
        FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->keep) FUZ_CUDA(ctx, cudaFree(ctx->keep));
        ctx->keep = nullptr; ctx->keep_cap = 0;
        cudaError_t e = cudaMalloc(&ctx->keep, need + (need >> 2));
        if (e != cudaSuccess) return fuz_fail(ctx, FUZ_E_CUDA, "inter-stage buffer of %zu bytes: %s", need, cudaGetErrorString(e));
        ctx->keep_cap = need + (need >> 2);
    }
    *row_off = reinterpret_cast<int32_t *>(ctx->keep);
    if (at_off) *at_off = reinterpret_cast<int32_t *>(ctx->keep + sz_off);
    *dup = ctx->keep + 2 * sz_off;
    return FUZ_OK;
}

// ---- per-launch profile: fuz_profile(ctx, 1) starts a capture, fuz_profile_report prints
// "name ms" for every launch since (device time between consecutive marks on the stream).
void fuz_profile_mark(fuz_ctx *ctx, const char *name) {
    if (ctx->prof_used == ctx->prof_marks.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        ctx->prof_marks.push_back({name, e});
    }
    ctx->prof_marks[ctx->prof_used].first = name;
    cudaEventRecord(ctx->prof_marks[ctx->prof_used].second, ctx->stream);
    ctx->prof_used++;
}

extern "C" int fuz_profile(fuz_ctx *ctx, int enable) {
    if (!ctx) return FUZ_E_ARG;
    ctx->profile = enable != 0;
    ctx->prof_used = 0;
    if (enable) {
        if (!ctx->prof_start) FUZ_CUDA(ctx, cudaEventCreate(&ctx->prof_start));
        FUZ_CUDA(ctx, cudaEventRecord(ctx->prof_start, ctx->stream));
    }
    return FUZ_OK;
}

// Writes "name\tms\n" lines into buf (truncated to cap); returns the number of marks.
extern "C" int64_t fuz_profile_report(fuz_ctx *ctx, char *buf, int64_t cap) {
    if (!ctx || !buf || cap < 1) return -1;
    cudaStreamSynchronize(ctx->stream);
    int64_t w = 0;
    buf[0] = 0;
    cudaEvent_t prev = ctx->prof_start;
    for (size_t i = 0; i < ctx->prof_used; i++) {
        float ms = 0;
        cudaEventElapsedTime(&ms, prev, ctx->prof_marks[i].second);
        prev = ctx->prof_marks[i].second;
        int n = snprintf(buf + w, (size_t)(cap - w), "%s\t%.4f\n", ctx->prof_marks[i].first, ms);
        if (n < 0 || w + n >= cap) break;
        w += n;
    }
    return (int64_t)ctx->prof_used;
}

int fuz_arena_commit(fuz_ctx *ctx, const FuzLayout &l) {
    if (l.off <= ctx->arena_cap) return FUZ_OK;
    FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->arena) FUZ_CUDA(ctx, cudaFree(ctx->arena));
    ctx->arena = nullptr;
    ctx->arena_cap = 0;
    size_t want = l.off + (l.off >> 2) + (1 << 20);
    cudaError_t e = cudaMalloc(&ctx->arena, want);
    if (e != cudaSuccess) return fuz_fail(ctx, FUZ_E_CUDA, "scratch arena of %zu bytes: %s", want, cudaGetErrorString(e));
    ctx->arena_cap = want;
    return FUZ_OK;
}

// ------------------------------------------------------------------ single-CTA scan
// One CTA of 1024 threads walks the array in chunks of 16384 carrying the running total.
// The arrays scanned on this path (records, tiles, sites, q_ids) are at most a few
// million entries; the scan is never the dominant kernel.  The tail of the scan also
// publishes the total into the status block (row counts + capacity checks), which saves
// one tiny kernel launch per scan.
__global__ void __launch_bounds__(1024) k_scan_i32(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                   int64_t n_cap, const int64_t *__restrict__ d_n, int fin_op,
                                                   int64_t fin_cap, fuz_status *st) {
    fuz_pdl_enter();
    int64_t n = d_n ? *d_n : n_cap;
    if (n > n_cap) n = n_cap;
    if (n < 0) n = 0;
    if (st && st->error) n = 0;
    const long long total = fuz_cta_scan_i32(in, out, n);
    if (threadIdx.x == 0) fuz_scan_publish(st, fin_op, fin_cap, total);
}

// ------------------------------------------------------------------ multi-CTA scan
// Arrays of millions of entries with a host-known length (overlap lines): single pass with
// decoupled look-back.  Tiles of 4096 entries are handed out by an atomic counter (a tile only
// waits for tiles that already run); a tile publishes its aggregate, then its inclusive prefix,
// flag and value packed in one 64-bit word; warp 0 looks back 32 tiles at a time.
#define FUZ_SCAN_TILE 4096
__global__ void __launch_bounds__(1024) k_scan_wide(const int32_t *__restrict__ in, int32_t *__restrict__ out, int64_t n_cap,
                                                    const int64_t *__restrict__ d_n, unsigned long long *state, unsigned int *counter,
                                                    int fin_op, int64_t fin_cap, fuz_status *st) {
    fuz_pdl_enter();
    int64_t n = d_n ? *d_n : n_cap;                // a device-side length: the grid covers n_cap, tiles past n carry zeros
    if (n > n_cap) n = n_cap;
    if (n < 0) n = 0;
    if (st && st->error) n = 0;
    __shared__ int s_tile;
    __shared__ int s_warp[32];
    __shared__ int s_prefix, s_agg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (int)atomicAdd(counter, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int64_t i0 = (int64_t)tile * FUZ_SCAN_TILE + (int64_t)tid * 4;
    int4 v = make_int4(0, 0, 0, 0);
    const bool vec = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (i0 + 4 <= n && vec) v = *reinterpret_cast<const int4 *>(in + i0);
    else if (i0 < n) {
        v.x = in[i0];
        if (i0 + 1 < n) v.y = in[i0 + 1];
        if (i0 + 2 < n) v.z = in[i0 + 2];
        if (i0 + 3 < n) v.w = in[i0 + 3];
    }
    const int s = v.x + v.y + v.z + v.w;
    const int incl = fuz_warp_incl_scan(s, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int t = s_warp[lane];
        const int ti = fuz_warp_incl_scan(t, lane);
        s_warp[lane] = ti - t;
        const int agg = __shfl_sync(0xffffffffu, ti, 31);
        int prefix = 0;
        if (tile > 0) {
            volatile unsigned long long *vs = state;
            if (lane == 0) vs[tile] = (1ull << 32) | (unsigned int)agg;
            for (int base = tile - 1;; base -= 32) {
                const int idx = base - lane;
                unsigned long long w;
                do {
                    w = idx >= 0 ? vs[idx] : (2ull << 32);
                } while (__any_sync(0xffffffffu, (w >> 32) == 0));
                const uint32_t done = __ballot_sync(0xffffffffu, (w >> 32) == 2);
                const int first = done ? __ffs(done) - 1 : 31;
                prefix += fuz_warp_sum(lane <= first ? (int)(unsigned int)w : 0);
                if (done) break;
            }
        }
        if (lane == 0) {
            reinterpret_cast<volatile unsigned long long *>(state)[tile] = (2ull << 32) | (unsigned int)(prefix + agg);
            s_prefix = prefix;
            s_agg = agg;
        }
    }
    __syncthreads();
    const int excl = s_prefix + s_warp[warp] + (incl - s);
    if (i0 + 4 <= n && vec) {
        *reinterpret_cast<int4 *>(out + i0) = make_int4(excl, excl + v.x, excl + v.x + v.y, excl + v.x + v.y + v.z);
    } else if (i0 < n) {
        out[i0] = excl;
        if (i0 + 1 < n) out[i0 + 1] = excl + v.x;
        if (i0 + 2 < n) out[i0 + 2] = excl + v.x + v.y;
        if (i0 + 3 < n) out[i0 + 3] = excl + v.x + v.y + v.z;
    }
    if (tid == 0 && (int64_t)(tile + 1) * FUZ_SCAN_TILE >= n && ((int64_t)tile * FUZ_SCAN_TILE < n || tile == 0)) {
        out[n] = s_prefix + s_agg;
        fuz_scan_publish(st, fin_op, fin_cap, (long long)(s_prefix + s_agg));
    }
}

int fuz_scan_i32(fuz_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t n_cap, const int64_t *d_n, int fin_op,
                 int64_t fin_cap) {
    fuz_launch(ctx, k_scan_i32, 1, 1024, 0, ctx->stream, d_in, d_out, n_cap, d_n, fin_op, fin_cap, ctx->d_status);
    FUZ_LAUNCH_CHECK(ctx, "k_scan_i32");
    return FUZ_OK;
}

// exclusive scan of n entries (host-known n), out[n] = total; uses the context's tile-state buffer:
// main stream only, one scan at a time
int fuz_scan_i32_wide(fuz_ctx *ctx, const int32_t *d_in, int32_t *d_out, int64_t n, int fin_op, int64_t fin_cap, const int64_t *d_n) {
    if (n <= 16384) return fuz_scan_i32(ctx, d_in, d_out, n, d_n, fin_op, fin_cap);
    const int64_t tiles = (n + FUZ_SCAN_TILE - 1) / FUZ_SCAN_TILE;
    const size_t need = 8 * (size_t)tiles + 64;
    if (need > ctx->scan_state_cap) {
        FUZ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->scan_state) FUZ_CUDA(ctx, cudaFree(ctx->scan_state));
        ctx->scan_state = nullptr; ctx->scan_state_cap = 0;
        FUZ_CUDA(ctx, cudaMalloc(&ctx->scan_state, need * 2));
        ctx->scan_state_cap = need * 2;
    }
    FUZ_CUDA(ctx, cudaMemsetAsync(ctx->scan_state, 0, need, ctx->stream));
    unsigned long long *state = reinterpret_cast<unsigned long long *>(ctx->scan_state) + 1;
    fuz_launch(ctx, k_scan_wide, (unsigned)tiles, 1024, 0, ctx->stream, d_in, d_out, n, d_n, state,
               reinterpret_cast<unsigned int *>(ctx->scan_state), fin_op, fin_cap, (fin_op || d_n) ? ctx->d_status : (fuz_status *)nullptr);
    FUZ_LAUNCH_CHECK(ctx, "k_scan_wide");
    return FUZ_OK;
}
