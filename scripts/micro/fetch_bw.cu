// Zero-copy read bandwidth of page-locked host memory from SM loads, in the access pattern of
// k_fetch_records: per record read `need` bytes, skip the rest.  Variants: loads in flight per
// lane, cache hints, grid size.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a fetch_bw.cu -o fetch_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int UNROLL, int HINT>
__global__ void __launch_bounds__(256) k_fetch(const uint8_t *__restrict__ src, uint8_t *__restrict__ dst, int n_rec,
                                               int64_t stride, int64_t need) {
    const int lane = threadIdx.x & 31;
    const int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp_g; r < n_rec; r += n_warps) {
        const int64_t a0 = r * stride, end = a0 + need;
        for (int64_t o = a0 + 16 * lane; o < end; o += UNROLL * 512) {
            uint4 w[UNROLL];
#pragma unroll
            for (int j = 0; j < UNROLL; j++)
                if (o + 512 * j < end) {
                    const uint4 *p = reinterpret_cast<const uint4 *>(src + o + 512 * j);
                    if (HINT == 0) w[j] = __ldcs(p);
                    else if (HINT == 1) w[j] = *p;
                    else asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                                      : "=r"(w[j].x), "=r"(w[j].y), "=r"(w[j].z), "=r"(w[j].w) : "l"(p));
                }
#pragma unroll
            for (int j = 0; j < UNROLL; j++)
                if (o + 512 * j < end) *reinterpret_cast<uint4 *>(dst + o + 512 * j) = w[j];
        }
    }
}

template <int U, int H>
void run(const char *name, const uint8_t *h, uint8_t *d, int n_rec, int64_t stride, int64_t need, int blocks) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_fetch<U, H><<<blocks, 256>>>(h, d, n_rec, stride, need);
    cudaDeviceSynchronize();
    float best = 1e9f;
    for (int it = 0; it < 5; it++) {
        cudaEventRecord(a);
        k_fetch<U, H><<<blocks, 256>>>(h, d, n_rec, stride, need);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    printf("%-28s blocks %5d: %.3f ms  %.1f GB/s (err %s)\n", name, blocks, best, n_rec * (double)need / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int n_rec = 20000;
    const int64_t stride = 15824, need = 5888;   // multiples of 16
    uint8_t *h, *d;
    cudaHostAlloc(&h, n_rec * stride, cudaHostAllocDefault);
    cudaMalloc(&d, n_rec * stride);
    for (int64_t i = 0; i < n_rec * stride; i += 4096) h[i] = (uint8_t)i;
    {   // DMA reference: whole buffer
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaMemcpy(d, h, n_rec * stride, cudaMemcpyHostToDevice);
        cudaEventRecord(a); cudaMemcpyAsync(d, h, n_rec * stride, cudaMemcpyHostToDevice); cudaEventRecord(b);
        cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b);
        printf("cudaMemcpy whole buffer: %.3f ms %.1f GB/s\n", ms, n_rec * (double)stride / ms / 1e6);
    }
    for (int blocks : {148, 296, 592, 1184, 2368}) {
        run<4, 0>("ldcs x4", h, d, n_rec, stride, need, blocks);
        run<8, 0>("ldcs x8", h, d, n_rec, stride, need, blocks);
        run<4, 1>("ld x4", h, d, n_rec, stride, need, blocks);
        run<4, 2>("ld.nc.L2::256B x4", h, d, n_rec, stride, need, blocks);
        run<2, 0>("ldcs x2", h, d, n_rec, stride, need, blocks);
    }
    return 0;
}
