// Micro-benchmark: latency (cycles per dependent step) of warp-wide primitives used on the
// sequential sweep of k_ctg_phase.  nvcc -arch=sm_100a -O3 lat.cu -o lat && ./lat
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *out, int iters) {
    __shared__ volatile int sm[1024];
    int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + 3) & 1023;
    __syncthreads();
    if (threadIdx.x >= 32) return;
    int v = lane;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) v = __reduce_add_sync(0xffffffffu, v & 3) + lane;
    long long t1 = clock64();
    int w = lane;
    for (int i = 0; i < iters; i++) {
        int s = w & 3;
        for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
        w = s + lane;
    }
    long long t2 = clock64();
    int p = lane;
    for (int i = 0; i < iters; i++) p = sm[p];
    long long t3 = clock64();
    unsigned b = lane;
    for (int i = 0; i < iters; i++) b = __ballot_sync(0xffffffffu, (b >> lane) & 1) + lane;
    long long t4 = clock64();
    int a = lane;
    for (int i = 0; i < iters; i++) a = (a * 3 + 1) ^ (a >> 3);
    long long t5 = clock64();
    int r = lane;
    for (int i = 0; i < iters; i++) r = __shfl_sync(0xffffffffu, r + 1, (r + i) & 31);
    long long t6 = clock64();
    if (lane == 0) {
        out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; out[5] = t6 - t5;
        out[6] = v + w + p + b + a + r;
    }
}
int main() {
    long long *d, h[8];
    cudaMalloc(&d, 64);
    int iters = 10000;
    k<<<1, 64>>>(d, iters);
    k<<<1, 64>>>(d, iters);
    cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
    printf("cycles/step: redux %.1f  shfl5sum %.1f  lds-chain %.1f  ballot %.1f  alu3 %.1f  shfl %.1f\n", (double)h[0] / iters,
           (double)h[1] / iters, (double)h[2] / iters, (double)h[3] / iters, (double)h[4] / iters, (double)h[5] / iters);
    // globaltimer vs clock64 rate
    return 0;
}
