// Micro-benchmark: how long after an mbarrier phase completes does each of W waiting warps wake up?  (B200)
// One CTA, 16 warps.  Warp 15 lane 0 completes the phase (local arrive); warps 0..W-1 sleep in
// mbarrier.try_wait.parity (the same loop the v2 kernel uses) and stamp clock64 when they are released.
// Prints, per number of waiters, the release latency of the first / median / last warp and of warp 0.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_wake mbar_wake.cu
#include <cstdio>
#include <cstdint>
#include <algorithm>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void wake_kernel(int W, int iters, int mode, long long* out) {
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ long long t0s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar))); }
    __syncthreads();
    for (int i = 0; i < iters; ++i) {
        const uint32_t parity = i & 1;
        if (warp == 15) {
            if (lane == 0) {
                long long t = clock64();
                while (clock64() - t < 3000) {}
                t0s = clock64();
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            }
        } else if (warp < W) {
            uint32_t done = 0;
            if (mode == 0) {
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
            } else {
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
            }
            const long long t1 = clock64();
            if (lane == 0) out[i * 16 + warp] = t1;
        }
        __syncthreads();
        if (tid == 0) out[i * 16 + 15] = t0s;
        __syncthreads();
    }
}
int main() {
    const int iters = 200;
    long long* d; cudaMalloc(&d, iters * 16 * sizeof(long long));
    static long long h[200 * 16];
    for (int mode = 0; mode < 2; ++mode)
    for (int W : {1, 2, 4, 5, 8, 12, 15}) {
        cudaMemset(d, 0, iters * 16 * sizeof(long long));
        wake_kernel<<<1, 512>>>(W, iters, mode, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double first = 0, last = 0, w0 = 0, med = 0; int n = 0;
        for (int i = 20; i < iters; ++i) {
            long long v[16]; for (int w = 0; w < W; ++w) v[w] = h[i * 16 + w] - h[i * 16 + 15];
            w0 += v[0]; std::sort(v, v + W); first += v[0]; last += v[W - 1]; med += v[W / 2]; ++n;
        }
        printf("%s waiters %2d: release latency (cycles) first %.0f  median %.0f  last %.0f  warp0 %.0f\n",
               mode ? "test_wait spin" : "try_wait      ", W, first / n, med / n, last / n, w0 / n);
    }
    return 0;
}
