// Micro-benchmark: cost of fanning a value out to N CTAs of a 16-CTA cluster (B200).
// CTA 0, warp 0: lanes 0..N-1 each send to CTA lane (st.async + complete_tx, or plain remote store + remote arrive);
// CTA N-1 answers with a single st.async back to CTA 0.  Reports issue time of the fan-out instruction and the
// full round trip (fan-out + wake + pong).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_par(unsigned long long* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__global__ void __cluster_dims__(16, 1, 1) fan_kernel(int mode, int iters, int N, int nwarps, long long* out) {
    __shared__ __align__(16) float buf[64];
    __shared__ __align__(8) unsigned long long mb_in, mb_back;
    const uint32_t rank = ctarank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mb_in)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mb_back)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads(); csync();
    const int bytes = (mode == 1 ? 16 : 4) * nwarps;
    long long t_issue = 0, t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        const uint32_t par = i & 1;
        if (rank == 0) {
            if (warp < nwarps) {
                if (warp == 0 && lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 4;" ::"r"(smem_u32(&mb_back)) : "memory");
                const long long a = clock64();
                if (lane < N) {
                    const uint32_t dst = mapa(smem_u32(&buf[warp * 4]), lane), dmb = mapa(smem_u32(&mb_in), lane);
                    if (mode == 0) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst), "r"(i), "r"(dmb) : "memory");
                    else if (mode == 1) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%1,%1,%1}, [%2];" ::"r"(dst), "r"(i), "r"(dmb) : "memory");
                }
                t_issue += clock64() - a;
                if (warp == 0) wait_par(&mb_back, par);
            }
            __syncthreads();
        }
        if (rank != 0 && rank < (uint32_t)N && warp == 0) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mb_in)), "r"(bytes) : "memory");
            wait_par(&mb_in, par);
            if (rank == (uint32_t)N - 1 && lane == 0) {
                const uint32_t dst = mapa(smem_u32(&buf[32]), 0), dmb = mapa(smem_u32(&mb_back), 0);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(dst), "r"(i), "r"(dmb) : "memory");
            }
        }
        if (rank == 0 && N == 1 && warp == 0) { /* self only: nothing answers */ }
    }
    long long t1 = clock64();
    if (rank == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t_issue; }
    __syncthreads(); csync();
}
int main() {
    long long* out; cudaMalloc(&out, 64);
    cudaFuncSetAttribute(fan_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int iters = 2000;
    for (int mode = 0; mode < 2; ++mode)
        for (int N : {2, 4, 8, 16})
            for (int nw : {1, 4}) {
                cudaMemset(out, 0, 64);
                fan_kernel<<<16, 512>>>(mode, iters, N, nw, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
                printf("%s fan-out to %2d CTAs from %d warp(s): issue %.0f cycles, round trip %.0f cycles %s\n", mode ? "st.async.v4 " : "st.async.b32", N, nw,
                       h[1] / (double)iters, h[0] / (double)iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
