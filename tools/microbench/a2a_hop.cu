// Micro-benchmark: period of one all-to-all vector exchange inside a 16-CTA cluster (B200), as a function of the
// message granularity.  Every CTA owns 16 of the 256 values of a vector and needs the whole vector before the next
// step (the v2 kernel's hop).  1024 bytes land in every CTA per hop as
//   mode 0: 256 st.async.b32   (each with its own mbarrier complete_tx)      -- what swn_core_kernel did first
//   mode 1: 128 st.async.b64
//   mode 2:  64 st.async.v4.b32
// The loop is a chain of dependent hops (wait for the vector, publish the next), so cycles / iteration = hop period.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o a2a_hop a2a_hop.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(16, 1, 1) a2a_kernel(int mode, int iters, int waiters, long long* out) {
    __shared__ __align__(16) float vec[2][512];
    __shared__ __align__(8) unsigned long long mbar[2];
    const uint32_t rank = ctarank();
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[i])));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[i])), "r"(mode >= 3 ? 2048 : 1024) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    csync();
    // producers: warps 0-3 (128 lanes).  A value index owned by this CTA: 16*rank + k.
    // mode 0: lane -> (k = tid >> 3, destinations (tid & 7) and (tid & 7) + 8)             2 x b32 per lane
    // mode 1: lane -> (k pair = tid >> 4, destination tid & 15)                            1 x b64 per lane
    // mode 2: lanes with tid < 64 -> (k quad = tid >> 4, destination tid & 15)             1 x v4 per lane, 64 lanes
    long long t0 = 0;
    float acc = (float)tid;
    for (int it = 0; it < iters; ++it) {
        const int b = it & 1;
        if (it == 16 && tid == 0) t0 = clock64();
        if (warp < 4) {
            const uint32_t vb = smem_u32(&vec[b][0]), mb = smem_u32(&mbar[b]);
            if (mode == 0) {
                const int k = 16 * rank + (tid >> 3);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t d = (tid & 7) + 8 * h;
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                                 ::"r"(mapa(vb + 4 * k, d)), "r"(__float_as_uint(acc)), "r"(mapa(mb, d)) : "memory");
                }
            } else if (mode == 1) {
                const int k = 16 * rank + 2 * (tid >> 4);
                const uint32_t d = tid & 15;
                unsigned long long w = ((unsigned long long)__float_as_uint(acc) << 32) | __float_as_uint(acc + 1.f);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                             ::"r"(mapa(vb + 4 * k, d)), "l"(w), "r"(mapa(mb, d)) : "memory");
            } else if (mode == 3 || mode == 4) {
                // two columns (512 values per hop): mode 3 = 512 x b32 (two vectors, as the kernel does), mode 4 = 256 x b64
                // (row-major {col0, col1} pairs, one message per row and destination)
                const int k = 16 * rank + (tid >> 3);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t d = (tid & 7) + 8 * h;
                    if (mode == 3) {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                                         ::"r"(mapa(vb + 4 * (k + 256 * c), d)), "r"(__float_as_uint(acc)), "r"(mapa(mb, d)) : "memory");
                    } else {
                        unsigned long long w = ((unsigned long long)__float_as_uint(acc) << 32) | __float_as_uint(acc + 1.f);
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                                     ::"r"(mapa(vb + 8 * k, d)), "l"(w), "r"(mapa(mb, d)) : "memory");
                    }
                }
            } else if (tid < 64) {
                const int k = 16 * rank + 4 * (tid >> 4);
                const uint32_t d = tid & 15;
                const uint32_t u = __float_as_uint(acc);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];"
                             ::"r"(mapa(vb + 4 * k, d)), "r"(u), "r"(u + 1), "r"(u + 2), "r"(u + 3), "r"(mapa(mb, d)) : "memory");
            }
        }
        // consumers: `waiters` warps sleep on the barrier, then read the vector
        if (warp < waiters) {
            uint32_t done = 0;
            const uint32_t parity = (it >> 1) & 1;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(smem_u32(&mbar[b])), "r"(parity) : "memory");
            acc = acc * 0.5f + vec[b][(tid * 7) & (mode >= 3 ? 511 : 255)];
        }
        __syncthreads();      // everyone has read vec[b] ... (the real kernel has the same CTA-level ordering per step)
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar[b])), "r"(mode >= 3 ? 2048 : 1024) : "memory");
    }
    if (tid == 0) out[rank] = clock64() - t0;
    if (acc == 1234.5f) out[16] = 1;
    csync();
}
int main() {
    long long* d; cudaMalloc(&d, 32 * sizeof(long long));
    long long h[32];
    const int iters = 2016;
    cudaFuncSetAttribute(a2a_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int waiters : {4, 16})
        for (int mode = 0; mode < 5; ++mode) {
            a2a_kernel<<<16, 512>>>(mode, iters, waiters, d);
            if (cudaGetLastError() != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            printf("waiting warps %2d, %s: %.0f cycles per hop\n", waiters,
                   mode == 0 ? "256 x b32  " : mode == 1 ? "128 x b64  " : mode == 2 ? " 64 x v4.b32" : mode == 3 ? "512 x b32 (2 columns)" : "256 x b64 (2 columns)", (double)h[0] / (iters - 16));
        }
    return 0;
}
