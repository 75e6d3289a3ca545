// Micro-benchmark: one-way signalling latency between two CTAs (ping-pong / 2), B200.
//   mode 0: L2, 64-bit {val,tag} word, st.relaxed.gpu + ld.relaxed.gpu poll        (v1 exchange)
//   mode 1: DSMEM, st.relaxed.cluster.shared::cluster.u64 + local ld.volatile.shared poll (v2 exchange)
//   mode 2: DSMEM, st.async ... mbarrier::complete_tx + mbarrier.try_wait.parity
//   mode 3: like mode 1 but `bg` extra warps per CTA spin on local shared memory (poll pressure)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop_latency hop_latency.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

__global__ void __cluster_dims__(16, 1, 1) hop_kernel(int mode, int iters, int partner, int bg, unsigned long long* g, long long* out) {
    __shared__ __align__(16) unsigned long long slot[4];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ volatile int stop;
    const uint32_t rank = ctarank();
    const int tid = threadIdx.x;
    if (tid == 0) {
        slot[0] = slot[1] = 0; stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    csync();
    if (tid >= 32) {                       // background pollers (mode 3): hammer local shared memory
        if (tid < 32 + 32 * bg && (rank == 0 || rank == (uint32_t)partner)) {
            unsigned long long acc = 0;
            while (!stop) {
                unsigned long long a, b;
                asm volatile("ld.volatile.shared.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "r"(smem_u32(&slot[2])));
                acc += a + b;
            }
            if (acc == 12345) out[7] = 1;
        }
    } else if (tid == 0 && (rank == 0 || rank == (uint32_t)partner)) {
        const bool ping = rank == 0;
        const uint32_t peer = ping ? partner : 0;
        const uint32_t rslot = mapa(smem_u32(&slot[0]), peer), rbar = mapa(smem_u32(&mbar), peer);
        unsigned long long* gmine = g + (ping ? 0 : 16), *gpeer = g + (ping ? 16 : 0);
        long long t0 = clock64();
        for (int i = 1; i <= iters; ++i) {
            const unsigned long long word = ((unsigned long long)i << 32) | 7ull;
            if (ping) {
                // send
                if (mode == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(gpeer), "l"(word) : "memory");
                else if (mode == 2) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(rslot), "l"(word), "r"(rbar) : "memory");
                else asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(rslot), "l"(word) : "memory");
            }
            // receive
            if (mode == 0) {
                unsigned long long v;
                do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(gmine) : "memory"); } while ((v >> 32) != (unsigned)i);
            } else if (mode == 2) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 8;" ::"r"(smem_u32(&mbar)) : "memory");
                uint32_t done = 0;
                const uint32_t parity = (i - 1) & 1;
                while (!done)
                    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
            } else {
                unsigned long long v;
                do { asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(&slot[0])) : "memory"); } while ((v >> 32) != (unsigned)i);
            }
            if (!ping) {
                if (mode == 0) asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(gpeer), "l"(word) : "memory");
                else if (mode == 2) asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(rslot), "l"(word), "r"(rbar) : "memory");
                else asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(rslot), "l"(word) : "memory");
            }
        }
        long long t1 = clock64();
        if (ping) out[0] = t1 - t0;
        stop = 1;
    }
    __syncthreads();
    csync();
}

int main() {
    unsigned long long* g; long long* out;
    cudaMalloc(&g, 4096); cudaMalloc(&out, 64);
    cudaFuncSetAttribute(hop_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    const int iters = 2000;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[] = {"L2 LL (st/ld.relaxed.gpu)", "DSMEM LL + local poll", "DSMEM st.async + mbarrier", "DSMEM LL + local poll, bg pollers"};
    for (int mode = 0; mode < 4; ++mode)
        for (int partner : {1, 8, 15})
            for (int bg : {0, 4, 15}) {
                if (mode != 3 && bg != 0) continue;
                if (mode == 3 && bg == 0) continue;
                cudaMemset(g, 0, 4096); cudaMemset(out, 0, 64);
                hop_kernel<<<16, 512>>>(mode == 3 ? 1 : mode, iters, partner, bg, g, out);
                cudaError_t e = cudaDeviceSynchronize();
                long long h[8]; cudaMemcpy(h, out, 64, cudaMemcpyDeviceToHost);
                printf("%-36s partner %2d bg %2d : one-way %.0f cycles (%.0f ns @%d MHz) %s\n", names[mode], partner, bg,
                       h[0] / (2.0 * iters), h[0] / (2.0 * iters) / (clk / 1e6) , clk / 1000, e == cudaSuccess ? "" : cudaGetErrorString(e));
            }
    return 0;
}
