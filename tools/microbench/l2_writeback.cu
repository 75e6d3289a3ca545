// Micro-benchmark: when does a buffer that is rewritten in place and consumed from L2 microseconds later reach DRAM?
// (The batch kernel's exchange tiles - 5.2 MB rewritten every 77 us - were written back to HBM in full: 123 GB per cfg3
// launch, profiles/r01_v3e_traffic_cfg3.csv.)  Half of the CTAs produce tiles, the other half consume the tile produced one
// round earlier; run under  ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,gpu__time_duration.sum  per mode.
//   store: 0 = st.global.cg   1 = st.global (write-back default)   2 = st.global with an L2::evict_last cache hint
//          6 = st.global.cg of single floats, every instruction touching 4 of the 32 bytes of a sector (the scattered 4-byte
//              stores with which the gated cells publish their tf32 hi / lo tiles)
//          3 = st.global.cg with an L2::evict_last cache hint   4 = red.global.add.v4.f32 (fire-and-forget reduction, as the
//          batch kernel's scatter into its L2-resident rings)   5 = ld.global.cg + add + st.global.cg (same update, no reduction)
//   load : 0 = none   1 = ld.global.cg   2 = cp.async.bulk (TMA engine) into shared memory
//          4 = cp.async.bulk with an L2::evict_last cache hint
//          3 = EVERY consumer CTA reads EVERY tile with ld.global.cg (as the 24 CTAs of a cell group all pull the same tile:
//              readers on both dies)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_writeback l2_writeback.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(256) probe(float4* buf, unsigned int* flags, int n_tiles, int tile_f4, int rounds, int smode, int lmode,
                                             float* sink, int sleep_ns, const float4* table, long long table_f4, int pollute_f4, int cyclic) {
    extern __shared__ __align__(128) float4 stage[];
    __shared__ __align__(8) unsigned long long bar;
    const int half = gridDim.x / 2;
    const bool producer = (int)blockIdx.x < half;
    const int lane_cta = producer ? blockIdx.x : blockIdx.x - half;
    unsigned long long pol = 0;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (!producer && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    uint32_t phase = 0;
    for (int r = 0; r < rounds; ++r) {
        // idle time between two rewrites of a tile (the batch kernel rewrites a tile every ~77 us)
        for (int z = 0; z < sleep_ns; z += 1000) __nanosleep(1000);
        if (!producer && pollute_f4 > 0) {
            // background stream through L2 (the conditioning table of the real kernel): pollute_f4 float4 per CTA and round
            const long long base = ((long long)r * half + lane_cta) * pollute_f4 % (table_f4 - pollute_f4);
            for (int i = threadIdx.x; i < pollute_f4; i += blockDim.x) { const float4 v = __ldg(table + base + i); acc += v.y; }
        }
        const int t_step = (!producer && lmode == 3) ? 1 : half;
        // cyclic: a round only touches `half` tiles, the next round the next ones (a large ring buffer that is walked
        // through once every n_tiles / half rounds - the dilation-49 rings of the batch kernel: 59 MB, one lap per 60 iterations)
        const int t_first = cyclic ? (int)(((long long)r * half + lane_cta) % n_tiles) : ((!producer && lmode == 3) ? 0 : lane_cta);
        const int t_last = cyclic ? t_first + 1 : n_tiles;
        for (int t = t_first; t < t_last; t += t_step) {
            float4* tile = buf + (size_t)t * tile_f4;
            if (producer) {
                if (smode == 6) {
                    float* tf = reinterpret_cast<float*>(tile);
                    const int nsec = tile_f4 / 2;                       // 32-byte sectors
                    for (int j = 0; j < 8; ++j)
                        for (int i = threadIdx.x; i < nsec; i += blockDim.x) __stcg(tf + (size_t)i * 8 + j, (float)(r + j));
                } else
                for (int i = threadIdx.x; i < tile_f4; i += blockDim.x) {
                    const float4 v = make_float4((float)r, (float)i, (float)t, 1.f);
                    float4* p = tile + i;
                    if (smode == 0) __stcg(p, v);
                    else if (smode == 1) *p = v;
                    else if (smode == 2)
                        asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
                    else if (smode == 4)
                        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                    else if (smode == 5) {
                        float4 o = __ldcg(p);
                        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
                        __stcg(p, o);
                    } else
                        asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    __threadfence();
                    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + t), "r"((unsigned)r + 1u) : "memory");
                }
            } else if (lmode != 0) {
                if (threadIdx.x == 0) {
                    unsigned v;
                    do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + t) : "memory"); } while (v < (unsigned)r + 1u);
                }
                __syncthreads();
                if (lmode == 1 || lmode == 3) {
                    for (int i = threadIdx.x; i < tile_f4; i += blockDim.x) { const float4 v = __ldcg(tile + i); acc += v.x + v.w; }
                } else {
                    if (threadIdx.x == 0) {
                        const uint32_t bytes = (uint32_t)tile_f4 * 16u;
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
                        if (lmode == 4)
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                                         ::"r"(smem_u32(stage)), "l"(tile), "r"(bytes), "r"(smem_u32(&bar)), "l"(pol) : "memory");
                        else
                            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                         ::"r"(smem_u32(stage)), "l"(tile), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
                    }
                    uint32_t done = 0;
                    while (!done)
                        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
                    phase ^= 1u;
                    acc += stage[threadIdx.x % tile_f4].x;
                    __syncthreads();
                }
            }
        }
    }
    if (acc == 12345.678f) sink[0] = acc;
}

int main(int argc, char** argv) {
    const int smode = argc > 1 ? atoi(argv[1]) : 0, lmode = argc > 2 ? atoi(argv[2]) : 0;
    const int rounds = argc > 3 ? atoi(argv[3]) : 2000;
    const int sleep_ns = argc > 4 ? atoi(argv[4]) : 0;
    const int pollute_kb = argc > 5 ? atoi(argv[5]) : 0;     // KB streamed through L2 per round (all consumers together)
    const int cyclic = argc > 7 ? atoi(argv[7]) : 0;
    const int n_tiles = argc > 6 ? atoi(argv[6]) : 72, tile_bytes = 72 * 1024, tile_f4 = tile_bytes / 16;        // 5.2 MB in 72 tiles of 72 KB
    float4* buf; unsigned int* flags; float* sink;
    cudaMalloc(&buf, (size_t)n_tiles * tile_bytes); cudaMalloc(&flags, n_tiles * 4); cudaMalloc(&sink, 4);
    const long long table_bytes = 1LL << 30; float4* table; cudaMalloc(&table, table_bytes); cudaMemset(table, 0, table_bytes);
    cudaMemset(buf, 0, (size_t)n_tiles * tile_bytes); cudaMemset(flags, 0, n_tiles * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, tile_bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<<<144, 256, tile_bytes>>>(buf, flags, n_tiles, tile_f4, rounds, smode, lmode, sink, sleep_ns, table, table_bytes / 16, pollute_kb * 1024 / 16 / 72, cyclic);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    printf("tiles %d cyclic %d | pollute %d KB/round | sleep %d ns | store mode %d load mode %d: %d rounds x %.1f MB rewritten = %.1f GB of stores, %.2f ms (%s)\n", n_tiles, cyclic, pollute_kb, sleep_ns, smode, lmode, rounds,
           n_tiles * tile_bytes / 1e6, (double)rounds * n_tiles * tile_bytes / 1e9, ms, cudaGetErrorString(e));
    return e != cudaSuccess;
}
