// Standalone check of the tensor-core contraction planned for the batched gated cell:
//   D[128 x N] = W[128 x K] . X[K x N]  in fp32-equivalent precision by three tf32 passes
//   (W_hi.X_hi + W_lo.X_hi + W_hi.X_lo), tcgen05.mma kind::tf32, cta_group::1,
//   A (weights, hi and lo halves) resident in TENSOR MEMORY, B (activations) in shared memory in the
//   MN-major no-swizzle canonical layout, D in tensor memory, drained with tcgen05.ld.
// Prints the max error against a double-precision reference for 1 pass and 3 passes and the cycles per tile.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_tf32 umma_tf32.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

constexpr int M = 128, K = 192, N = 40, NP = 48;             // NP: N padded to a multiple of 16 (UMMA M=128 constraint)
constexpr int KG = K / 8;                                    // k-steps of one tf32 MMA (K = 8)
constexpr int SBO = 128, LBO = (NP / 4) * 128;               // bytes: next 4-column group / next 8-k group
constexpr int TILE_BYTES = KG * LBO;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, int bk = 0) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(((bk ? (NP / 8) * 128 : LBO) >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    return d;                                                // layout type 0 = no swizzle
}

__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

__global__ void __launch_bounds__(128, 1) umma_test(const float* __restrict__ W, const float* __restrict__ X, float* __restrict__ out1,
                                                    float* __restrict__ out3, long long* cycles, int reps, int bk, float* __restrict__ dbg, float* __restrict__ outs, float* __restrict__ dump) {
    extern __shared__ __align__(1024) uint8_t smem[];
    float* bhi = reinterpret_cast<float*>(smem);
    float* blo = reinterpret_cast<float*>(smem + TILE_BYTES);
    float* ahi = reinterpret_cast<float*>(smem + 2 * TILE_BYTES);        // K-major canonical tile of W_hi (SS form check)
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_slot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);

    // ---- A: row = TMEM lane, k = column; hi half at columns [0,K), lo half at [K,2K)
    {
        const float* wr = W + (size_t)tid * K;
        for (int k0 = 0; k0 < K; k0 += 4) {
            uint32_t hi[4], lo[4];
            for (int i = 0; i < 4; ++i) {
                const float w = wr[k0 + i];
                const uint32_t h = __float_as_uint(w) & 0xFFFFE000u;           // tf32: sign, 8 exponent, 10 mantissa bits
                hi[i] = h;
                lo[i] = __float_as_uint(w - __uint_as_float(h));
            }
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lane_base + k0), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lane_base + K + k0), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // ---- B: X[k][n] -> canonical MN-major tiles (8 k x 4 n core matrices of 128 bytes), hi and lo
    for (int i = tid; i < TILE_BYTES / 4; i += 128) { bhi[i] = 0.f; blo[i] = 0.f; }
    __syncthreads();
    for (int e = tid; e < K * N; e += 128) {
        const int k = e / N, n = e - k * N;
        const float x = X[e];
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        const int off = bk ? (((k >> 2) * (NP / 8) * 128 + (n >> 3) * 128 + (n & 7) * 16 + (k & 3) * 4) >> 2)
                           : (((k >> 3) * LBO + (n >> 2) * SBO + (k & 7) * 16 + (n & 3) * 4) >> 2);
        bhi[off] = h;
        blo[off] = x - h;
    }
    for (int e = tid; e < M * K; e += 128) {
        const int r = e / K, k = e - r * K;
        const float h = __uint_as_float(__float_as_uint(W[e]) & 0xFFFFE000u);
        ahi[((k >> 2) * 2048 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4) >> 2] = h;
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // read A back from tensor memory (sanity of the st / ld addressing)
    {
        uint32_t v[4];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(lane_base + 8));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        dbg[tid] = __uint_as_float(v[1]);                                   // W_hi[tid][9]
    }

    // instruction descriptor: D fp32, A/B tf32, A K-major, B MN-major, N = NP, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((bk ? 0u : 1u) << 16) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t d_tmem = tbase + 2 * K;                                     // columns [2K, 2K+NP)
    {
        const uint32_t seven = __float_as_uint(7.0f);
        for (int c0 = 0; c0 < 64; c0 += 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lane_base + 2 * K + c0), "r"(seven), "r"(seven), "r"(seven), "r"(seven) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    uint32_t phase = 0;
    for (int mode = 0; mode < 3; ++mode) {                                     // 0: one pass (hi.hi), 1: three passes, 2: one pass, A from smem
        long long t0 = 0;
        const int nrep = mode == 1 ? reps : 1;
        if (tid == 0) t0 = clock64();
        for (int rep = 0; rep < nrep; ++rep) {
            if (tid == 0) {
                for (int ks = 0; ks < KG; ++ks) {
                    const int kstep = bk ? 2 * (NP / 8) * 128 : LBO;
                    const uint64_t dh = make_desc(smem_u32(bhi) + ks * kstep, bk), dl = make_desc(smem_u32(blo) + ks * kstep, bk);
                    if (mode == 2) {
                        uint64_t da = 0;
                        const uint32_t sa = smem_u32(ahi) + ks * 2 * 2048;
                        da |= (uint64_t)((sa >> 4) & 0x3FFF);
                        da |= (uint64_t)((2048 >> 4) & 0x3FFF) << 16;          // LBO: next 4-k core matrix
                        da |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;           // SBO: next 8-row group
                        da |= (uint64_t)1 << 46;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                                     ::"r"(d_tmem), "l"(da), "l"(dh), "r"(idesc), "r"((uint32_t)(ks > 0)), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
                        continue;
                    }
                    mma_tf32_ts(d_tmem, tbase + ks * 8, dh, idesc, ks > 0);
                    if (mode == 1) {
                        mma_tf32_ts(d_tmem, tbase + K + ks * 8, dh, idesc, 1);
                        mma_tf32_ts(d_tmem, tbase + ks * 8, dl, idesc, 1);
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            // everybody waits for the MMAs
            uint32_t done = 0;
            while (!done)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar)), "r"(phase) : "memory");
            phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        if (tid == 0 && mode == 1) cycles[0] = (clock64() - t0) / nrep;
        if (mode == 0 && dump) {
            for (int c0 = 0; c0 < 512; c0 += 16) {
                uint32_t v[16];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                               "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(lane_base + c0));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                for (int i = 0; i < 16; ++i) dump[(size_t)tid * 512 + c0 + i] = __uint_as_float(v[i]);
            }
        }
        // ---- drain D: thread = row, NP columns
        float* out = mode == 0 ? out1 : (mode == 1 ? out3 : outs);
        for (int c0 = 0; c0 < NP; c0 += 16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(lane_base + 2 * K + c0));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 16; ++i)
                if (c0 + i < N) out[(size_t)tid * N + c0 + i] = __uint_as_float(v[i]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main() {
    std::vector<float> W((size_t)M * K), X((size_t)K * N);
    srand(1);
    for (auto& v : W) v = (rand() / (float)RAND_MAX - 0.5f) * 0.3f;
    for (auto& v : X) v = (rand() / (float)RAND_MAX - 0.5f) * 2.0f;
    float *dW, *dX, *d1, *d3, *dd, *ds; long long* dc;
    cudaMalloc(&dd, 128 * 4); cudaMalloc(&ds, (size_t)M * N * 4); cudaMemset(ds, 0, (size_t)M * N * 4);
    cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&d1, (size_t)M * N * 4); cudaMalloc(&d3, (size_t)M * N * 4); cudaMalloc(&dc, 8);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(d1, 0, (size_t)M * N * 4); cudaMemset(d3, 0, (size_t)M * N * 4);
    const int smem = 2 * TILE_BYTES + M * K * 4;
    cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    float* ddump; cudaMalloc(&ddump, 128 * 512 * 4); cudaMemset(ddump, 0, 128 * 512 * 4);
    const int bk = getenv("BK") ? atoi(getenv("BK")) : 0;
    printf("B operand: %s-major\n", bk ? "K" : "MN");
    umma_test<<<1, 128, smem>>>(dW, dX, d1, d3, dc, 200, bk, dd, ds, ddump);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> o1((size_t)M * N), o3((size_t)M * N);
    long long cyc = 0;
    cudaMemcpy(o1.data(), d1, o1.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o3.data(), d3, o3.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    std::vector<float> dbg(128), os((size_t)M * N);
    cudaMemcpy(dbg.data(), dd, 128 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(os.data(), ds, os.size() * 4, cudaMemcpyDeviceToHost);
    {
        int bad = 0;
        for (int r = 0; r < 128; ++r) {
            uint32_t u; memcpy(&u, &W[(size_t)r * K + 9], 4); u &= 0xFFFFE000u; float h; memcpy(&h, &u, 4);
            if (dbg[r] != h) ++bad;
        }
        printf("A readback from TMEM: %d of 128 rows differ (row 3: got %.6f want ~%.6f)\n", bad, dbg[3], W[3 * K + 9]);
    }
    {
        std::vector<float> dump(128 * 512);
        cudaMemcpy(dump.data(), ddump, dump.size() * 4, cudaMemcpyDeviceToHost);
        // which columns differ from what was stored (A hi / lo) or are non-zero beyond them
        int nz_d = 0;
        for (int c = 384; c < 512; ++c) for (int r = 0; r < 128; ++r) if (dump[r * 512 + c] != 0.f) ++nz_d;
        int changed_a = 0;
        for (int r = 0; r < 128; ++r) for (int c = 0; c < 192; ++c) {
            uint32_t u; memcpy(&u, &W[(size_t)r * K + c], 4); u &= 0xFFFFE000u; float h; memcpy(&h, &u, 4);
            if (dump[r * 512 + c] != h) ++changed_a;
        }
        printf("TMEM dump after the first batch: %d non-zero words in columns [384,512), %d words of A_hi changed\n", nz_d, changed_a);
        printf("row 5, columns 384..391: "); for (int c = 384; c < 392; ++c) printf("%g ", dump[5 * 512 + c]); printf("\n");
    }
    double es = 0;
    double e1 = 0, e3 = 0, ef = 0, scale = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0; float f = 0.f;
            for (int k = 0; k < K; ++k) { r += (double)W[(size_t)m * K + k] * X[(size_t)k * N + n]; f = fmaf(W[(size_t)m * K + k], X[(size_t)k * N + n], f); }
            scale = fmax(scale, fabs(r));
            e1 = fmax(e1, fabs(o1[(size_t)m * N + n] - r));
            e3 = fmax(e3, fabs(o3[(size_t)m * N + n] - r));
            es = fmax(es, fabs(os[(size_t)m * N + n] - r));
            ef = fmax(ef, fabs(f - r));
        }
    printf("max |ref| %.4f ; max abs error: 1 tf32 pass %.3e, 3 tf32 passes %.3e, fp32 fmaf chain %.3e\n", scale, e1, e3, ef);
    printf("A from shared memory (1 pass): max abs error %.3e ; outs[5][7] = %.7f\n", es, os[5 * N + 7]);
    printf("sample: out3[5][7] = %.7f  out1[5][7] = %.7f\n", o3[5 * N + 7], o1[5 * N + 7]);
    printf("cycles per 128x%dx%d tile, 3 passes (%d MMAs incl. commit + wait): %lld\n", NP, K, 3 * KG, cyc);
    const bool ok = e3 < 5e-6 * fmax(scale, 1.0) && e1 < 5e-3 * fmax(scale, 1.0);
    printf(ok ? "UMMA_OK\n" : "UMMA_MISMATCH\n");
    return ok ? 0 : 2;
}
