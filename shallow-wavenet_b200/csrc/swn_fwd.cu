// Teacher-forced, time-parallel forward of the two vocoders: what
// CSWNV.forward (reference src/nets/cswnv_shift1.py:191-267) and DSWNV.forward (src/nets/dswnv.py:250-276)
// compute for whole utterances - stride-1 outputs, zero conv history before the first sample (no warm-up).
//
// Unlike the generation loop this IS dense linear algebra: per layer the dilated causal conv is
//   Q[2H x T] = W[2H x K*H] . im2col(h)[K*H x T]
// and the skip / head 1x1 convs are [S x L*H].[L*H x T], [n1 x S].[S x T], [n2 x n1].[n1 x T].  They all run on
// ONE hand-written tcgen05 GEMM kernel (fwd_gemm_kernel):
//   * D[128 rows x 256 times] accumulates in tensor memory; fp32-equivalent precision by three tf32 passes
//     (W_hi.X_hi + W_lo.X_hi + W_hi.X_lo, tcgen05.mma kind::tf32, both operands from shared memory);
//   * A (weights) is pre-split into tf32 hi / lo planes and pre-tiled on the host into the K-major no-swizzle
//     core-matrix layout, one contiguous 16 KB block per (row tile, k block): one cp.async.bulk per stage;
//   * B (activations) lives in HBM time-major as hi / lo planes [rows][ld]; the im2col of the dilated conv is free:
//     tap j of the block is the same tile read `delay_j` rows earlier (rows before the utterance start are zeros =
//     the zero padding of CausalConv1d, cswnv_shift1.py:68-92).  Four loader warps gather it in 16-byte pieces
//     straight into the core-matrix layout;
//   * 2-stage mbarrier pipeline (loaders -> MMA issuer -> tcgen05.commit frees the stage), one MMA-issuing lane; two
//     CTAs per SM (96 KB of shared memory and 256 of the 512 tensor-memory columns each), so that the epilogue of one
//     tile overlaps the main loop of another (a cp.async-pipelined loader with 4 stages and one CTA per SM measured slower);
//   * the epilogue (the same four warps, one TMEM lane = one output row each) is fused: gate with the hoisted
//     conditioning + carry (rows are interleaved z/c so that a lane pair holds one gate pair), or bias + relu, and
//     writes the next operand already split into tf32 hi / lo planes.
// Elementwise helpers: input stage (causal conv over the given audio as a FIR / table gather + softsign),
// expansion of the frame-rate conditioning table to sample rate for one layer, one-hot -> index, Laplace parameters.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "swn_fwd.h"

namespace swn {

namespace {

constexpr int FW_BM = 128, FW_BN = 256, FW_BK = 16, FW_STAGES = 2;
constexpr int FW_A_PLANE = FW_BM * FW_BK * 4;                  // bytes of one (hi or lo) weight block
constexpr int FW_B_PLANE = FW_BN * FW_BK * 4;
constexpr int FW_STAGE_BYTES = 2 * FW_A_PLANE + 2 * FW_B_PLANE; // 48 KB
constexpr int FW_THREADS = 160;                                 // warps 0-3: loaders, then epilogue; warp 4: MMA issuer
constexpr unsigned long long FW_TIMEOUT = 4000000000ULL;        // cycles a barrier wait may take before the kernel gives up

__device__ __forceinline__ uint32_t fw_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fw_mbar_init(unsigned long long* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fw_smem(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void fw_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fw_smem(b)) : "memory");
}
__device__ __forceinline__ void fw_mbar_expect(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fw_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool fw_mbar_try(unsigned long long* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(fw_smem(b)), "r"(parity) : "memory");
    return done != 0;
}
// Wait with a watchdog: a lost signal becomes an error status, never a hang.
__device__ __noinline__ bool fw_mbar_wait(unsigned long long* b, uint32_t parity, int32_t* status, int where) {
    if (fw_mbar_try(b, parity)) return true;
    const long long t0 = clock64();
    uint32_t spins = 0;
    for (;;) {
        if (fw_mbar_try(b, parity)) return true;
        if (((++spins) & 63u) == 0) {
            if (((volatile int32_t*)status)[0] != 0) return false;
            if ((unsigned long long)(clock64() - t0) > FW_TIMEOUT) {
                if (atomicCAS((int*)status, 0, 2) == 0) { status[1] = where; status[2] = (int)parity; status[3] = blockIdx.x; }
                return false;
            }
        }
    }
}
__device__ __forceinline__ void fw_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(fw_smem(dst)), "l"(src), "r"(bytes), "r"(fw_smem(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: 128-byte core matrices (8 rows x 4 tf32), `lbo` bytes to the
// next core matrix along k, 128 bytes to the next group of 8 rows (layout validated by tools/microbench/umma_tf32.cu
// and used by the batch kernel, swn_batch.cuh b3_udesc)
__device__ __forceinline__ uint64_t fw_desc(uint32_t saddr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                        // descriptor version (Blackwell)
    return d;
}
__device__ __forceinline__ void fw_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                 ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}
__device__ __forceinline__ void fw_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fw_smem(bar)) : "memory");
}
__device__ __forceinline__ float fw_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void fw_tmem_ld16(uint32_t a, float* r) {
    uint32_t x[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),
                   "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15])
                 : "r"(a));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(x[i]);
}

}  // namespace

// ------------------------------------------------------------------ the GEMM
__global__ void __launch_bounds__(FW_THREADS, 2) fwd_gemm_kernel(const __grid_constant__ FwdGemm G) {
    extern __shared__ __align__(1024) uint8_t fw_smem_raw[];
    uint8_t* stages = fw_smem_raw;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(fw_smem_raw + FW_STAGES * FW_STAGE_BYTES);
    unsigned long long* full = bars;                    // [STAGES] operands of the stage have landed
    unsigned long long* empty = bars + FW_STAGES;       // [STAGES] the MMAs that read the stage have completed
    unsigned long long* acc_full = bars + 2 * FW_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * FW_STAGES + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // tile -> (row tile mt, utterance b, first time t0); row tiles are the fastest index so that the CTAs that share
    // an activation tile run at the same time (L2 reuse)
    const long long tile = blockIdx.x;
    const int mt = (int)(tile % G.n_mt);
    const long long tt = tile / G.n_mt;
    const int tiles_per_utt = (G.T + FW_BN - 1) / FW_BN;
    const int b = (int)(tt / tiles_per_utt);
    const int t0 = (int)(tt - (long long)b * tiles_per_utt) * FW_BN;
    const long long row_base = (long long)b * G.T;      // first row of the utterance in every activation buffer

    if (tid == 0) {
        for (int s = 0; s < FW_STAGES; ++s) { fw_mbar_init(&full[s], 129); fw_mbar_init(&empty[s], 1); }
        fw_mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(fw_smem(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tmem_slot;
    const int n_kb = G.n_kb, kb_per_tap = G.Cpad / FW_BK;
    bool ok = true;

    if (warp < 4) {
        // =============================== loaders: A by one bulk copy, B gathered into the core-matrix layout
        const float* Ablk = G.A + (size_t)mt * n_kb * (2 * FW_BM * FW_BK);
        // piece p of the block: row n = p >> 2 (time), kq = p & 3 (4 channels = 16 bytes); 128 threads cover 32 rows
        const int kq = tid & 3, nr = tid >> 2;
        for (int kb = 0; kb < n_kb && ok; ++kb) {
            const int s = kb % FW_STAGES;
            const uint32_t ph = (uint32_t)(kb / FW_STAGES) & 1u;
            ok = fw_mbar_wait(&empty[s], ph ^ 1u, G.status, 1);
            if (!ok) break;
            uint8_t* st = stages + s * FW_STAGE_BYTES;
            if (tid == 0) {
                fw_mbar_expect(&full[s], 2 * FW_A_PLANE);
                fw_bulk_g2s(st, Ablk + (size_t)kb * (2 * FW_BM * FW_BK), 2 * FW_A_PLANE, &full[s]);
            }
            const int tap = kb / kb_per_tap;
            const int c0 = (kb - tap * kb_per_tap) * FW_BK + kq * 4;          // channel of this piece inside the tap
            const int delay = G.delay[tap];
            const bool cvalid = c0 < G.C;                                      // (buffers are padded to 4 channels with zeros)
            const size_t col = (size_t)G.b_col0 + (size_t)tap * G.tap_col_stride + c0;
            float4* bh = reinterpret_cast<float4*>(st + 2 * FW_A_PLANE);
            float4* bl = reinterpret_cast<float4*>(st + 2 * FW_A_PLANE + FW_B_PLANE);
            float4 vh[FW_BN / 32], vl[FW_BN / 32];
#pragma unroll
            for (int i = 0; i < FW_BN / 32; ++i) {
                const int n = nr + 32 * i;
                const int ts = t0 + n - delay;                                 // source time inside the utterance
                vh[i] = make_float4(0.f, 0.f, 0.f, 0.f); vl[i] = vh[i];
                if (cvalid && ts >= 0 && ts < G.T) {
                    const size_t off = (size_t)(row_base + ts) * G.ldb + col;
                    vh[i] = __ldg(reinterpret_cast<const float4*>(G.Bhi + off));
                    vl[i] = __ldg(reinterpret_cast<const float4*>(G.Blo + off));
                }
            }
#pragma unroll
            for (int i = 0; i < FW_BN / 32; ++i) {
                const int n = nr + 32 * i;
                const int o = (kq * (FW_BN >> 3) + (n >> 3)) * 8 + (n & 7);    // float4 index: [k/4][n/8][n%8]
                bh[o] = vh[i]; bl[o] = vl[i];
            }
            asm volatile("fence.proxy.async;" ::: "memory");       // generic-proxy stores -> visible to the MMA
            fw_mbar_arrive(&full[s]);
        }
    } else if (lane == 0) {
        // =============================== MMA issuer: D += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo per 8-wide k step
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(FW_BN >> 3) << 17) | ((uint32_t)(FW_BM >> 4) << 24);
        constexpr uint32_t A_LBO = (FW_BM >> 3) * 128, B_LBO = (FW_BN >> 3) * 128;
        for (int kb = 0; kb < n_kb && ok; ++kb) {
            const int s = kb % FW_STAGES;
            const uint32_t ph = (uint32_t)(kb / FW_STAGES) & 1u;
            ok = fw_mbar_wait(&full[s], ph, G.status, 2);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = fw_smem(stages + s * FW_STAGE_BYTES);
            const uint32_t a_hi = sa, a_lo = sa + FW_A_PLANE, b_hi = sa + 2 * FW_A_PLANE, b_lo = b_hi + FW_B_PLANE;
#pragma unroll
            for (int ks = 0; ks < FW_BK / 8; ++ks) {
                const uint64_t dah = fw_desc(a_hi + ks * 2 * A_LBO, A_LBO), dal = fw_desc(a_lo + ks * 2 * A_LBO, A_LBO);
                const uint64_t dbh = fw_desc(b_hi + ks * 2 * B_LBO, B_LBO), dbl = fw_desc(b_lo + ks * 2 * B_LBO, B_LBO);
                fw_mma_ss(tbase, dah, dbh, idesc, (kb | ks) != 0);
                fw_mma_ss(tbase, dal, dbh, idesc, 1);
                fw_mma_ss(tbase, dah, dbl, idesc, 1);
            }
            fw_commit(&empty[s]);
        }
        if (ok) fw_commit(acc_full);
    }

    if (warp < 4) {
        // =============================== epilogue: thread = TMEM lane = output row of the tile
        ok = ok && fw_mbar_wait(acc_full, 0, G.status, 3);
        ok = __all_sync(0xffffffffu, ok);
        if (ok) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tbase + ((uint32_t)(32 * warp) << 16);
            const int r = tid;                                              // row inside the tile
            if (G.mode == FWD_GATE) {
                // rows are interleaved: r = 2 * i + {0: gate z, 1: candidate} of pair mt * 64 + i (cswnv_shift1.py:275-279)
                const int pair = mt * 64 + (r >> 1);
                const bool pv = pair < G.H;
                const int zc = r & 1;
                const float bias = pv ? __ldg(G.bias + zc * G.H + pair) : 0.f;
                for (int c0 = 0; c0 < FW_BN; c0 += 16) {
                    float v[16];
                    fw_tmem_ld16(taddr + c0, v);
                    // the z lane keeps times c0..c0+7, the candidate lane c0+8..c0+15: both end up with (qz, qc) of 8 times
                    float qz[8], qc[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float mine_lo = v[i] + bias, mine_hi = v[8 + i] + bias;
                        const float other = __shfl_xor_sync(0xffffffffu, zc ? mine_lo : mine_hi, 1);
                        qz[i] = zc ? other : mine_lo;
                        qc[i] = zc ? mine_hi : other;
                    }
                    if (pv) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int t = t0 + c0 + zc * 8 + i;
                            if (t < G.T) {
                                const size_t row = (size_t)(row_base + t);
                                const float pz = __ldg(G.p + row * (2 * G.H) + pair), pc = __ldg(G.p + row * (2 * G.H) + G.H + pair);
                                const size_t ci = row * G.ldb + G.carry_col0 + pair;
                                const float carry = __ldg(G.Bhi + ci) + __ldg(G.Blo + ci);
                                const float z = 1.f / (1.f + expf(-(pz * qz[i])));
                                const float hn = (1.f - z) * tanhf(pc * qc[i]) + z * carry;
                                const float h = fw_hi(hn);
                                const size_t oi = row * G.ldo + G.out_col0 + pair;
                                G.out_hi[oi] = h; G.out_lo[oi] = hn - h;
                            }
                        }
                    }
                }
            } else {
                const int orow = mt * FW_BM + r;
                const bool rv = orow < G.rows;
                const float bias = rv ? __ldg(G.bias + orow) : 0.f;
                for (int c0 = 0; c0 < FW_BN; c0 += 16) {
                    float v[16];
                    fw_tmem_ld16(taddr + c0, v);
                    if (rv) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int t = t0 + c0 + i;
                            if (t < G.T) {
                                const size_t row = (size_t)(row_base + t);
                                float x = v[i] + bias;
                                if (G.mode == FWD_RELU) {
                                    x = fmaxf(x, 0.f);
                                    const float h = fw_hi(x);
                                    const size_t oi = row * G.ldo + G.out_col0 + orow;
                                    G.out_hi[oi] = h; G.out_lo[oi] = x - h;
                                } else {
                                    G.out_f[row * G.ldf + orow] = x;
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase) : "memory");
}

cudaError_t launch_fwd_gemm(const FwdGemm& g, cudaStream_t st) {
    static bool attr_set = false;
    const size_t smem = (size_t)FW_STAGES * FW_STAGE_BYTES + 128;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(fwd_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long tiles = (long long)g.n_mt * g.nbatch * ((g.T + FW_BN - 1) / FW_BN);
    if (tiles <= 0 || tiles > 0x7fffffffLL) return cudaErrorInvalidValue;
    fwd_gemm_kernel<<<(unsigned)tiles, FW_THREADS, smem, st>>>(g);
    return cudaGetLastError();
}

// Host-side packing of a weight matrix W[rows][ktot] (row-major, ktot = ntaps * cpad with zero padding) into the
// kernel's A operand: [row tile][k block][hi | lo][k/4][r/8][r%8][k%4].  `row_of(tile_row)` maps a tile row to the
// logical row (or -1 for a zero row) so that gate rows can be interleaved.
void pack_fwd_weights(const float* W, int ktot, int n_mt, const int* row_map, float* out) {
    const int n_kb = ktot / FW_BK;
    const size_t blk = (size_t)2 * FW_BM * FW_BK;
    for (int mt = 0; mt < n_mt; ++mt)
        for (int kb = 0; kb < n_kb; ++kb) {
            float* hi = out + ((size_t)mt * n_kb + kb) * blk;
            float* lo = hi + FW_BM * FW_BK;
            for (int r = 0; r < FW_BM; ++r) {
                const int src = row_map[mt * FW_BM + r];
                for (int k = 0; k < FW_BK; ++k) {
                    const float w = src >= 0 ? W[(size_t)src * ktot + kb * FW_BK + k] : 0.f;
                    uint32_t u;
                    memcpy(&u, &w, 4);
                    u &= 0xFFFFE000u;
                    float h;
                    memcpy(&h, &u, 4);
                    const size_t o = ((size_t)(k >> 2) * (FW_BM >> 3) + (r >> 3)) * 32 + (r & 7) * 4 + (k & 3);
                    hi[o] = h; lo[o] = w - h;
                }
            }
        }
}
int fwd_bk() { return FW_BK; }
int fwd_bm() { return FW_BM; }

// ------------------------------------------------------------------ elementwise helpers
// one-hot float [B][Q][T] -> class index [B][T]; bad[0] counts columns that are not one-hot
__global__ void fwd_onehot_kernel(const float* __restrict__ x, int Q, int T, long long n, int32_t* __restrict__ idx,
                                  int32_t* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long b = i / T;
    const int t = (int)(i - b * T);
    int best = -1, ones = 0;
    bool other = false;
    for (int q = 0; q < Q; ++q) {
        const float v = __ldg(x + ((size_t)b * Q + q) * T + t);
        if (v == 1.f) { ++ones; if (best < 0) best = q; }
        else if (v != 0.f) other = true;
    }
    if (ones != 1 || other) { atomicAdd(bad, 1); if (best < 0) best = 0; }
    idx[i] = best;
}

// int64 indices -> int32 with range check (the reference's OneHot does x % depth, dswnv.py:77)
__global__ void fwd_index_kernel(const long long* __restrict__ x, int Q, long long n, int32_t* __restrict__ idx) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long v = x[i] % Q;
    if (v < 0) v += Q;
    idx[i] = (int32_t)v;
}

// h0[t] = softsign(b_c + sum_j tap_j(a[t + seg - 1 - (K-1) + j])), zero conv padding before the first sample
// (cswnv_shift1.py:204-207 with [:, :, seg-1:]; dswnv.py:258-261).  Writes tf32 hi / lo planes, column block 0.
__global__ void fwd_input_kernel(FwdInput I) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)I.nbatch * I.T * I.H;
    if (i >= n) return;
    const int r = (int)(i % I.H);
    const long long bt = i / I.H;
    const int b = (int)(bt / I.T), t = (int)(bt - (long long)b * I.T);
    float acc = __ldg(I.in_bias + r);
    for (int j = 0; j < I.K; ++j) {
        const int tau = t + I.seg - 1 - (I.K - 1) + j;
        if (tau < 0) continue;
        if (I.kind == 0) {
            const float a = __ldg(I.audio_f + (size_t)b * I.Ta + tau);
            acc += __ldg(I.in_tab + (j * 2) * I.H + r) + __ldg(I.in_tab + (j * 2 + 1) * I.H + r) * a;
        } else {
            const int q = __ldg(I.audio_i + (size_t)b * I.Ta + tau);
            acc += __ldg(I.in_tab + ((size_t)j * I.Q + q) * I.H + r);
        }
    }
    const float v = acc / (1.f + fabsf(acc));
    const float h = fw_hi(v);
    const size_t o = (size_t)bt * I.ld + r;
    I.out_hi[o] = h; I.out_lo[o] = v - h;
}

// p_l[t][row] = const_l[row] + sum_s gain[(t + s + seg) % U] * P[b][(t + s + seg) / U][l][s][row]
// (the seg-window x'[c*seg+s, t] = x_up[c, t+s+seg] of forward(), cswnv_shift1.py:193-201; dswnv.py:252 with seg = 1),
// plus the one-hot columns of in_x when audio_in_flag (dswnv.py:254-255).
__global__ void fwd_cond_kernel(FwdCond C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int twoH = 2 * C.H;
    const long long n = (long long)C.nbatch * C.T * twoH;
    if (i >= n) return;
    const int row = (int)(i % twoH);
    const long long bt = i / twoH;
    const int b = (int)(bt / C.T), t = (int)(bt - (long long)b * C.T);
    float acc = __ldg(C.pconst + C.l * twoH + row);
    for (int s = 0; s < C.seg; ++s) {
        const int tau = t + s + C.seg;
        const int f = tau / C.U, j = tau - f * C.U;
        acc = fmaf(__ldg(C.gain + j), __ldg(C.P + ((((size_t)b * C.Tf + f) * C.L + C.l) * C.seg + s) * twoH + row), acc);
    }
    if (C.xa_tab) acc += __ldg(C.xa_tab + ((size_t)C.l * C.Q + __ldg(C.audio_i + (size_t)b * C.Ta + t)) * twoH + row);
    C.out[i] = acc;
}

// Laplace parameters of CSWNV.forward (cswnv_shift1.py:229-242): log_b = logsigmoid(out[seg:2seg]), b = exp(log_b), and
// the clipped pair of clip=True (clamp at log(7.07e-7)).
__global__ void fwd_laplace_kernel(const float* __restrict__ out, int n2, int seg, long long rows, float* __restrict__ bb,
                                   float* __restrict__ log_b, float* __restrict__ b_clip, float* __restrict__ log_b_clip) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * seg) return;
    const long long r = i / seg;
    const int s = (int)(i - r * seg);
    const float x = out[r * n2 + seg + s];
    const float lb = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
    bb[i] = expf(lb);
    log_b[i] = lb;
    if (b_clip) {
        const float lc = fmaxf(lb, -14.162084148244246758816564788835f);
        log_b_clip[i] = lc; b_clip[i] = expf(lc);
    }
}

static inline unsigned fw_blocks(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

cudaError_t launch_fwd_onehot(const float* x, int B, int Q, int T, int32_t* idx, int32_t* bad, cudaStream_t st) {
    const long long n = (long long)B * T;
    fwd_onehot_kernel<<<fw_blocks(n, 256), 256, 0, st>>>(x, Q, T, n, idx, bad);
    return cudaGetLastError();
}
cudaError_t launch_fwd_index(const long long* x, int Q, long long n, int32_t* idx, cudaStream_t st) {
    fwd_index_kernel<<<fw_blocks(n, 256), 256, 0, st>>>(x, Q, n, idx);
    return cudaGetLastError();
}
cudaError_t launch_fwd_input(const FwdInput& I, cudaStream_t st) {
    fwd_input_kernel<<<fw_blocks((long long)I.nbatch * I.T * I.H, 256), 256, 0, st>>>(I);
    return cudaGetLastError();
}
cudaError_t launch_fwd_cond(const FwdCond& C, cudaStream_t st) {
    fwd_cond_kernel<<<fw_blocks((long long)C.nbatch * C.T * 2 * C.H, 256), 256, 0, st>>>(C);
    return cudaGetLastError();
}
cudaError_t launch_fwd_laplace(const float* out, int n2, int seg, long long rows, float* b, float* log_b, float* b_clip,
                               float* log_b_clip, cudaStream_t st) {
    fwd_laplace_kernel<<<fw_blocks(rows * seg, 256), 256, 0, st>>>(out, n2, seg, rows, b, log_b, b_clip, log_b_clip);
    return cudaGetLastError();
}

}  // namespace swn
