// Pipelined batch generation kernel (variant 3) of the B200-native shallow-wavenet path.
//
// Same algorithm as swn_gen.cuh (CSWNV.batch_fast_generate, src/nets/cswnv_shift1.py:314-402;
// DSWNV.batch_fast_generate, src/nets/dswnv.py:312-374) but organised for THROUGHPUT over many
// utterances: the batch is cut into M pipeline slots of BS utterances (NC = SEG*BS columns);
// the L layer groups and the head CTAs form a ring pipeline, every stage working on a different
// slot at any time.  Differences to variant 1:
//   * a layer CTA evaluates all taps of its gate pairs (and its out_skip rows) as ONE
//     register-tiled fp32 GEMM  [rows x H] . [H x NC]  against weights that stay in shared memory
//     (k-major, so a thread reads its TR rows with 128-bit loads) - no shuffle reductions;
//   * the CTA is warp-specialised: GEMM warps, one loader lane that waits for the upstream
//     counter and pulls the next slot's input tile with one TMA bulk copy (double buffered,
//     mbarrier completion), and epilogue warps that prefetch conditioning + ring partial sums,
//     apply the gate, publish, and scatter the delayed taps into the ring off the critical path;
//   * tiles travel through L2 as plain fp32 [H][NC] blocks plus one monotonic counter per
//     (stage, slot) (release add / acquire poll) instead of {value,tag} words;
//   * out_skip of the last layer is evaluated by the last layer's CTAs on the rows they own
//     (partial sums over k, added by the head in a fixed order); every head CTA serves whole
//     slots on its own, streaming out_1 / out_2 through a TMA ring.
// Column order inside a slot: col = s*BS + b (time-major), so that the BS utterances of one time
// column are contiguous in tiles and rings.
#pragma once
#include "swn_gen.cuh"

namespace swn {

constexpr int B3_NGW = 8;                 // GEMM warps 0..7
constexpr int B3_LOADER = 8;              // warp 8: lane 0 is the loader
constexpr int B3_EPI0 = 9;                // epilogue warps 9..15
constexpr int B3_NEPI = (SWN_NWARPS - B3_EPI0) * 32;
constexpr int B3_CNT_STRIDE = 8;          // one counter per 32-byte sector
constexpr int B3_RCH = 32;                // head: weight rows per streamed chunk
constexpr int B3_NST = 4;                 // head: chunks in flight
constexpr int B3_HIST = 32;

struct BCta {
    int32_t role, layer, hindex;          // hindex: group index of a layer CTA / head CTA index
    int32_t dil;
    int32_t pair0, npairs, skip0, nskip;
    int32_t rows, RP;                     // GEMM rows (2*K*np + nskip) and their padding to 8
    int32_t ring_len, ring_in_smem;
    int32_t nprod_in;                     // producers of the input tile (counter increments per iteration)
    int32_t is_last;
    int32_t sm_w, sm_x, sm_o, sm_pre, sm_hn, sm_wl, sm_ring, sm_bar;
    int64_t w_off;                        // [Hp][RP]
    int64_t b_off;                        // [np][2]
    int64_t wl_off;                       // last layer: [np][S] columns of out_skip[L-1]
    int64_t ring_off;                     // [RL][2*np][Bp] in ring_scratch
    // head CTA shared-memory layout
    int32_t sm_wring, sm_x1, sm_x2, sm_oc, sm_tin, sm_hist, sm_lpb, sm_misc;
};

struct BatchParams {
    int32_t kind, H, Hp, S, Sp, K, L, lpc, n1, n1p, n2, n2p, Q, U, R, B, M, Tf;
    int32_t t_first, it0, n_it, n_iter;
    int32_t NH, G;                        // head CTAs, CTAs per layer group
    int64_t n_out;
    const BCta* desc;
    const float* wbuf;
    const float* P;                       // [B][Tf][L][SEG][2H]
    const float* pconst;                  // [L][2H]
    const float* gain;                    // [U]
    const float* in_tab;
    const float* in_bias;
    int64_t w1_off, b1_off, w2_off, b2_off, cs_off;    // head weights inside wbuf
    float* tiles;                         // [L+1][M][H][NC]
    unsigned int* cnt;                    // [L+1][M] * B3_CNT_STRIDE
    float* skipbuf;                       // [L-1][M][S][BS]
    float* lastskip;                      // [G][M][S][BS]
    float* ring_scratch;
    const float* noise;
    unsigned long long seed;
    const void* forced;
    const void* seedw;
    void* out;
    float* head_out;
    int32_t* status;
    unsigned long long timeout_cycles;
    unsigned long long* prof;             // optional [prof_count][M][grid][16]
    int32_t prof_start, prof_count;
};

// ------------------------------------------------------------------ primitives
__device__ __forceinline__ uint32_t b3_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void b3_mbar_init(unsigned long long* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b3_smem(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void b3_mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b3_smem(b)) : "memory");
}
__device__ __forceinline__ void b3_mbar_arm(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b3_smem(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool b3_mbar_try(unsigned long long* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(b3_smem(b)), "r"(parity) : "memory");
    return done != 0;
}
// wait with an escape: returns false when another CTA / role aborted
__device__ __forceinline__ bool b3_mbar_wait(unsigned long long* b, uint32_t parity, const int32_t* status) {
    uint32_t spins = 0;
    while (!b3_mbar_try(b, parity)) {
        if (((++spins) & 63u) == 0 && ((volatile const int32_t*)status)[0] != 0) return false;
    }
    return true;
}
// warp-uniform variant: every lane polls (so each observes the completion itself), the verdict is shared
__device__ __forceinline__ bool b3_wait_warp(unsigned long long* b, uint32_t parity, const int32_t* status) {
    const bool ok = b3_mbar_wait(b, parity, status);
    return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ void b3_tma_load(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(b3_smem(smem_dst)), "l"(gsrc), "r"(bytes), "r"(b3_smem(bar)) : "memory");
}
__device__ __forceinline__ void b3_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Spin until the monotonic counter reached `target` (acquire); false = watchdog / abort.
static __device__ __noinline__ bool b3_cnt_wait(const unsigned int* p, unsigned int target, const BatchParams& P, int where) {
    const long long t_start = clock64();
    uint32_t spins = 0;
    for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v >= target) return true;
        if (((++spins) & 63u) == 0) {
            volatile int32_t* st = P.status;
            if (st[0] != 0) return false;
            if ((unsigned long long)(clock64() - t_start) > P.timeout_cycles) {
                if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)target; st[3] = blockIdx.x; }
                return false;
            }
        }
    }
}
__device__ __forceinline__ void b3_signal(unsigned int* p) {
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

#define B3_PROF(ev)                                                                                      \
    do {                                                                                                 \
        if (P.prof && it >= P.prof_start && it < P.prof_start + P.prof_count)                            \
            P.prof[((((size_t)(it - P.prof_start) * P.M + m) * gridDim.x + blockIdx.x) << 4) + (ev)] = gtime(); \
    } while (0)

__host__ __device__ constexpr int b3_pick_tc(int nc) {
    return (nc % 5 == 0) ? 5 : (nc % 4 == 0) ? 4 : (nc % 3 == 0) ? 3 : (nc % 2 == 0) ? 2 : 1;
}

// ------------------------------------------------------------------ layer CTA
template <int SEG, int BS>
__device__ void b3_layer(const BatchParams& P, const BCta& D, float* smem) {
    constexpr int NC = SEG * BS;
    constexpr int TC = b3_pick_tc(NC);
    constexpr int NCG = NC / TC;                    // column groups (NCG*TC == NC)
    static_assert(NCG <= 32, "too many column groups");
    constexpr int RGW = 32 / NCG;                   // row groups handled by one warp pass
    constexpr int TR = (NC >= 40) ? 8 : 4;
    constexpr int RPW = RGW * TR;                   // rows per warp pass
    constexpr int OS = NC | 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = D.layer, np = D.npairs, K = P.K, H = P.H, Hp = P.Hp, dil = D.dil, RL = D.ring_len, RP = D.RP;
    const int M = P.M, Bp = M * BS, np2 = 2 * np;

    float* Wt = smem + D.sm_w;
    float* Xs = smem + D.sm_x;
    float* Os = smem + D.sm_o;
    float4* pre = reinterpret_cast<float4*>(smem + D.sm_pre);
    float* hn = smem + D.sm_hn;
    float* Wl = smem + D.sm_wl;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + D.sm_bar);
    unsigned long long* x_full = bars;              // [2]
    unsigned long long* x_empty = bars + 2;         // [2]
    unsigned long long* o_full = bars + 4;
    unsigned long long* o_empty = bars + 5;
    unsigned long long* ebar = bars + 6;            // barrier among the epilogue warps (abortable, unlike bar.sync)

    const int nrb = (RP + RPW - 1) / RPW;
    const int ngwa = nrb < B3_NGW ? nrb : B3_NGW;   // GEMM warps that own rows

    // ---- one-time staging
    {
        const float4* src = reinterpret_cast<const float4*>(P.wbuf + D.w_off);
        float4* dst = reinterpret_cast<float4*>(Wt);
        const int n4 = Hp * RP / 4;
        for (int i = tid; i < n4; i += SWN_NTHREADS) dst[i] = src[i];
        for (int i = tid; i < 2 * Hp * NC; i += SWN_NTHREADS) Xs[i] = 0.f;
        if (D.is_last) {
            const float* s2 = P.wbuf + D.wl_off;
            for (int i = tid; i < np * P.S; i += SWN_NTHREADS) Wl[i] = s2[i];
        }
        if (D.ring_in_smem) {
            float* rg = smem + D.sm_ring;
            for (int i = tid; i < RL * np2 * Bp; i += SWN_NTHREADS) rg[i] = 0.f;
        }
        if (tid == 0) {
            b3_mbar_init(&x_full[0], 1); b3_mbar_init(&x_full[1], 1);
            b3_mbar_init(&x_empty[0], ngwa + 1); b3_mbar_init(&x_empty[1], ngwa + 1);
            b3_mbar_init(o_full, ngwa); b3_mbar_init(o_empty, 1);
            b3_mbar_init(ebar, SWN_NWARPS - B3_EPI0);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    b3_fence_proxy_async();         // the zeroed X buffers are overwritten by the async proxy later
    __syncthreads();

    const int nq = P.n_it * M;
    const uint32_t tile_bytes = (uint32_t)(H * NC * sizeof(float));
    const float* tile_in = P.tiles + (size_t)l * M * H * NC;
    float* tile_out = P.tiles + (size_t)(l + 1) * M * H * NC;

    if (warp < B3_NGW) {
        // =============================== GEMM warps
        if (warp >= ngwa) return;
        const int rg = lane / NCG, cg = lane - rg * NCG;
        const bool lane_ok = rg < RGW;
        for (int q = 0; q < nq; ++q) {
            const int buf = q & 1;
            if (!b3_wait_warp(&x_full[buf], (q >> 1) & 1, P.status)) return;
            const float* X = Xs + buf * Hp * NC;
            bool first = true;
            for (int rb = warp; rb < nrb; rb += ngwa) {
                const int r0 = (rb * RGW + rg) * TR;
                const bool act = lane_ok && r0 < RP;
                float acc[TR][TC];
#pragma unroll
                for (int i = 0; i < TR; ++i)
#pragma unroll
                    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;
                if (act) {
                    const float* wp = Wt + r0;
                    const float* xp = X + cg * TC;
#pragma unroll 4
                    for (int k = 0; k < Hp; ++k) {
                        float w[TR], x[TC];
#pragma unroll
                        for (int i = 0; i < TR; i += 4) {
                            const float4 v = *reinterpret_cast<const float4*>(wp + (size_t)k * RP + i);
                            w[i] = v.x; w[i + 1] = v.y; w[i + 2] = v.z; w[i + 3] = v.w;
                        }
                        if constexpr (TC == 4) {
                            const float4 v = *reinterpret_cast<const float4*>(xp + k * NC);
                            x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
                        } else if constexpr (TC == 2) {
                            const float2 v = *reinterpret_cast<const float2*>(xp + k * NC);
                            x[0] = v.x; x[1] = v.y;
                        } else {
#pragma unroll
                            for (int j = 0; j < TC; ++j) x[j] = xp[k * NC + j];
                        }
#pragma unroll
                        for (int i = 0; i < TR; ++i)
#pragma unroll
                            for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(w[i], x[j], acc[i][j]);
                    }
                }
                if (first) {
                    first = false;
                    if (q > 0 && !b3_wait_warp(o_empty, (q - 1) & 1, P.status)) return;
                }
                if (act) {
#pragma unroll
                    for (int i = 0; i < TR; ++i)
#pragma unroll
                        for (int j = 0; j < TC; ++j) Os[(r0 + i) * OS + cg * TC + j] = acc[i][j];
                }
            }
            __syncwarp();
            if (lane == 0) { b3_mbar_arrive(o_full); b3_mbar_arrive(&x_empty[buf]); }
        }
        return;
    }

    if (warp == B3_LOADER) {
        // =============================== loader
        if (lane != 0) return;
        for (int q = 0; q < nq; ++q) {
            const int it = q / M, m = q - it * M, buf = q & 1;
            if (q >= 2 && !b3_mbar_wait(&x_empty[buf], ((q >> 1) - 1) & 1, P.status)) return;
            if (!b3_cnt_wait(P.cnt + ((size_t)l * M + m) * B3_CNT_STRIDE, (unsigned)D.nprod_in * (unsigned)(it + 1), P, 100 + l)) return;
            B3_PROF(0);
            b3_fence_proxy_async();
            b3_mbar_arm(&x_full[buf], tile_bytes);
            b3_tma_load(Xs + buf * Hp * NC, tile_in + (size_t)m * H * NC, tile_bytes, &x_full[buf]);
        }
        return;
    }

    // =================================== epilogue warps
    const int te = tid - B3_EPI0 * 32;
    const float* bias = P.wbuf + D.b_off;
    float* ringS = smem + D.sm_ring;
    float* ringG = P.ring_scratch + D.ring_off;
    const int twoH = 2 * H;
    const int nitems = np * NC;
    uint32_t ep = 0;
    auto epi_sync = [&]() -> bool {
        __syncwarp();
        if (lane == 0) b3_mbar_arrive(ebar);
        const bool ok = b3_wait_warp(ebar, ep & 1, P.status);
        ++ep;
        return ok;
    };
    for (int q = 0; q < nq; ++q) {
        const int it = q / M, m = q - it * M, buf = q & 1;
        const int t0 = P.t_first + it * SEG;
        const bool head_active = it >= P.it0;
        // ---- P1: everything the gate needs that does not depend on the GEMM
        for (int item = te; item < nitems; item += B3_NEPI) {
            const int pi = item / NC, col = item - pi * NC;
            const int s = col / BS, b = col - s * BS;
            const int t = t0 + s;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0) {
                int bglob = m * BS + b; if (bglob >= P.B) bglob = P.B - 1;
                const int i = D.pair0 + pi;
                float az = __ldg(P.pconst + l * twoH + i), ac = __ldg(P.pconst + l * twoH + H + i);
#pragma unroll
                for (int sp = 0; sp < SEG; ++sp) {
                    const int tau = t + sp;
                    int f = 0, j = 0;
                    if (tau >= P.R) { f = (tau - P.R) / P.U; j = (tau - P.R) - f * P.U; }
                    const float g = __ldg(P.gain + j);
                    const float* pr = P.P + ((((size_t)bglob * P.Tf + f) * P.L + l) * SEG + sp) * twoH;
                    az = fmaf(g, __ldg(pr + i), az);
                    ac = fmaf(g, __ldg(pr + H + i), ac);
                }
                const int slot = (t - P.t_first) % RL;
                const size_t ro = ((size_t)slot * np2 + pi * 2) * Bp + m * BS + b;
                float rz, rc;
                if (D.ring_in_smem) { rz = ringS[ro]; rc = ringS[ro + Bp]; ringS[ro] = 0.f; ringS[ro + Bp] = 0.f; }
                else { rz = __ldcg(ringG + ro); rc = __ldcg(ringG + ro + Bp); __stcg(ringG + ro, 0.f); __stcg(ringG + ro + Bp, 0.f); }
                v = make_float4(az, ac, rz + bias[pi * 2], rc + bias[pi * 2 + 1]);
            }
            pre[item] = v;
        }
        if (te == 0) B3_PROF(3);
        if (!b3_wait_warp(&x_full[buf], (q >> 1) & 1, P.status)) return;
        if (!b3_wait_warp(o_full, q & 1, P.status)) return;
        if (te == 0) B3_PROF(4);
        const float* X = Xs + buf * Hp * NC;
        float* tout = tile_out + (size_t)m * H * NC;
        // ---- P2: gate (cswnv_shift1.py:281-285) and publish
        for (int item = te; item < nitems; item += B3_NEPI) {
            const int pi = item / NC, col = item - pi * NC;
            const int s = col / BS, b = col - s * BS;
            const int t = t0 + s;
            float hv = 0.f;
            if (t >= 0) {
                const float4 v = pre[item];
                float qz = v.z, qc = v.w;
                for (int j = K - 1; j >= 0; --j) {
                    const int delta = (K - 1 - j) * dil;
                    if (delta > s) break;
                    const float* o = Os + ((j * np + pi) * 2) * OS + (s - delta) * BS + b;
                    qz += o[0]; qc += o[OS];
                }
                const float z = sigmoidf_(qz * v.x);
                hv = (1.f - z) * tanhf(qc * v.y) + z * X[(D.pair0 + pi) * NC + col];
            }
            __stcg(tout + (size_t)(D.pair0 + pi) * NC + col, hv);
            if (D.is_last && s == SEG - 1) hn[pi * BS + b] = hv;
        }
        // rows of out_skip[l-1] applied to the last time column of every utterance
        if (l >= 1 && head_active) {
            float* sk = P.skipbuf + ((size_t)(l - 1) * M + m) * P.S * BS;
            for (int e = te; e < D.nskip * BS; e += B3_NEPI) {
                const int r = e / BS, b = e - r * BS;
                __stcg(sk + (size_t)(D.skip0 + r) * BS + b, Os[(K * np2 + r) * OS + (SEG - 1) * BS + b]);
            }
        }
        if (D.is_last) {
            if (!epi_sync()) return;
            if (head_active) {
                // out_skip[L-1] restricted to the rows of h this CTA owns (partial sums over k)
                float* ls = P.lastskip + ((size_t)D.hindex * M + m) * P.S * BS;
                for (int e = te; e < P.S * BS; e += B3_NEPI) {
                    const int r = e / BS, b = e - r * BS;
                    float a = 0.f;
                    for (int pi = 0; pi < np; ++pi) a = fmaf(Wl[pi * P.S + r], hn[pi * BS + b], a);
                    __stcg(ls + e, a);
                }
            }
        }
        if (!epi_sync()) return;
        if (te == 0) { b3_signal(P.cnt + ((size_t)(l + 1) * M + m) * B3_CNT_STRIDE); B3_PROF(5); }
        // ---- P3 (off the critical path): delayed taps into the ring of partial sums
        if (dil >= SEG) {
            // every (tap, column) has its own target slot
            const int n3 = (K - 1) * SEG * np2 * BS;
            for (int e = te; e < n3; e += B3_NEPI) {
                const int b = e % BS;
                int r = e / BS;
                const int r2 = r % np2; r /= np2;
                const int s = r % SEG;
                const int j = r / SEG;                                   // tap 0..K-2
                const int delta = (K - 1 - j) * dil;
                const int slot = (t0 + s + delta - P.t_first) % RL;
                const float val = Os[(j * np2 + r2) * OS + s * BS + b];
                const size_t ro = ((size_t)slot * np2 + r2) * Bp + m * BS + b;
                if (D.ring_in_smem) ringS[ro] += val;
                else atomicAdd(ringG + ro, val);                         // RED: fire and forget, one writer per element
            }
        } else {
            // several (tap, column) pairs share a target: one thread sums them in a fixed order
            const int ntau = (K - 1) * dil;
            const int n3 = ntau * np2 * BS;
            for (int e = te; e < n3; e += B3_NEPI) {
                const int b = e % BS;
                int r = e / BS;
                const int r2 = r % np2;
                const int trel = SEG + r / np2;
                float val = 0.f;
                for (int j = K - 2; j >= 0; --j) {
                    const int s = trel - (K - 1 - j) * dil;
                    if (s < 0) break;
                    if (s < SEG) val += Os[(j * np2 + r2) * OS + s * BS + b];
                }
                const int slot = (t0 + trel - P.t_first) % RL;
                const size_t ro = ((size_t)slot * np2 + r2) * Bp + m * BS + b;
                if (D.ring_in_smem) ringS[ro] += val;
                else atomicAdd(ringG + ro, val);
            }
        }
        if (!epi_sync()) return;
        if (te == 0) { b3_mbar_arrive(o_empty); b3_mbar_arrive(&x_empty[buf]); B3_PROF(6); }
    }
}

// ------------------------------------------------------------------ head CTA
// sums of N accumulators over the 32 lanes with N-1 + log2(32/N) shuffles; returns the total of
// accumulator `which` (a function of the lane), valid in every lane
template <int N>
__device__ __forceinline__ float b3_reduce_t(float (&v)[N], int lane, int& which) {
    which = 0;
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = hi ? v[i] : v[i + n / 2];
            const float keep = hi ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
        which += hi ? n / 2 : 0;
        off >>= 1;
    }
    float r = v[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// out[b][row] = act(W[row][:] . x[b][:] + bias[row]) for the rows of one streamed chunk.
// Lane owns k = c*64 + 2*lane, +1; the CB columns of this warp's group are register resident.
template <int CB, bool RELU>
__device__ __forceinline__ void b3_head_rows(const float* __restrict__ wch, int nrows, int kp, const float2 (&xr)[4][CB],
                                             const float* __restrict__ bias, int row0, float* __restrict__ out, int ostride,
                                             int b0, int wr, int wstep, int lane) {
    const int nchk = kp >> 6;
    const int lowmask = (32 / CB) - 1;
    for (int r = wr; r < nrows; r += wstep) {
        const float2* w2 = reinterpret_cast<const float2*>(wch + (size_t)r * kp);
        float acc[CB];
#pragma unroll
        for (int b = 0; b < CB; ++b) acc[b] = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c < nchk) {
                const float2 w = w2[c * 32 + lane];
#pragma unroll
                for (int b = 0; b < CB; ++b) acc[b] = fmaf(w.x, xr[c][b].x, fmaf(w.y, xr[c][b].y, acc[b]));
            }
        }
        int which;
        float v = b3_reduce_t<CB>(acc, lane, which);
        if ((lane & lowmask) == 0) {
            v += __ldg(bias + row0 + r);
            if (RELU) v = fmaxf(v, 0.f);
            out[(size_t)(b0 + which) * ostride + row0 + r] = v;
        }
    }
}

template <int SEG, int BS>
__device__ void b3_head(const BatchParams& P, const BCta& D, float* smem) {
    constexpr int NC = SEG * BS;
    constexpr int CB = BS > 8 ? 8 : BS;             // columns per warp group
    constexpr int NGRP = BS / CB;
    static_assert(BS % CB == 0 && (CB & (CB - 1)) == 0, "batch tile must be a power of two");
    constexpr int WPG = SWN_NWARPS / NGRP;          // warps per column group
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = P.H, S = P.S, Sp = P.Sp, K = P.K, L = P.L, n1 = P.n1, n1p = P.n1p, n2 = P.n2, n2p = P.n2p, M = P.M;
    const int hc = D.hindex;
    const bool is_c = P.kind == 0;

    float* wring = smem + D.sm_wring;               // [NST][RCH * kpmax]
    float* x1 = smem + D.sm_x1;                     // [BS][Sp]
    float* x2 = smem + D.sm_x2;                     // [BS][n1p]
    float* oc = smem + D.sm_oc;                     // [BS][n2p]
    float* tin = smem + D.sm_tin;                   // [H][NC]
    float* hist = smem + D.sm_hist;                 // [MO][BS][HIST]
    float* lpb = smem + D.sm_lpb;                   // [MO][BS][8]
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + D.sm_misc);
    unsigned long long* w_full = bars;              // [NST]
    unsigned long long* w_empty = bars + B3_NST;    // [NST]
    int* flag = reinterpret_cast<int*>(bars + 2 * B3_NST);

    const int kpmax = Sp > n1p ? Sp : n1p;
    const int chunk_floats = B3_RCH * kpmax;
    const int nc1 = (n1 + B3_RCH - 1) / B3_RCH, nc2 = (n2 + B3_RCH - 1) / B3_RCH, nct = nc1 + nc2;
    const int nown = hc < M ? (M - hc + P.NH - 1) / P.NH : 0;       // slots served: m = hc + i*NH
    const long long total_chunks = (long long)nown * P.n_iter * nct;

    for (int i = tid; i < BS * Sp; i += SWN_NTHREADS) x1[i] = 0.f;
    for (int i = tid; i < BS * n1p; i += SWN_NTHREADS) x2[i] = 0.f;
    for (int i = tid; i < nown * BS * B3_HIST; i += SWN_NTHREADS) hist[i] = is_c ? 0.f : __int_as_float(P.Q / 2);
    for (int i = tid; i < nown * BS * 8; i += SWN_NTHREADS) lpb[i] = 0.f;
    if (tid == 0) {
        for (int i = 0; i < B3_NST; ++i) { b3_mbar_init(&w_full[i], 1); b3_mbar_init(&w_empty[i], SWN_NWARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (P.seedw) {
        for (int e = tid; e < nown * BS * SEG; e += SWN_NTHREADS) {
            const int mb = e / SEG, j = e - mb * SEG;
            const int ml = mb / BS, b = mb - ml * BS;
            int bglob = (hc + ml * P.NH) * BS + b; if (bglob >= P.B) bglob = P.B - 1;
            const float v = is_c ? ((const float*)P.seedw)[(size_t)bglob * SEG + j]
                                 : __int_as_float((int)(((const long long*)P.seedw)[bglob] % P.Q));
            hist[(size_t)mb * B3_HIST + ((j - SEG + B3_HIST) % B3_HIST)] = v;
            const int mm = P.lpc - SEG + j;
            if (is_c && mm >= 0) lpb[(size_t)mb * 8 + mm] = v;
        }
        __syncthreads();
    }

    // ---- weight stream (out_1 rows, then out_2 rows, cyclically): chunk sequence numbers
    long long pseq = 0, cseq = 0;
    auto chunk_src = [&](int ci, const float*& src, uint32_t& bytes) {
        if (ci < nc1) {
            const int r0 = ci * B3_RCH, nr = min(B3_RCH, n1 - r0);
            src = P.wbuf + P.w1_off + (size_t)r0 * Sp; bytes = (uint32_t)(nr * Sp * sizeof(float));
        } else {
            const int r0 = (ci - nc1) * B3_RCH, nr = min(B3_RCH, n2 - r0);
            src = P.wbuf + P.w2_off + (size_t)r0 * n1p; bytes = (uint32_t)(nr * n1p * sizeof(float));
        }
    };
    auto produce = [&]() -> bool {             // thread 0 only: keep up to NST chunks in flight
        while (pseq < total_chunks && pseq < cseq + B3_NST) {
            const int st = (int)(pseq % B3_NST);
            if (pseq >= B3_NST && !b3_mbar_wait(&w_empty[st], (uint32_t)((pseq / B3_NST) - 1) & 1, P.status)) return false;
            const float* src; uint32_t bytes;
            chunk_src((int)(pseq % nct), src, bytes);
            b3_mbar_arm(&w_full[st], bytes);
            b3_tma_load(wring + (size_t)st * chunk_floats, src, bytes, &w_full[st]);
            ++pseq;
        }
        return true;
    };
    if (tid == 0) produce();

    // Input stage for iteration `itn` of slot m: h0 = softsign(causal(...)) (cswnv_shift1.py:352 / dswnv.py:345)
    auto input_stage = [&](int itn, int m, int ml) {
        const int t0n = P.t_first + itn * SEG;
        const float* hs = hist + (size_t)ml * BS * B3_HIST;
        for (int e = tid; e < H * NC; e += SWN_NTHREADS) {
            const int c = e / H, r = e - c * H;
            const int s = c / BS, b = c - s * BS;
            const int t = t0n + s;
            float v = 0.f;
            if (t >= 0) {
                float acc = __ldg(P.in_bias + r);
                for (int j = 0; j < K; ++j) {
                    const int tau = t - (K - 1) + j;
                    if (tau < 0) continue;
                    const int n = tau - P.R - 1;
                    if (is_c) {
                        const float a = (n >= -SEG) ? hs[b * B3_HIST + ((n + B3_HIST) % B3_HIST)] : 0.f;
                        acc += __ldg(P.in_tab + (j * 2) * H + r) + __ldg(P.in_tab + (j * 2 + 1) * H + r) * a;
                    } else {
                        const int qq = (n >= -SEG) ? __float_as_int(hs[b * B3_HIST + ((n + B3_HIST) % B3_HIST)]) : P.Q / 2;
                        acc += __ldg(P.in_tab + ((size_t)j * P.Q + qq) * H + r);
                    }
                }
                v = softsignf_(acc);
            }
            tin[r * NC + c] = v;
        }
        __syncthreads();
        float* dst = P.tiles + (size_t)m * H * NC;                   // layer 0 input
        for (int e = tid; e < H * NC; e += SWN_NTHREADS) __stcg(dst + e, tin[e]);
        __syncthreads();
        if (tid == 0) b3_signal(P.cnt + (size_t)m * B3_CNT_STRIDE);
    };

    for (int i = 0; i < nown; ++i) input_stage(0, hc + i * P.NH, i);

    const int wg = warp / WPG, wr = warp - wg * WPG;                  // column group and row phase of this warp
    const int b0 = wg * CB;
    for (int it = 0; it < P.n_it; ++it) {
        const bool active = it >= P.it0;
        const int istep = it - P.it0;
        for (int ml = 0; ml < nown; ++ml) {
            const int m = hc + ml * P.NH;
            // ---- wait for the last layer (flow control during warm-up, data when active)
            if (tid == 0)
                *flag = b3_cnt_wait(P.cnt + ((size_t)L * M + m) * B3_CNT_STRIDE, (unsigned)P.G * (unsigned)(it + 1), P, 200) ? 1 : 0;
            __syncthreads();
            if (*flag == 0) return;
            if (tid == 0) B3_PROF(0);
            if (active) {
                // ---- relu(sum_l skip_l) (cswnv_shift1.py:367 inner sum), fixed summation order
                for (int e = tid; e < S * BS; e += SWN_NTHREADS) {
                    const int r = e / BS, b = e - r * BS;
                    float a = __ldg(P.wbuf + P.cs_off + r);
                    for (int l2 = 0; l2 + 1 < L; ++l2) a += __ldcg(P.skipbuf + ((size_t)l2 * M + m) * S * BS + e);
                    for (int g = 0; g < P.G; ++g) a += __ldcg(P.lastskip + ((size_t)g * M + m) * S * BS + e);
                    x1[b * Sp + r] = fmaxf(a, 0.f);
                }
                __syncthreads();
                if (tid == 0) B3_PROF(1);
                bool ok = true;
                // ---- stage B: relu(out_1(.)), stage C: out_2(.)
                for (int stg = 0; stg < 2; ++stg) {
                    const int kp = stg == 0 ? Sp : n1p;
                    const float* xin = stg == 0 ? x1 : x2;
                    float2 xr[4][CB];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
#pragma unroll
                        for (int b = 0; b < CB; ++b)
                            xr[c][b] = (c < (kp >> 6)) ? reinterpret_cast<const float2*>(xin + (size_t)(b0 + b) * kp)[c * 32 + lane]
                                                       : make_float2(0.f, 0.f);
                    const int ncs = stg == 0 ? nc1 : nc2;
                    for (int ci = 0; ci < ncs; ++ci) {
                        if (tid == 0) ok = produce();
                        const int st = (int)(cseq % B3_NST);
                        ok = b3_mbar_wait(&w_full[st], (uint32_t)(cseq / B3_NST) & 1, P.status) && ok;
                        if (!ok) break;
                        const float* wch = wring + (size_t)st * chunk_floats;
                        const int r0 = ci * B3_RCH;
                        if (stg == 0)
                            b3_head_rows<CB, true>(wch, min(B3_RCH, n1 - r0), kp, xr, P.wbuf + P.b1_off, r0, x2, n1p, b0, wr, WPG, lane);
                        else
                            b3_head_rows<CB, false>(wch, min(B3_RCH, n2 - r0), kp, xr, P.wbuf + P.b2_off, r0, oc, n2p, b0, wr, WPG, lane);
                        __syncwarp();
                        if (lane == 0) b3_mbar_arrive(&w_empty[st]);
                        ++cseq;
                    }
                    if (__syncthreads_or(!ok)) return;
                }
                if (tid == 0) B3_PROF(2);
                float* hs = hist + (size_t)ml * BS * B3_HIST;
                if (is_c) {
                    if (tid < BS) {
                        const int b = tid;
                        int bglob = m * BS + b;
                        const bool real = bglob < P.B;
                        if (!real) bglob = P.B - 1;
                        const long long nbase = (long long)istep * SEG;
                        float eps[SEG], smp[SEG];
#pragma unroll
                        for (int j = 0; j < SEG; ++j)
                            eps[j] = P.noise ? P.noise[(size_t)bglob * P.n_out + nbase + j] : laplace_eps(P.seed, bglob, nbase + j);
                        const float* forced = P.forced ? (const float*)P.forced + (size_t)bglob * P.n_out + nbase : nullptr;
                        laplace_step<SEG>(P.lpc, oc + b * n2p, lpb + ((size_t)ml * BS + b) * 8, eps, forced, smp);
#pragma unroll
                        for (int j = 0; j < SEG; ++j) {
                            const float fed = forced ? forced[j] : smp[j];
                            hs[b * B3_HIST + (int)((nbase + j) % B3_HIST)] = fed;
                            if (real) ((float*)P.out)[(size_t)bglob * P.n_out + nbase + j] = smp[j];
                        }
                    }
                    if (P.head_out) {
                        for (int e = tid; e < BS * n2; e += SWN_NTHREADS) {
                            const int b = e / n2, r = e - b * n2;
                            const int bglob = m * BS + b;
                            if (bglob < P.B) P.head_out[((size_t)bglob * P.n_iter + istep) * n2 + r] = oc[b * n2p + r];
                        }
                    }
                } else {
                    // Gumbel-max: argmax_k(logit_k - log q_k)  (dswnv.py:363-369)
                    for (int b = warp; b < BS; b += SWN_NWARPS) {
                        int bglob = m * BS + b;
                        const bool real = bglob < P.B;
                        if (!real) bglob = P.B - 1;
                        float best = -CUDART_INF_F; int bi = 0x7fffffff;
                        for (int r = lane; r < n2; r += 32) {
                            const float o = oc[b * n2p + r];
                            if (P.head_out && real) P.head_out[((size_t)bglob * P.n_iter + istep) * n2 + r] = o;
                            const float qv = P.noise ? P.noise[((size_t)bglob * P.n_out + istep) * P.Q + r]
                                                     : softmax_q(P.seed, bglob, istep, r);
                            const float g = o - logf(qv);
                            if (g > best || (g == best && r < bi)) { best = g; bi = r; }
                        }
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) {
                            const float ob = __shfl_xor_sync(0xffffffffu, best, o2);
                            const int oi = __shfl_xor_sync(0xffffffffu, bi, o2);
                            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                        }
                        if (lane == 0) {
                            int fed = bi;
                            if (real) {
                                ((long long*)P.out)[(size_t)bglob * P.n_out + istep] = bi;
                                if (P.forced) fed = (int)((const long long*)P.forced)[(size_t)bglob * P.n_out + istep];
                            }
                            hs[b * B3_HIST + (istep % B3_HIST)] = __int_as_float(fed);
                        }
                    }
                }
                __syncthreads();
                if (tid == 0) B3_PROF(3);
            }
            if (it + 1 < P.n_it) input_stage(it + 1, m, ml);
            if (tid == 0) B3_PROF(4);
        }
    }
}

// ------------------------------------------------------------------ kernel
template <int SEG, int BS>
__global__ void __launch_bounds__(SWN_NTHREADS, 1) swn_batch_kernel(const __grid_constant__ BatchParams P) {
    extern __shared__ __align__(16) float smem[];
    const BCta& D = P.desc[blockIdx.x];
    if (D.role == ROLE_LAYER) b3_layer<SEG, BS>(P, D, smem);
    else if (D.role == ROLE_HEAD) b3_head<SEG, BS>(P, D, smem);
}

}  // namespace swn
