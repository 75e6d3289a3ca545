// Latency path of the generation loop (small batch, one sample per step): the whole dependent
// chain of a time step runs inside ONE 16-CTA thread-block cluster, the rest of the chip helps.
//
// Why: profiles/r01_v1_stage_timeline.txt -- with the layers spread over the chip a vector hop
// through L2 costs 1.0-1.6 us and a time step is a chain of ten of them.  Inside a cluster a hop is
// a distributed-shared-memory store plus a local shared-memory poll.
//
//   core cluster (16 CTAs)   every CTA owns H/16 gate pairs of EVERY layer, S/16 rows of every
//                            out_skip, 1/16 of out_1 / out_2 and of the input-stage table.
//                            Per layer only the taps with delay 0 (and delay 1 for dilation-1
//                            layers) are evaluated here; they live in TENSOR MEMORY (tcgen05.st
//                            once, tcgen05.ld prefetched into registers while the previous hop
//                            is in flight) -- TMEM is used as 256 KB of extra weight-stationary
//                            storage per SM, which is what makes the critical weights of all six
//                            layers fit in one cluster.  Vectors are exchanged as 64-bit
//                            {fp32, tag} words written into every peer's shared memory
//                            (st.shared::cluster) and polled locally.
//   helper CTAs (rest)       own the remaining taps (delay >= 2 steps) in shared memory in the
//                            scatter form of v1; they receive the layer inputs through L2 and
//                            return finished partial sums through L2 with >= 2 steps of slack.
#pragma once
#include "swn_gen.cuh"

namespace swn {

constexpr int CORE_CL = 16;        // CTAs in the core cluster
constexpr int CORE_KMAX = 8;       // max kernel_size handled by the helper inner loop

struct CoreCta {                   // one per core CTA (global memory)
    int32_t pair0, npairs;         // gate pairs of every layer (npairs <= 16: one per warp)
    int32_t skip0, nskip;          // rows of every out_skip
    int32_t b0, nb;                // rows of out_1
    int32_t c0, nc;                // rows of out_2 (CSWNV: all rows)
    int32_t i0, ni;                // rows of the input stage
    int64_t tap_off;               // [n_core_taps][npairs][2][Hp]
    int64_t bias_off;              // [L][npairs][2]
    int64_t skip_off;              // [L][nskip][Hp]
    int64_t skipc_off;             // [nskip] sum of all skip biases
    int64_t w1_off, b1_off;        // [nb][Sp], [nb]
    int64_t w2_off, b2_off;        // [nc][n1p], [nc]
    int64_t tin_off;               // DSWNV: [Q][ni] newest-tap table slice
    int32_t sm_skip, sm_w1, sm_w2, sm_tin, sm_ring, sm_hx, sm_x1, sm_x2, sm_gx, sm_misc, sm_pre, sm_tmem_slot;
};

struct HelpCta {                   // one per helper CTA
    int32_t layer, dil, pair0, npairs, ntaps, ring_len, dmin;   // dmin = smallest helper delay
    int32_t sm_w, sm_h, sm_ring, ring_in_smem;
    int64_t ring_off;              // float offset into ring_scratch when the ring does not fit in shared memory
    int64_t w_off;                 // [ntaps][npairs][2][Hp], taps ordered by increasing delay
    int32_t delay[CORE_KMAX];
};

struct CoreParams {
    int32_t kind, H, Hp, S, Sp, K, L, lpc, n1, n1p, n2, Q, U, R, B, Tf;
    int32_t it0, n_it, n_iter, n_core_taps, n_help, audio_in;
    int64_t n_out;
    int32_t dil[16];
    int32_t tap_base[16];          // first core tap index of layer l (tap 0 = delay 0, tap 1 = delay 1)
    int32_t has_help[16];
    int32_t hslots[16];            // depth of the core->helper input ring of layer l (iterations)
    int32_t hacc_nt[16];           // time slots of the helper->core partial-sum ring of layer l
    int32_t hdmin[16];             // smallest helper delay of layer l (no helper sum exists before that time)
    int64_t hin_off[16];           // word offsets into hin_ll / hacc_ll
    int64_t hacc_off[16];
    const CoreCta* cdesc;
    const HelpCta* hdesc;
    const float* wbuf;
    const float* P; const float* pconst; const float* gain;
    const float* in_tab; const float* in_bias; const float* xa_tab;
    unsigned long long* hin_ll;    // [l][slot][NC][H]    layer inputs for the helpers
    unsigned long long* hacc_ll;   // [l][tslot][2H][NC]  finished partial sums for the core
    float* ring_scratch;
    const float* noise; unsigned long long seed;
    const void* forced; const void* seedw; void* out; float* head_out;
    int32_t* status; unsigned long long timeout_cycles;
    unsigned long long* prof; int32_t prof_start, prof_count;
};

// ------------------------------------------------------------------ cluster / DSMEM / TMEM primitives
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_ll(uint32_t remote_addr, float val, uint32_t tag) {
    unsigned long long v = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(val);
    asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(remote_addr), "l"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_ll_u32(uint32_t remote_addr, uint32_t val, uint32_t tag) {
    unsigned long long v = ((unsigned long long)tag << 32) | (unsigned long long)val;
    asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(remote_addr), "l"(v) : "memory");
}
__device__ __forceinline__ ulonglong2 ld_smem_ll2(const unsigned long long* p) {    // two adjacent LL words
    ulonglong2 v;
    asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_smem_ll(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}

__device__ __forceinline__ void tmem_alloc_all(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free_all(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld4(uint32_t a, float* r) {
    uint32_t x0, x1, x2, x3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(a));
    r[0] = __uint_as_float(x0); r[1] = __uint_as_float(x1); r[2] = __uint_as_float(x2); r[3] = __uint_as_float(x3);
}
__device__ __forceinline__ void tmem_st4(uint32_t a, const float* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(__float_as_uint(r[0])),
                 "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])) : "memory");
}
// N floats (multiple of 4) per thread, asynchronous: tmem_wait_ld() before use
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t a, float (&r)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 4) tmem_ld4(a + i, &r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_st(uint32_t a, const float (&r)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 4) tmem_st4(a + i, &r[i]);
}

#define SWN_CPROF(ev)                                                                              \
    do {                                                                                           \
        if (P.prof && threadIdx.x == 0 && it >= P.prof_start && it < P.prof_start + P.prof_count) \
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 8 + (ev)] = gtime();   \
    } while (0)

// Wait until the lane's slice of an LL vector in LOCAL shared memory carries `tag`, return it in
// registers (lane owns k = c*64 + 2*lane, +1).  Polling local SMEM costs no off-SM traffic.
template <int NCH>
__device__ __forceinline__ bool wait_vec(const unsigned long long* vec, uint32_t tag, float2 (&h)[NCH], int lane,
                                         const CoreParams& P, int where) {
    uint32_t spins = 0;
    long long t_start = 0;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        const unsigned long long* p = vec + c * 64 + 2 * lane;
        const int k0 = c * 64 + 2 * lane;
        const bool pad0 = k0 >= P.H, pad1 = k0 + 1 >= P.H;      // padded words are never written (value 0)
        for (;;) {
            const ulonglong2 v = ld_smem_ll2(p);
            if ((pad0 || (uint32_t)(v.x >> 32) == tag) && (pad1 || (uint32_t)(v.y >> 32) == tag)) {
                h[c].x = __uint_as_float((uint32_t)v.x); h[c].y = __uint_as_float((uint32_t)v.y);
                break;
            }
            if (((++spins) & 1023u) == 0) {
                volatile int32_t* st = P.status;
                if (t_start == 0) t_start = clock64();
                if (st[0] != 0) return false;
                if ((unsigned long long)(clock64() - t_start) > (P.timeout_cycles >> 1)) {
                    if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)tag; st[3] = blockIdx.x; }
                    return false;
                }
            }
        }
    }
    return true;
}

// Two adjacent LL words in local shared memory (with the watchdog); padded words are not tag-checked.
__device__ __forceinline__ bool wait_pair(const unsigned long long* p, uint32_t tag, bool pad0, bool pad1, ulonglong2& v,
                                          const CoreParams& P, int where) {
    uint32_t spins = 0;
    long long t_start = 0;
    for (;;) {
        v = ld_smem_ll2(p);
        if ((pad0 || (uint32_t)(v.x >> 32) == tag) && (pad1 || (uint32_t)(v.y >> 32) == tag)) return true;
        if (((++spins) & 1023u) == 0) {
            volatile int32_t* st = P.status;
            if (t_start == 0) t_start = clock64();
            if (st[0] != 0) return false;
            if ((unsigned long long)(clock64() - t_start) > (P.timeout_cycles >> 1)) {
                if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)tag; st[3] = blockIdx.x; }
                return false;
            }
        }
    }
}

__device__ __forceinline__ bool wait_gword(const unsigned long long* p, uint32_t tag, uint32_t& payload, const CoreParams& P,
                                           int where) {
    uint32_t spins = 0;
    long long t_start = 0;
    for (;;) {
        const unsigned long long v = ld_ll(p);
        if ((uint32_t)(v >> 32) == tag) { payload = (uint32_t)v; return true; }
        if (((++spins) & 255u) == 0) {
            volatile int32_t* st = P.status;
            if (t_start == 0) t_start = clock64();
            if (st[0] != 0) return false;
            if ((unsigned long long)(clock64() - t_start) > P.timeout_cycles) {
                if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)tag; st[3] = blockIdx.x; }
                return false;
            }
        }
    }
}

// ------------------------------------------------------------------ helper CTA
// Owns the taps of one layer whose contribution is needed >= 2 steps in the future.  Scatter form:
// the input vector of time `it` adds W_j . h into the partial sum of time it + delay_j; when the
// smallest delay has been applied the sum of that target time is complete and is handed to the core.
template <int NCH, int BS>
__device__ void helper_cta(const CoreParams& P, const HelpCta& D, float* smem) {
    constexpr int NC = BS, HP = NCH * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = D.layer, np = D.npairs, H = P.H, RL = D.ring_len, nt = D.ntaps;
    float* sW = smem + D.sm_w;
    {
        const float4* src = reinterpret_cast<const float4*>(P.wbuf + D.w_off);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = tid; i < nt * np * 2 * HP / 4; i += SWN_NTHREADS) dst[i] = src[i];
    }
    float* hb = smem + D.sm_h;                                 // [2][NC][HP]
    for (int i = tid; i < 2 * NC * HP; i += SWN_NTHREADS) hb[i] = 0.f;
    float* ring;                                               // [RL][np][2][BS]
    if (D.ring_in_smem) {
        ring = smem + D.sm_ring;
        for (int i = tid; i < RL * np * 2 * BS; i += SWN_NTHREADS) ring[i] = 0.f;
    } else {
        ring = P.ring_scratch + D.ring_off;                    // zeroed by the host
    }
    __syncthreads();
    const unsigned long long* in_ll = P.hin_ll + P.hin_off[l];
    unsigned long long* acc_ll = P.hacc_ll + P.hacc_off[l];
    const int nslot = P.hslots[l], NT = P.hacc_nt[l];
    // lane <-> (tap j, column c) for the ring update
    const int my_j = lane / NC, my_c = lane - my_j * NC;
    const bool my_live = my_j < nt;
    const int my_delay = my_live ? D.delay[my_j] : 0;
    int hsel = 0;
    for (int it = 0; it < P.n_it; ++it) {
        const uint32_t tag = (uint32_t)it + 1u;
        float* hbuf = hb + hsel * NC * HP;
        hsel ^= 1;
        bool ok = true;
        {
            const unsigned long long* src = in_ll + (size_t)(it % nslot) * NC * H;
            for (int i = tid; i < NC * H; i += SWN_NTHREADS) {
                uint32_t bits;
                ok = wait_gword(src + i, tag, bits, P, 300 + l) && ok;
                const int c = i / H, k = i - c * H;
                hbuf[c * HP + k] = __uint_as_float(bits);
            }
        }
        if (__syncthreads_or(!ok)) return;
        float2 h[NC][NCH];
#pragma unroll
        for (int c = 0; c < NC; ++c) load_frag<NCH>(h[c], hbuf + c * HP, lane);
        for (int pi = warp; pi < np; pi += SWN_NWARPS) {
            // partial dot products of every tap first, then the (independent, pipelined) reductions
            float mz = 0.f, mc = 0.f;
#pragma unroll
            for (int j = 0; j < CORE_KMAX; ++j) {
                if (j < nt) {
                    const float* wr = sW + ((size_t)(j * np + pi) * 2) * HP;
                    float2 wz[NCH], wc[NCH];
                    load_frag<NCH>(wz, wr, lane);
                    load_frag<NCH>(wc, wr + HP, lane);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const float a = warp_sum(dot_frag<NCH>(wz, h[c])), b = warp_sum(dot_frag<NCH>(wc, h[c]));
                        if (lane == j * NC + c) { mz = a; mc = b; }
                    }
                }
            }
            if (my_live) {
                const int T = it + my_delay;
                float* r0 = ring + ((size_t)((T % RL) * np + pi) * 2) * BS + my_c;
                const float a = r0[0] + mz, b = r0[BS] + mc;
                if (my_j == 0) {
                    // smallest delay: the sum for target time T is complete -> hand it to the core, clear the slot
                    r0[0] = 0.f; r0[BS] = 0.f;
                    unsigned long long* dst = acc_ll + ((size_t)(T % NT) * 2 * H) * NC;
                    st_ll(dst + (size_t)(D.pair0 + pi) * NC + my_c, a, (uint32_t)T + 1u);
                    st_ll(dst + (size_t)(H + D.pair0 + pi) * NC + my_c, b, (uint32_t)T + 1u);
                } else {
                    r0[0] = a; r0[BS] = b;
                }
            }
            __syncwarp();
        }
    }
}

// ------------------------------------------------------------------ core CTA
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct PreSlot {            // what a gate lane needs besides the mat-vec, prefetched one step ahead
    float pz, pc;           // conditioning table entries (before gain / constant)
    unsigned long long hz, hc;   // helper partial sums (LL words, tag = time + 1)
};

template <int NCH, int BS>
__device__ void core_cta(const CoreParams& P, const CoreCta& D, float* smem) {
    constexpr int NC = BS, HP = NCH * 64, FPT = 2 * NCH, TCOLS = 2 * FPT;     // TMEM columns per tap per thread
    constexpr int HIST = 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = P.H, S = P.S, Sp = P.Sp, K = P.K, L = P.L, n1 = P.n1, n1p = P.n1p, n2 = P.n2, Q = P.Q;
    const uint32_t rank = cluster_ctarank();
    const bool is_c = P.kind == 0;
    const int np = D.npairs;
    const bool has_pair = warp < np;
    const int nip = (D.ni + 3) & ~3;

    // ---- shared memory carve-up
    float* sSkip = smem + D.sm_skip;                       // [L][nskip][HP]
    float* sW1 = smem + D.sm_w1;                           // [nb][Sp]
    float* sW2 = smem + D.sm_w2;                           // [nc][n1p]
    float* sTin = smem + D.sm_tin;                         // DSWNV [Q][nip]
    float* ring = smem + D.sm_ring;                        // [L][2][16][2][BS]  delay-1 partial sums
    unsigned long long* hx = reinterpret_cast<unsigned long long*>(smem + D.sm_hx);   // [L+1][NC][HP] LL vectors
    unsigned long long* x1 = reinterpret_cast<unsigned long long*>(smem + D.sm_x1);   // [NC][Sp]
    unsigned long long* x2 = reinterpret_cast<unsigned long long*>(smem + D.sm_x2);   // [NC][n1p]
    unsigned long long* gx = reinterpret_cast<unsigned long long*>(smem + D.sm_gx);   // [NC][256] Gumbel scores
    PreSlot* pre = reinterpret_cast<PreSlot*>(smem + D.sm_pre);                        // [2][L][16][NC]
    float* misc = smem + D.sm_misc;
    float* hist = misc;                                    // [NC][HIST] fed-back samples / indices
    float* lpb = hist + NC * HIST;                         // [NC][8]
    float* o2 = lpb + NC * 8;                              // [NC][16] out_2 rows (CSWNV)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + D.sm_tmem_slot);

    auto stage = [&](float* dst, const float* src, int n) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n / 4; i += SWN_NTHREADS) d4[i] = s4[i];
    };
    stage(sSkip, P.wbuf + D.skip_off, L * D.nskip * HP);
    stage(sW1, P.wbuf + D.w1_off, D.nb * Sp);
    stage(sW2, P.wbuf + D.w2_off, D.nc * n1p);
    if (!is_c) stage(sTin, P.wbuf + D.tin_off, Q * nip);
    for (int i = tid; i < L * 2 * 16 * 2 * BS; i += SWN_NTHREADS) ring[i] = 0.f;
    {
        // zero all LL buffers (tags start at 1; padded words stay {0,0} and are never tag-checked)
        const int nw = (L + 1) * NC * HP + NC * Sp + NC * n1p + NC * 256;
        for (int i = tid; i < nw; i += SWN_NTHREADS) hx[i] = 0ull;
        unsigned long long* pz = reinterpret_cast<unsigned long long*>(pre);
        for (int i = tid; i < (int)(2 * L * 16 * NC * sizeof(PreSlot) / 8); i += SWN_NTHREADS) pz[i] = 0ull;
    }
    for (int i = tid; i < NC * (HIST + 8 + 16); i += SWN_NTHREADS) misc[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < NC * HIST; i += SWN_NTHREADS) hist[i] = is_c ? 0.f : __int_as_float(Q / 2);
    if (warp == 0) tmem_alloc_all(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t taddr0 = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * 128);
    // ---- tensor memory <- the core taps of this warp's gate pair (thread-private columns)
    if (has_pair) {
        for (int c = 0; c < P.n_core_taps; ++c) {
            const float* wr = P.wbuf + D.tap_off + ((size_t)(c * np + warp) * 2) * HP;
            float w[TCOLS];
#pragma unroll
            for (int cc = 0; cc < NCH; ++cc) {
                const float2 a = reinterpret_cast<const float2*>(wr)[cc * 32 + lane];
                const float2 b = reinterpret_cast<const float2*>(wr + HP)[cc * 32 + lane];
                w[2 * cc] = a.x; w[2 * cc + 1] = a.y; w[FPT + 2 * cc] = b.x; w[FPT + 2 * cc + 1] = b.y;
            }
            tmem_st<TCOLS>(taddr0 + c * TCOLS, w);
        }
        tmem_wait_st();
    }
    if (P.seedw) {
        for (int e = tid; e < NC; e += SWN_NTHREADS) {
            const int bglob = min(e, P.B - 1);
            const float v = is_c ? ((const float*)P.seedw)[bglob]
                                 : __int_as_float((int)(((const long long*)P.seedw)[bglob] % Q));
            hist[e * HIST + (HIST - 1)] = v;                       // generated-index -1
            if (is_c && P.lpc > 0) lpb[e * 8 + P.lpc - 1] = v;
        }
    }
    __syncthreads();
    cluster_sync_all();                                            // every CTA's buffers are initialised

    const uint32_t peer = (uint32_t)(lane & (CORE_CL - 1));
    const uint32_t hx_peer = mapa(smem_u32(hx), peer), x1_peer = mapa(smem_u32(x1), peer),
                   x2_peer = mapa(smem_u32(x2), peer), gx_peer = mapa(smem_u32(gx), peer);
    const float* bias = P.wbuf + D.bias_off;
    const int twoH = 2 * H;
    const int col = lane;                                          // gate lane <-> column
    const int my_i = D.pair0 + warp;

    // ---- prefetch (cp.async) of the gate operands of (layer l, time tt) into pre[tt&1][l][warp][col]
    auto prefetch_gate = [&](int l, int tt) {
        if (has_pair && col < NC && tt < P.n_it) {
            PreSlot* ps = pre + (((size_t)(tt & 1) * L + l) * 16 + warp) * NC + col;
            const int bglob = min(col, P.B - 1);
            int f = 0;
            if (tt >= P.R) f = (tt - P.R) / P.U;
            const float* pr = P.P + (((size_t)bglob * P.Tf + f) * L + l) * twoH;
            cp_async4(&ps->pz, pr + my_i);
            cp_async4(&ps->pc, pr + H + my_i);
            if (P.has_help[l] && tt >= P.hdmin[l]) {
                const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((size_t)(tt % P.hacc_nt[l]) * twoH) * NC;
                cp_async8(&ps->hz, src + (size_t)my_i * NC + col);
                cp_async8(&ps->hc, src + (size_t)(H + my_i) * NC + col);
            }
        }
        cp_async_commit();
    };

    // ---- input stage: h0 rows of this CTA for time itn, published into every peer's hx[0]
    // cpn[c] = bias + all taps but the newest (known one step ahead); the newest tap is added here.
    float cpn[NC];
    float vnew = 0.f;
    const int in_row = D.i0 + warp;
    const bool has_in = warp < D.ni;
    if (has_in && is_c) vnew = __ldg(P.in_tab + ((K - 1) * 2 + 1) * H + in_row);
    auto input_partial = [&](int itn) {                             // for time itn, taps j < K-1
        if (!has_in) return;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float acc = __ldg(P.in_bias + in_row);
            if (is_c) acc += __ldg(P.in_tab + ((K - 1) * 2) * H + in_row);       // cb of the newest tap (always valid)
            for (int j = 0; j + 1 < K; ++j) {
                const int tau = itn - (K - 1) + j;
                if (tau < 0) continue;                               // conv zero padding
                const int n = tau - P.R - 1;
                if (is_c) {
                    const float a = (n >= -1) ? hist[c * HIST + ((n + HIST) % HIST)] : 0.f;
                    acc += __ldg(P.in_tab + (j * 2) * H + in_row) + __ldg(P.in_tab + (j * 2 + 1) * H + in_row) * a;
                } else {
                    const int q = (n >= -1) ? __float_as_int(hist[c * HIST + ((n + HIST) % HIST)]) : Q / 2;
                    acc += __ldg(P.in_tab + ((size_t)j * Q + q) * H + in_row);
                }
            }
            cpn[c] = acc;
        }
    };
    auto input_publish = [&](int itn) {
        if (!has_in) return;
        const uint32_t tagn = (uint32_t)itn + 1u;
        const int n = itn - P.R - 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float acc = cpn[c];
            if (is_c) {
                const float a = (n >= -1) ? hist[c * HIST + ((n + HIST) % HIST)] : 0.f;
                acc = fmaf(vnew, a, acc);
            } else {
                const int q = (n >= -1) ? __float_as_int(hist[c * HIST + ((n + HIST) % HIST)]) : Q / 2;
                acc += sTin[q * nip + warp];
            }
            const float v = softsignf_(acc);
            if (lane < CORE_CL) st_cluster_ll(hx_peer + (uint32_t)((c * HP + in_row) * 8), v, tagn);
            if (lane == 16 && P.has_help[0])
                st_ll(P.hin_ll + P.hin_off[0] + ((size_t)(itn % P.hslots[0]) * NC + c) * H + in_row, v, tagn);
        }
    };
    for (int l = 0; l < L; ++l) prefetch_gate(l, 0);
    input_partial(0);
    input_publish(0);
    input_partial(1);

    bool ok = true;
    for (int it = 0; it < P.n_it; ++it) {
        const bool active = it >= P.it0;
        const uint32_t tag = (uint32_t)it + 1u;
        const int istep = it - P.it0;
        const int t = it;
        float sk[NC];                                              // running sum of this warp's skip row
#pragma unroll
        for (int c = 0; c < NC; ++c) sk[c] = 0.f;
        float w[TCOLS];
        if (has_pair) tmem_ld<TCOLS>(taddr0 + P.tap_base[0] * TCOLS, w);
        float gq = 1.f;
        if (t >= P.R) gq = __ldg(P.gain + (t - P.R) % P.U); else gq = __ldg(P.gain);
        cp_async_wait_all();                                       // gate operands of this step (issued one step ago)
        __syncwarp();
        SWN_CPROF(0);
        for (int l = 0; l < L; ++l) {
            // ---- gate operands that do not depend on the layer input
            float pz = 0.f, pc = 0.f, hz = 0.f, hcv = 0.f, rz = 0.f, rc = 0.f, bz = 0.f, bc = 0.f;
            if (has_pair && col < NC) {
                const PreSlot ps = pre[(((size_t)(t & 1) * L + l) * 16 + warp) * NC + col];
                pz = fmaf(gq, ps.pz, __ldg(P.pconst + l * twoH + my_i));
                pc = fmaf(gq, ps.pc, __ldg(P.pconst + l * twoH + H + my_i));
                bz = __ldg(bias + (l * np + warp) * 2); bc = __ldg(bias + (l * np + warp) * 2 + 1);
                if (P.has_help[l] && t >= P.hdmin[l]) {
                    uint32_t a = (uint32_t)ps.hz, b = (uint32_t)ps.hc;
                    if ((uint32_t)(ps.hz >> 32) != (uint32_t)t + 1u || (uint32_t)(ps.hc >> 32) != (uint32_t)t + 1u) {
                        // the helper was late when the prefetch was issued: wait for it here
                        const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((size_t)(t % P.hacc_nt[l]) * twoH) * NC;
                        ok = wait_gword(src + (size_t)my_i * NC + col, (uint32_t)t + 1u, a, P, 400 + l) && ok;
                        ok = wait_gword(src + (size_t)(H + my_i) * NC + col, (uint32_t)t + 1u, b, P, 400 + l) && ok;
                    }
                    hz = __uint_as_float(a); hcv = __uint_as_float(b);
                }
                if (P.dil[l] == 1) {
                    const float* r0 = ring + (((size_t)(l * 2 + (t & 1)) * 16 + warp) * 2) * BS + col;
                    rz = r0[0]; rc = r0[BS];
                }
            }
            // ---- wait for the layer input in local shared memory
            float2 h[NC][NCH];
#pragma unroll
            for (int c = 0; c < NC; ++c) ok = wait_vec<NCH>(hx + ((size_t)l * NC + c) * HP, tag, h[c], lane, P, 500 + l) && ok;
            if (!__all_sync(0xffffffffu, ok)) { ok = false; break; }
            if (has_pair) {
                tmem_wait_ld();
                float myz = 0.f, myc = 0.f;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float az = 0.f, ac = 0.f, bz2 = 0.f, bc2 = 0.f;
#pragma unroll
                    for (int cc = 0; cc < NCH; ++cc) {
                        az = fmaf(w[2 * cc], h[c][cc].x, az); bz2 = fmaf(w[2 * cc + 1], h[c][cc].y, bz2);
                        ac = fmaf(w[FPT + 2 * cc], h[c][cc].x, ac); bc2 = fmaf(w[FPT + 2 * cc + 1], h[c][cc].y, bc2);
                    }
                    const float vz = warp_sum(az + bz2), vc = warp_sum(ac + bc2);
                    if (lane == c) { myz = vz; myc = vc; }
                }
                // ---- gate (cswnv_shift1.py:281-285); lane c holds column c
                float hn = 0.f;
                if (col < NC) {
                    const float gz = (myz + bz + rz + hz) * pz;
                    const float gc = (myc + bc + rc + hcv) * pc;
                    const float z = sigmoidf_(gz);
                    const float carry = __uint_as_float((uint32_t)ld_smem_ll(hx + ((size_t)l * NC + col) * HP + my_i));
                    hn = (1.f - z) * tanhf(gc) + z * carry;
                }
                // ---- fan out: lanes 0..15 write the value into the 16 CTAs, lane 16 feeds the helpers
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float v = __shfl_sync(0xffffffffu, hn, c);
                    if (lane < CORE_CL) st_cluster_ll(hx_peer + (uint32_t)((((l + 1) * NC + c) * HP + my_i) * 8), v, tag);
                    if (lane == 16 && l + 1 < L && P.has_help[l + 1])
                        st_ll(P.hin_ll + P.hin_off[l + 1] + ((size_t)(it % P.hslots[l + 1]) * NC + c) * H + my_i, v, tag);
                }
                // ---- off the critical path: prefetch next step's gate operands, the delay-1 tap, next layer's tap
                prefetch_gate(l, t + 1);
                if (P.dil[l] == 1 && K >= 2) {
                    float w1[TCOLS];
                    tmem_ld<TCOLS>(taddr0 + (P.tap_base[l] + 1) * TCOLS, w1);
                    tmem_wait_ld();
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float az = 0.f, ac = 0.f;
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            az = fmaf(w1[2 * cc], h[c][cc].x, az); az = fmaf(w1[2 * cc + 1], h[c][cc].y, az);
                            ac = fmaf(w1[FPT + 2 * cc], h[c][cc].x, ac); ac = fmaf(w1[FPT + 2 * cc + 1], h[c][cc].y, ac);
                        }
                        const float vz = warp_sum(az), vc = warp_sum(ac);
                        if (lane == c) {
                            float* r0 = ring + (((size_t)(l * 2 + ((t + 1) & 1)) * 16 + warp) * 2) * BS + c;
                            r0[0] = vz; r0[BS] = vc;
                        }
                    }
                }
                if (l + 1 < L) tmem_ld<TCOLS>(taddr0 + P.tap_base[l + 1] * TCOLS, w);
            } else {
                cp_async_commit();                                 // keep the group count uniform
            }
            // ---- skip row of this warp: out_skip[l-1] applied to the input of layer l
            if (active && l >= 1 && warp < D.nskip) {
                float2 ws[NCH];
                load_frag<NCH>(ws, sSkip + ((size_t)(l - 1) * D.nskip + warp) * HP, lane);
#pragma unroll
                for (int c = 0; c < NC; ++c) sk[c] += dot_frag<NCH>(ws, h[c]);
            }
            __syncwarp();
        }
        if (__syncthreads_or(!ok)) { ok = false; break; }
        SWN_CPROF(1);
        if (active) {
            // ---- out_skip[L-1] on the output of the last layer; relu(sum of skips) -> x1 on every CTA
            {
                float2 hl[NC][NCH];
#pragma unroll
                for (int c = 0; c < NC; ++c) ok = wait_vec<NCH>(hx + ((size_t)L * NC + c) * HP, tag, hl[c], lane, P, 510) && ok;
                if (warp < D.nskip) {
                    float2 ws[NCH];
                    load_frag<NCH>(ws, sSkip + ((size_t)(L - 1) * D.nskip + warp) * HP, lane);
                    const float cst = __ldg(P.wbuf + D.skipc_off + warp);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const float v = warp_sum(sk[c] + dot_frag<NCH>(ws, hl[c])) + cst;
                        if (lane < CORE_CL) st_cluster_ll(x1_peer + (uint32_t)((c * Sp + D.skip0 + warp) * 8), fmaxf(v, 0.f), tag);
                    }
                }
            }
            SWN_CPROF(2);
            // ---- out_1 row of this warp (relu) -> x2 on every CTA
            if (warp < D.nb) {
                float acc1[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) acc1[c] = 0.f;
                for (int cc = 0; cc < (Sp >> 6); ++cc) {
                    const float2 wv = reinterpret_cast<const float2*>(sW1 + (size_t)warp * Sp)[cc * 32 + lane];
                    const bool pad = (cc * 64 + 2 * lane) >= S, pad1 = (cc * 64 + 2 * lane + 1) >= S;   // never written
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const unsigned long long* p = x1 + (size_t)c * Sp + cc * 64 + 2 * lane;
                        ulonglong2 v;
                        ok = wait_pair(p, tag, pad, pad1, v, P, 520) && ok;
                        acc1[c] = fmaf(wv.x, __uint_as_float((uint32_t)v.x), acc1[c]);
                        acc1[c] = fmaf(wv.y, __uint_as_float((uint32_t)v.y), acc1[c]);
                    }
                }
                const float b1 = __ldg(P.wbuf + D.b1_off + warp);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float v = warp_sum(acc1[c]) + b1;
                    if (lane < CORE_CL) st_cluster_ll(x2_peer + (uint32_t)((c * n1p + D.b0 + warp) * 8), fmaxf(v, 0.f), tag);
                }
            }
            SWN_CPROF(3);
            // ---- out_2 rows
            for (int r = warp; r < D.nc; r += SWN_NWARPS) {
                float acc2[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) acc2[c] = 0.f;
                for (int cc = 0; cc < (n1p >> 6); ++cc) {
                    const float2 wv = reinterpret_cast<const float2*>(sW2 + (size_t)r * n1p)[cc * 32 + lane];
                    const bool pad = (cc * 64 + 2 * lane) >= n1, pad1 = (cc * 64 + 2 * lane + 1) >= n1;
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const unsigned long long* p = x2 + (size_t)c * n1p + cc * 64 + 2 * lane;
                        ulonglong2 v;
                        ok = wait_pair(p, tag, pad, pad1, v, P, 530) && ok;
                        acc2[c] = fmaf(wv.x, __uint_as_float((uint32_t)v.x), acc2[c]);
                        acc2[c] = fmaf(wv.y, __uint_as_float((uint32_t)v.y), acc2[c]);
                    }
                }
                const float b2 = __ldg(P.wbuf + D.b2_off + r);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float o = warp_sum(acc2[c]) + b2;
                    const int bglob = min(c, P.B - 1);
                    if (is_c) {
                        if (lane == 0) o2[c * 16 + r] = o;
                    } else {
                        const int row = D.c0 + r;
                        if (lane == 0 && P.head_out && c < P.B) P.head_out[((size_t)c * P.n_iter + istep) * n2 + row] = o;
                        const float q = P.noise ? P.noise[((size_t)bglob * P.n_out + istep) * Q + row]
                                                : softmax_q(P.seed, bglob, istep, row);
                        const float g = o - logf(q);
                        if (lane < CORE_CL) st_cluster_ll(gx_peer + (uint32_t)((c * 256 + row) * 8), g, tag);
                    }
                }
            }
            if (is_c) {
                __syncthreads();
                if (tid < NC) {
                    const int c = tid, bglob = min(c, P.B - 1);
                    float eps[1], smp[1];
                    eps[0] = P.noise ? P.noise[(size_t)bglob * P.n_out + istep] : laplace_eps(P.seed, bglob, istep);
                    const float* forced = P.forced ? (const float*)P.forced + (size_t)bglob * P.n_out + istep : nullptr;
                    laplace_step<1>(P.lpc, o2 + c * 16, lpb + c * 8, eps, forced, smp);
                    hist[c * HIST + (istep % HIST)] = forced ? forced[0] : smp[0];
                    if (rank == 0 && c < P.B) {
                        ((float*)P.out)[(size_t)c * P.n_out + istep] = smp[0];
                        if (P.head_out)
                            for (int r = 0; r < n2; ++r) P.head_out[((size_t)c * P.n_iter + istep) * n2 + r] = o2[c * 16 + r];
                    }
                }
            } else if (warp == 0) {
                // ---- argmax over the 256 published Gumbel scores (first maximum wins), one warp per CTA
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float best = -CUDART_INF_F; int bi = 0x7fffffff;
                    for (int r0 = 0; r0 < Q; r0 += 64) {
                        const unsigned long long* p = gx + (size_t)c * 256 + r0 + 2 * lane;
                        ulonglong2 v;
                        ok = wait_pair(p, tag, false, false, v, P, 540) && ok;
                        const float g0 = __uint_as_float((uint32_t)v.x), g1 = __uint_as_float((uint32_t)v.y);
                        const int i0 = r0 + 2 * lane;
                        if (g0 > best) { best = g0; bi = i0; }
                        if (g1 > best) { best = g1; bi = i0 + 1; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                    }
                    if (lane == 0) {
                        int fed = bi;
                        if (c < P.B) {
                            if (rank == 0) ((long long*)P.out)[(size_t)c * P.n_out + istep] = bi;
                            if (P.forced) fed = (int)((const long long*)P.forced)[(size_t)c * P.n_out + istep];
                        }
                        hist[c * HIST + (istep % HIST)] = __int_as_float(fed);
                    }
                }
            }
            if (__syncthreads_or(!ok)) { ok = false; break; }
        }
        SWN_CPROF(4);
        if (it + 1 < P.n_it) {
            input_publish(it + 1);
            input_partial(it + 2);
        }
        SWN_CPROF(5);
    }
    // ---- teardown
    cp_async_wait_all();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_free_all(tbase);
}

template <int NCH, int BS>
__global__ void __cluster_dims__(CORE_CL, 1, 1) __launch_bounds__(SWN_NTHREADS, 1)
swn_core_kernel(const __grid_constant__ CoreParams P) {
    extern __shared__ __align__(16) float smem[];
    if (cluster_id_x() == 0) {
        core_cta<NCH, BS>(P, P.cdesc[cluster_ctarank()], smem);
    } else {
        const int hidx = (int)blockIdx.x - CORE_CL;
        if (hidx < P.n_help && P.hdesc[hidx].npairs > 0) helper_cta<NCH, BS>(P, P.hdesc[hidx], smem);
    }
}

}  // namespace swn
