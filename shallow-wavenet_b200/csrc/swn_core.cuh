// Latency path of the generation loop (small batch, one sample per step): the whole dependent
// chain of a time step runs inside ONE 16-CTA thread-block cluster, the rest of the chip helps.
// This file: the structures shared with the host, the cluster / DSMEM / TMEM primitives and the helper CTA;
// the core CTA and the kernel are in swn_core2.cuh.
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

constexpr int CORE_MAXCL = 3;      // swn_coreseg.cuh: core clusters (utterances) of one launch that share the helper CTAs

struct CoreUtt {                   // swn_coreseg.cuh: what differs between the core clusters of one launch
    const float* P; const float* noise;
    const void* forced; const void* seedw; void* out; float* head_out;
    unsigned long long* hin_ll; unsigned long long* hacc_ll;
    int32_t b0, nb;                // first utterance (Philox counter) and number of utterances of the cluster
};

struct CoreParams {
    int32_t kind, H, Hp, S, Sp, K, L, lpc, n1, n1p, n2, Q, U, R, B, Tf;
    int32_t it0, n_it, n_iter, n_core_taps, n_help, audio_in;
    int32_t b0, help_delay_ns;     // index of this launch's first utterance in the caller's batch (Philox counter); test hook:
                                   // nanoseconds every helper CTA sleeps per turn (swn_gen_args::debug_helper_delay_ns)
    int64_t n_out;
    int32_t dil[16];
    int32_t tap_base[16];          // first core tap index of layer l (tap 0 = delay 0, tap 1 = delay 1)
    int32_t ncore[16];             // core taps of layer l (swn_coreseg.cuh: the taps with delay < seg)
    int32_t t_first, seg;          // swn_coreseg.cuh: time of the first column of block 0 (<= 0), samples per step
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
    int32_t n_cl, pad_cl;          // swn_coreseg.cuh: clusters 0 .. n_cl-1 are core clusters, one utterance each (cl[])
    CoreUtt cl[CORE_MAXCL];
};

// ------------------------------------------------------------------ cluster / DSMEM / TMEM primitives
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
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

// ---- mbarrier + st.async: the exchange inside the cluster.  A producer lane stores one fp32 into a
// peer's shared memory and, in the same instruction, credits 4 bytes to that peer's mbarrier; a
// consumer warp sleeps in mbarrier.try_wait (no issue slots, no shared-memory polling) and is woken
// when the whole vector has landed.  Measured one-way latency 339 cycles vs 428 for a polled
// {value,tag} word (tools/microbench/hop_latency.cu).
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arm(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long* b, uint32_t parity) {
    uint32_t done;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_mbar) : "memory");
}

// the same, executed only where `pred` is non-zero - as a predicated instruction, not a branch (a divergent `if` around the
// asm statement made the compiler re-materialise the lane id and reload spilled addresses in every unrolled iteration)
__device__ __forceinline__ void st_async_f32_if(uint32_t pred, uint32_t remote_addr, float v, uint32_t remote_mbar) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t"
                 "@p st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];\n\t}"
                 ::"r"(remote_addr), "r"(__float_as_uint(v)), "r"(remote_mbar), "r"(pred) : "memory");
}

__device__ __forceinline__ void st_async_f32x2(uint32_t remote_addr, float lo, float hi, uint32_t remote_mbar) {
    const unsigned long long w = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo);
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 ::"r"(remote_addr), "l"(w), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tmem_ld16(uint32_t a, float* r) {
    uint32_t x[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),
                   "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15])
                 : "r"(a));
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(x[i]);
}
// N floats (multiple of 16) per thread, asynchronous: tmem_wait_ld() before use
template <int N>
__device__ __forceinline__ void tmem_ldw(uint32_t a, float (&r)[N]) {
#pragma unroll
    for (int i = 0; i < N; i += 16) tmem_ld16(a + i, &r[i]);
}

// Sleep on a local mbarrier until the phase with `parity` has completed (with the watchdog).
static __device__ __noinline__ bool mbar_wait_slow(unsigned long long* b, uint32_t parity, const CoreParams& P, int where) {
    uint32_t spins = 0;
    long long t_start = clock64();
    for (;;) {
        if (mbar_try(b, parity)) return true;
        if (((++spins) & 15u) == 0) {
            volatile int32_t* st = P.status;
            if (st[0] != 0) return false;
            if ((unsigned long long)(clock64() - t_start) > (P.timeout_cycles >> 1)) {
                if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)parity; st[3] = blockIdx.x; }
                return false;
            }
        }
    }
}

static __device__ __noinline__ bool wait_gword(const unsigned long long* p, uint32_t tag, uint32_t& payload, const CoreParams& P,
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

__device__ __forceinline__ bool mbar_wait(unsigned long long* b, uint32_t parity, const CoreParams& P, int where) {
    if (mbar_try(b, parity)) return true;
    return mbar_wait_slow(b, parity, P, where);
}

// ------------------------------------------------------------------ helper CTA
// Owns the taps of one layer whose contribution is needed >= 2 steps in the future.  Scatter form:
// the input vector of time `it` adds W_j . h into the partial sum of time it + delay_j; when the
// smallest delay has been applied the sum of that target time is complete and is handed to the core.
template <int NCH, int BS, bool MULTI>
__device__ void helper_cta(const CoreParams& P, const HelpCta& D, float* smem) {
    constexpr int NC = BS, HP = NCH * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = D.layer, np = D.npairs, H = P.H, RL = D.ring_len, nt = D.ntaps;
    // MULTI: several core clusters (their own utterances) share this helper's weights; inputs, partial sums and hand-off
    // rings are per cluster, the clusters take turns (swn_coreseg.cuh has the same scheme)
    const int ncl = MULTI ? P.n_cl : 1;
    float* sW = smem + D.sm_w;
    {
        const float4* src = reinterpret_cast<const float4*>(P.wbuf + D.w_off);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = tid; i < nt * np * 2 * HP / 4; i += SWN_NTHREADS) dst[i] = src[i];
    }
    float* hb = smem + D.sm_h;                                 // [ncl][2][NC][HP]
    for (int i = tid; i < ncl * 2 * NC * HP; i += SWN_NTHREADS) hb[i] = 0.f;
    const size_t ring_f = (size_t)RL * np * 2 * BS;            // [ncl][RL][np][2][BS]
    float* ring0;
    if (D.ring_in_smem) {
        ring0 = smem + D.sm_ring;
        for (int i = tid; i < (int)(ncl * ring_f); i += SWN_NTHREADS) ring0[i] = 0.f;
    } else {
        ring0 = P.ring_scratch + D.ring_off;                   // zeroed by the host
    }
    __syncthreads();
    const int slot_mask = P.hslots[l] - 1, nt_mask = P.hacc_nt[l] - 1, rl_mask = RL - 1;   // powers of two (MULTI: RL need not be)
    // lane <-> (tap j, column c) for the ring update
    const int my_j = lane / NC, my_c = lane - my_j * NC;
    const bool my_live = my_j < nt;
    const int my_delay = my_live ? D.delay[my_j] : 0;
    int hsel = 0;
    constexpr int WPT = (NC * HP + SWN_NTHREADS - 1) / SWN_NTHREADS;          // input words per thread and turn
    unsigned long long pre[WPT];
    if (MULTI) {
        const unsigned long long* s0 = P.cl[0].hin_ll + P.hin_off[l];
#pragma unroll
        for (int r = 0; r < WPT; ++r) {
            const int i = tid + r * SWN_NTHREADS;
            pre[r] = i < NC * H ? ld_ll(s0 + i) : 0ull;
        }
    }
    for (int it = 0; it < P.n_it; ++it) {
        const uint32_t tag = (uint32_t)it + 1u;
        for (int q = 0; q < ncl; ++q) {
        const unsigned long long* in_ll = (MULTI ? P.cl[q].hin_ll : P.hin_ll) + P.hin_off[l];
        unsigned long long* acc_ll = (MULTI ? P.cl[q].hacc_ll : P.hacc_ll) + P.hacc_off[l];
        float* ring = ring0 + (size_t)q * ring_f;
        float* hbuf = hb + (size_t)(q * 2 + hsel) * NC * HP;
        bool ok = true;
        if (MULTI) {
            // the words of this turn were requested before the previous turn's arithmetic (below): with several clusters to serve
            // the two L2 round trips per turn are what made three clusters helper-bound (as in swn_coreseg.cuh)
            const unsigned long long* src = in_ll + (size_t)(it & slot_mask) * NC * H;
#pragma unroll
            for (int r = 0; r < WPT; ++r) {
                const int i = tid + r * SWN_NTHREADS;
                if (i < NC * H) {
                    uint32_t bits = (uint32_t)pre[r];
                    if ((uint32_t)(pre[r] >> 32) != tag) ok = wait_gword(src + i, tag, bits, P, 300 + l) && ok;
                    const int c = i / H, k = i - c * H;
                    hbuf[c * HP + k] = __uint_as_float(bits);
                }
            }
        } else {
            const unsigned long long* src = in_ll + (size_t)(it & slot_mask) * NC * H;
#pragma unroll
            for (int c = 0; c < NC; ++c)
                for (int k = tid; k < H; k += SWN_NTHREADS) {
                    unsigned long long v = ld_ll(src + c * H + k);
                    uint32_t bits = (uint32_t)v;
                    if ((uint32_t)(v >> 32) != tag) ok = wait_gword(src + c * H + k, tag, bits, P, 300 + l) && ok;
                    hbuf[c * HP + k] = __uint_as_float(bits);
                }
        }
        if (__syncthreads_or(!ok)) return;
        if (MULTI) {
            int qn = q + 1, itn = it;
            if (qn == ncl) { qn = 0; ++itn; }
            if (itn < P.n_it) {
                const unsigned long long* sn = P.cl[qn].hin_ll + P.hin_off[l] + (size_t)(itn & slot_mask) * NC * H;
#pragma unroll
                for (int r = 0; r < WPT; ++r) {
                    const int i = tid + r * SWN_NTHREADS;
                    pre[r] = i < NC * H ? ld_ll(sn + i) : 0ull;
                }
            }
        }
        if (P.help_delay_ns > 0) __nanosleep((unsigned)P.help_delay_ns);   // test hook: a late helper (the core must wait for its sums)
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
                float* r0 = ring + (((MULTI ? T % RL : T & rl_mask) * np + pi) * 2) * BS + my_c;
                const float a = r0[0] + mz, b = r0[BS] + mc;
                if (my_j == 0) {
                    // smallest delay: the sum for target time T is complete -> hand it to the core, clear the slot
                    r0[0] = 0.f; r0[BS] = 0.f;
                    unsigned long long* dst = acc_ll + (size_t)((T & nt_mask) * 2 * H) * NC;
                    st_ll(dst + (size_t)(D.pair0 + pi) * NC + my_c, a, (uint32_t)T + 1u);
                    st_ll(dst + (size_t)(H + D.pair0 + pi) * NC + my_c, b, (uint32_t)T + 1u);
                } else {
                    r0[0] = a; r0[BS] = b;
                }
            }
            __syncwarp();
        }
        }
        hsel ^= 1;
    }
}

// ------------------------------------------------------------------ cp.async helpers (core CTA, swn_core2.cuh)
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

}  // namespace swn
