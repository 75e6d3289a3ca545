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
    int32_t b0, pad_b0;            // index of this launch's first utterance in the caller's batch (Philox counter)
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

// SWN_LANE8: both rows (z, candidate) of a gate pair live in the 8 lanes of the pair group, lane `sub` holding the
// weights of k = 4 * (sub + 8 * i) .. + 3, i < 2 * NCH.  w[zc * 8 * NCH + 4 * i + j]; the layer input is read as
// float4 (the four pair groups read the same 128 bytes: broadcast).  Totals end up in all 8 lanes after 3 butterflies.
template <int NCH>
__device__ __forceinline__ void lane8_dot(const float* w, const float* hvec, int sub, float& tz, float& tc) {
    float az[4] = {0.f, 0.f, 0.f, 0.f}, ac[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2 * NCH; ++i) {
        const float4 hv = reinterpret_cast<const float4*>(hvec)[sub + 8 * i];
        az[0] = fmaf(w[4 * i + 0], hv.x, az[0]); ac[0] = fmaf(w[8 * NCH + 4 * i + 0], hv.x, ac[0]);
        az[1] = fmaf(w[4 * i + 1], hv.y, az[1]); ac[1] = fmaf(w[8 * NCH + 4 * i + 1], hv.y, ac[1]);
        az[2] = fmaf(w[4 * i + 2], hv.z, az[2]); ac[2] = fmaf(w[8 * NCH + 4 * i + 2], hv.z, ac[2]);
        az[3] = fmaf(w[4 * i + 3], hv.w, az[3]); ac[3] = fmaf(w[8 * NCH + 4 * i + 3], hv.w, ac[3]);
    }
    float z = (az[0] + az[1]) + (az[2] + az[3]), c = (ac[0] + ac[1]) + (ac[2] + ac[3]);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        z += __shfl_xor_sync(0xffffffffu, z, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    tz = z; tc = c;
}

// Sum 8 per-lane partials over the 32 lanes: afterwards lane L holds the total of value (L >> 2)
// (9 shuffles instead of 40).
__device__ __forceinline__ float reduce8(const float (&v)[8], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float u[4], t2[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b3 ? u[i] : u[i + 2], keep = b3 ? u[i + 2] : u[i];
        t2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = b2 ? t2[0] : t2[1], keep = b2 ? t2[1] : t2[0];
    float sres = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    sres += __shfl_xor_sync(0xffffffffu, sres, 2);
    sres += __shfl_xor_sync(0xffffffffu, sres, 1);
    return sres;
}

#ifndef SWN_FINE_PROF
#define SWN_FINE_PROF 0      // 1: clock64 stamps inside the stages (tools/gpu_prof_core.py); they cost ~1.4 us per step
#endif
#if SWN_FINE_PROF
#define SWN_CCLK(ev)                                                                               \
    do {                                                                                           \
        if (P.prof && threadIdx.x == 0 && l == 3 && it >= P.prof_start && it < P.prof_start + P.prof_count) \
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + (ev)] = (unsigned long long)clock64(); \
    } while (0)
#define SWN_HCLK(ev)                                                                               \
    do {                                                                                           \
        if (P.prof && threadIdx.x == 0 && it >= P.prof_start && it < P.prof_start + P.prof_count) \
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + (ev)] = (unsigned long long)clock64(); \
    } while (0)
#else
#define SWN_CCLK(ev) do {} while (0)
#define SWN_HCLK(ev) do {} while (0)
#endif
#define SWN_CPROF(ev)                                                                              \
    do {                                                                                           \
        if (P.prof && threadIdx.x == 0 && it >= P.prof_start && it < P.prof_start + P.prof_count) \
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + (ev)] = gtime();  \
    } while (0)

#ifndef SWN_ABL
#define SWN_ABL 0            // dev only: ablation timing (tools/ab_build.sh), results are wrong when non-zero
#endif
#ifndef SWN_LANE8
#define SWN_LANE8 0          // chain / delay-1 warps: a row is spread over the 8 lanes of its pair group (3-level reductions)
#endif
#ifndef SWN_LANE8_GUARD
#define SWN_LANE8_GUARD 0
#endif
#ifndef SWN_HBULK
#define SWN_HBULK 0          // helper sums brought in by bulk copies (TMA): 10.99 vs 8.88 us per step - the latency of a small bulk copy exceeds the stage
#endif
#ifndef SWN_ARM_BULK
#define SWN_ARM_BULK 1       // layer barriers re-armed once per step (after the last one completed) instead of per stage
#endif
#ifndef SWN_GN_TEAM
#define SWN_GN_TEAM 1        // Gumbel noise drawn one step ahead by a few lanes of warps 8-11
#endif
#ifndef SWN_LDTM_EARLY
#define SWN_LDTM_EARLY 0     // next layer's taps requested right after the FMA block instead of after the publish
#endif
#ifndef SWN_GOP
#define SWN_GOP 1            // gate operands combined by the delay-1 warps one step ahead: one LDS.128 per stage in the chain
#endif
#ifndef SWN_NOSYNC_L
#define SWN_NOSYNC_L 1       // no CTA barrier between the layer loop and the head
#endif
#ifndef SWN_FWD_HELP
#define SWN_FWD_HELP 0       // helper inputs forwarded by the helper-prefetch team instead of the chain warps
#endif
#ifndef SWN_PFG_W
#define SWN_PFG_W 4          // first warp of the team that prefetches the conditioning rows (4 = the delay-1 warps)
#endif
#ifndef SWN_PFH_W
#define SWN_PFH_W 4         // first warp of the team that fetches the helper sums (4 = the delay-1 warps)
#endif
#ifndef SWN_PUB64
#define SWN_PUB64 0          // layer outputs published as 8-byte messages (two rows each)
#endif
#ifndef SWN_HEAD2
#define SWN_HEAD2 1          // head: weight rows in registers before the vectors land, interleaved reductions, re-arming off warp 0
#endif
#ifndef SWN_GATE1
#define SWN_GATE1 1          // gate with one reciprocal (see the chain warp)
#endif
#ifndef SWN_HELP_L2
#define SWN_HELP_L2 1        // helper sums fetched with L2-scope loads instead of cp.async.ca (stale L1 lines)
#endif
#ifndef SWN_RELAY
#define SWN_RELAY 0          // background warps are released by a local relay barrier instead of the exchange mbarrier
#endif
#ifndef SWN_DEFER_BG
#define SWN_DEFER_BG 0      // 1 measured slower (11.31 vs 10.82 us per step on cfg2): the deferred work collides with the next stage
#endif
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

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
    const int slot_mask = P.hslots[l] - 1, nt_mask = P.hacc_nt[l] - 1, rl_mask = RL - 1;   // all powers of two
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
                float* r0 = ring + (((T & rl_mask) * np + pi) * 2) * BS + my_c;
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
    // TMEM columns per tap per thread: a quarter-warp-group thread keeps 8 rows (4 gate pairs) x FPT weights
    constexpr int NC = BS, HP = NCH * 64, FPT = 2 * NCH, TCOLS = 8 * FPT;
    constexpr int HIST = 16;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = P.H, S = P.S, Sp = P.Sp, K = P.K, L = P.L, n1 = P.n1, n1p = P.n1p, n2 = P.n2, Q = P.Q;
    const uint32_t rank = cluster_ctarank();
    const bool is_c = P.kind == 0;
    const int np = D.npairs;
    const int quarter = warp & 3;                               // TMEM lane quarter == the 4 gate pairs 4*quarter..+3
    const int nip = (D.ni + 3) & ~3;

    // ---- shared memory carve-up (identical offsets in all 16 CTAs)
    float* sSkip = smem + D.sm_skip;                       // [L][nskip][HP]
    float* sW1 = smem + D.sm_w1;                           // [nb][Sp]
    float* sW2 = smem + D.sm_w2;                           // [nc][n1p]
    float* sTin = smem + D.sm_tin;                         // DSWNV [Q][nip]
    float* ring = smem + D.sm_ring;                        // [L][2][16][2][BS]  delay-1 partial sums
    float* hx = smem + D.sm_hx;                            // [L+1][NC][HP] exchanged layer inputs
    float* x1 = smem + D.sm_x1;                            // [NC][Sp]  relu(sum of skips)
    float* x2 = smem + D.sm_x2;                            // [NC][n1p] relu(out_1)
    float* gx = smem + D.sm_gx;                            // [NC][256] Gumbel scores
    PreSlot* pre = reinterpret_cast<PreSlot*>(smem + D.sm_pre);                        // [2][L][16][NC]
    float4* gop = reinterpret_cast<float4*>(smem + D.sm_pre);                          // [2][L][16][NC] (SWN_GOP, same region)
    float* misc = smem + D.sm_misc;
    float* hist = misc;                                    // [NC][HIST] fed-back samples / indices
    float* lpb = hist + NC * HIST;                         // [NC][8]
    float* o2 = lpb + NC * 8;                              // [NC][16] out_2 rows (CSWNV)
    float4* lc = reinterpret_cast<float4*>(o2 + NC * 16);  // [L][16] per (layer, warp): conditioning const z/c, bias z/c
    float* cacc = reinterpret_cast<float*>(lc + L * 16);   // [8][16][NC]  input-stage partial sums by time slot
    float* cstage = cacc + 8 * 16 * NC;                    // [2][8][16][NC] gathered table entries in flight (cp.async)
    float* sfir = cstage + 2 * 8 * 16 * NC;                // CSWNV [8][2][16] FIR taps (cb, v) of this CTA's rows
    float* gnz = sfir + 8 * 2 * 16;                        // DSWNV [2][16][NC] Gumbel noise of this CTA's logit rows (SWN_GN_TEAM)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + D.sm_tmem_slot);
    unsigned long long* mb = reinterpret_cast<unsigned long long*>(tmem_slot + 2);     // [L+1] h vectors, then x1, x2, gx
    unsigned long long* mb_x1 = mb + (L + 1), *mb_x2 = mb + (L + 2), *mb_gx = mb + (L + 3);
    unsigned long long* mbg = mb + (L + 4);                // [L] relay barriers (SWN_RELAY)
    unsigned long long* mbh = mbg + L;                     // [L] completion of the helper-sum bulk copies (SWN_HBULK)
    const uint32_t bytes_h = (uint32_t)(H * NC * 4), bytes_x1 = (uint32_t)(S * NC * 4), bytes_x2 = (uint32_t)(n1 * NC * 4),
                   bytes_gx = (uint32_t)(Q * NC * 4);

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
        // exchanged vectors: padded elements stay 0 forever
        const int nw = (L + 1) * NC * HP + NC * Sp + NC * n1p + NC * 256;
        for (int i = tid; i < nw; i += SWN_NTHREADS) hx[i] = 0.f;
        unsigned long long* pz = reinterpret_cast<unsigned long long*>(pre);
        for (int i = tid; i < (int)(2 * L * 16 * NC * sizeof(PreSlot) / 8); i += SWN_NTHREADS) pz[i] = 0ull;
    }
    for (int i = tid; i < NC * (HIST + 8 + 16); i += SWN_NTHREADS) misc[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < NC * HIST; i += SWN_NTHREADS) hist[i] = is_c ? 0.f : __int_as_float(Q / 2);
    for (int i = tid; i < L * 16; i += SWN_NTHREADS) {
        const int l = i >> 4, wv = i & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wv < D.npairs) {
            v.x = P.pconst[l * 2 * H + D.pair0 + wv]; v.y = P.pconst[l * 2 * H + H + D.pair0 + wv];
            v.z = P.wbuf[D.bias_off + (l * D.npairs + wv) * 2]; v.w = P.wbuf[D.bias_off + (l * D.npairs + wv) * 2 + 1];
        }
        lc[i] = v;
    }
    for (int i = tid; i < 8 * 16 * NC + 2 * 8 * 16 * NC; i += SWN_NTHREADS) cacc[i] = 0.f;
    for (int i = tid; i < 8 * 2 * 16; i += SWN_NTHREADS) {
        const int j = i >> 5, w2 = (i >> 4) & 1, r = i & 15;
        sfir[i] = (is_c && j < K && r < D.ni) ? P.in_tab[(j * 2 + w2) * H + D.i0 + r] : 0.f;
    }
    if (tid == 0) {
        for (int i = 0; i < L + 4; ++i) mbar_init(mb + i, 1);
        for (int i = 0; i < L; ++i) { mbar_init(mbg + i, 1); mbar_init(mbh + i, 1); }
        mbar_fence_init();
        for (int i = 0; i <= L; ++i) mbar_arm(mb + i, bytes_h);          // phase 0 of every vector
        mbar_arm(mb_x1, bytes_x1); mbar_arm(mb_x2, bytes_x2); mbar_arm(mb_gx, bytes_gx);
    }
    if (warp == 0) tmem_alloc_all(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t taddr0 = tbase + ((uint32_t)(32 * quarter) << 16);
    // ---- tensor memory <- the core taps: warp q < 4 fills its lane quarter with the taps of gate pairs 4q..4q+3
    // (thread-private columns: tap c occupies columns [c*TCOLS, (c+1)*TCOLS), row rr = 2*pair + {z,c} at rr*FPT)
#if SWN_LANE8
    if (warp < 4) {
        for (int c = 0; c < P.n_core_taps; ++c) {
#pragma unroll
            for (int zc = 0; zc < 2; ++zc) {
#pragma unroll
                for (int i = 0; i < 2 * NCH; ++i) {
                    const int pl = 4 * quarter + (lane >> 3);
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (pl < np)
                        a = reinterpret_cast<const float4*>(P.wbuf + D.tap_off + ((size_t)(c * np + pl) * 2 + zc) * HP)[(lane & 7) + 8 * i];
                    const float wv[4] = {a.x, a.y, a.z, a.w};
                    tmem_st4(taddr0 + c * TCOLS + zc * 8 * NCH + 4 * i, wv);
                }
            }
        }
        tmem_wait_st();
    }
#else
    if (warp < 4) {
        for (int c = 0; c < P.n_core_taps; ++c) {
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const int pl = 4 * quarter + (rr >> 1);
                float wv[FPT];
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) {
                    float2 a = make_float2(0.f, 0.f);
                    if (pl < np)
                        a = reinterpret_cast<const float2*>(P.wbuf + D.tap_off + ((size_t)(c * np + pl) * 2 + (rr & 1)) * HP)[cc * 32 + lane];
                    wv[2 * cc] = a.x; wv[2 * cc + 1] = a.y;
                }
                tmem_st<FPT>(taddr0 + c * TCOLS + rr * FPT, wv);
            }
        }
        tmem_wait_st();
    }
#endif
    if (P.seedw) {
        for (int e = tid; e < NC; e += SWN_NTHREADS) {
            const int bglob = min(e, P.B - 1);
            const float v = is_c ? ((const float*)P.seedw)[bglob]
                                 : __int_as_float(pos_mod(((const long long*)P.seedw)[bglob], Q));
            hist[e * HIST + (HIST - 1)] = v;                       // generated-index -1
            if (is_c && P.lpc > 0) lpb[e * 8 + P.lpc - 1] = v;
        }
    }
    __syncthreads();
    cluster_sync_all();                                            // every CTA's buffers and barriers are ready

    const uint32_t peer = (uint32_t)(lane & (CORE_CL - 1));
    const uint32_t hx_peer = mapa(smem_u32(hx), peer), x1_peer = mapa(smem_u32(x1), peer),
                   x2_peer = mapa(smem_u32(x2), peer), gx_peer = mapa(smem_u32(gx), peer), mb_peer = mapa(smem_u32(mb), peer);
    const int twoH = 2 * H;
    // ---- roles inside the CTA (one warp group = 4 warps = one warp per SM sub-partition):
    //   warps 0-3   "chain" warps: each evaluates the delay-0 tap of 4 gate pairs (8 rows), gates, publishes
    //   warps 4-7   delay-1 taps of the dilation-1 layers + prefetch of next step's gate operands
    //   warps 8-15  two rows each of every out_skip
    // A lane of a chain / delay-1 warp belongs to pair-slot (lane >> 3) of its quarter; column = (lane & 7) for
    // per-column duties (NC <= 8).
    const int pslot = lane >> 3;                                   // 0..3
    const int plocal = 4 * quarter + pslot;                        // gate pair inside this CTA
    const bool pvalid = plocal < np;
    const int my_i = D.pair0 + plocal;                             // row index in H
    const int sub = lane & 7;                                      // lane inside the pair group
    const int skr0 = 2 * (warp - 8), skr1 = skr0 + 1;              // skip rows of warps 8..15
    const bool pub64 = ((np | D.pair0) & 1) == 0;                  // rows come in aligned pairs (SWN_PUB64)
    const uint32_t hxp0 = mapa(smem_u32(hx), (uint32_t)sub), hxp1 = mapa(smem_u32(hx), (uint32_t)sub + 8);
    const uint32_t mbp0 = mapa(smem_u32(mb), (uint32_t)sub), mbp1 = mapa(smem_u32(mb), (uint32_t)sub + 8);
    const float b1w = (warp < D.nb) ? __ldg(P.wbuf + D.b1_off + warp) : 0.f;
    const float b2w = (warp < D.nc) ? __ldg(P.wbuf + D.b2_off + warp) : 0.f;
    float cst0 = 0.f, cst1 = 0.f;
    if (warp >= 8) {
        if (skr0 < D.nskip) cst0 = __ldg(P.wbuf + D.skipc_off + skr0);
        if (skr1 < D.nskip) cst1 = __ldg(P.wbuf + D.skipc_off + skr1);
    }

    // ---- prefetch (cp.async) of the gate operands of (layer l, time tt) into pre[tt&1][l][pair][col], done by the
    // delay-1 warps (4-7): lane sub < NC of every pair group handles column sub
    // which warp team prefetches (any team of four warps covers the four quarters: quarter = warp & 3)
    const bool pf_lane = warp >= SWN_PFG_W && warp < SWN_PFG_W + 4 && pvalid && sub < NC;     // conditioning rows (cp.async)
    const bool pfh_lane = warp >= SWN_PFH_W && warp < SWN_PFH_W + 4 && pvalid && sub < NC;    // helper sums (L2 loads)
    const int bg_col = min(sub, P.B - 1);
    const float* ptab_b = P.P + (size_t)bg_col * P.Tf * L * twoH + my_i;
    auto prefetch_gate = [&](int l, int tt, const float* prow) {
        if (pf_lane && tt < P.n_it) {
            PreSlot* ps = pre + (((tt & 1) * L + l) * 16 + plocal) * NC + sub;
            cp_async4(&ps->pz, prow + l * twoH);
            cp_async4(&ps->pc, prow + l * twoH + H);
#if !SWN_HELP_L2
            if (P.has_help[l] && tt >= P.hdmin[l]) {
                const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((tt & (P.hacc_nt[l] - 1)) * twoH + my_i) * NC + sub;
                cp_async8(&ps->hz, src);
                cp_async8(&ps->hc, src + H * NC);
            }
#endif
        }
    };
    // The helpers' partial sums are re-written every ring period, and cp.async of 8 bytes can only be `.ca`: a line
    // still sitting in this SM's (incoherent) L1 from the previous period comes back with the old tag and sends the
    // chain warp into the slow polling path.  So they are fetched with L2-scope loads by the delay-1 warps, after
    // their own work for the stage.
    auto prefetch_help = [&](int l, int tt) {
#if SWN_HELP_L2
        if ((SWN_HELP_L2 ? pfh_lane : pf_lane) && tt < P.n_it && P.has_help[l] && tt >= P.hdmin[l]) {
            PreSlot* ps = pre + (((tt & 1) * L + l) * 16 + plocal) * NC + sub;
            const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((tt & (P.hacc_nt[l] - 1)) * twoH + my_i) * NC + sub;
            const unsigned long long a = ld_ll(src), b = ld_ll(src + H * NC);
            ps->hz = a; ps->hc = b;
        }
#endif
    };

    // ---- SWN_GOP: everything the gate of (layer l, time tt) needs besides the newest-tap product is known one step
    // ahead: conditioning row (x gain + constant), biases, helper sums, delay-1 sums.  The delay-1 warps fetch the
    // global parts (L2-scope loads, issued before their own FMA work), wait for a late helper themselves, and leave
    //   { pz, (bz + rz + hz) * pz, pc, (bc + rc + hc) * pc }
    // in gop[tt & 1][l][pair][col]; the chain warp's stage then needs one LDS.128 and two FMAs before the gate.
    long long hwait = 0;
    bool gok = true;
    // SWN_HBULK: the helper sums of this CTA's rows are two contiguous runs of {value, tag} words (z rows, c rows); one
    // lane asks for them with two bulk copies (L2 -> shared memory, no L1, no registers, no blocked warp) and the
    // committing lanes sleep on the copy barrier
    unsigned long long* hst = reinterpret_cast<unsigned long long*>(gnz + 2 * 16 * NC);   // [2][L][2][16 * NC]
    const bool hbulk = SWN_HBULK && ((D.pair0 * NC) & 1) == 0 && ((np * NC) & 1) == 0;
    auto gop_issue = [&](int l, int tt, const float* prow, float& prz, float& prc, unsigned long long& ha, unsigned long long& hb) {
        prz = 0.f; prc = 0.f; ha = 0ull; hb = 0ull;
        if (pf_lane && tt < P.n_it) {
            prz = __ldg(prow + l * twoH); prc = __ldg(prow + l * twoH + H);
            if (!hbulk && P.has_help[l] && tt >= P.hdmin[l]) {
                const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((tt & (P.hacc_nt[l] - 1)) * twoH + my_i) * NC + sub;
                ha = ld_ll(src); hb = ld_ll(src + H * NC);
            }
        }
        if (hbulk && tid == 128 && tt < P.n_it && P.has_help[l] && tt >= P.hdmin[l]) {
            const uint32_t bytes = (uint32_t)(np * NC * 8);
            const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((tt & (P.hacc_nt[l] - 1)) * twoH + D.pair0) * NC;
            unsigned long long* dst = hst + (size_t)(((tt & 1) * L + l) * 2) * 16 * NC;
            mbar_arm(mbh + l, 2 * bytes);
            bulk_g2s(dst, src, bytes, mbh + l);
            bulk_g2s(dst + 16 * NC, src + H * NC, bytes, mbh + l);
        }
    };
    auto gop_commit = [&](int l, int tt, float gain_tt, float prz, float prc, unsigned long long ha, unsigned long long hb,
                          float rz, float rc) {
        if (pf_lane && tt < P.n_it) {
            float hz = 0.f, hc = 0.f;
            if (P.has_help[l] && tt >= P.hdmin[l]) {
                const uint32_t tg = (uint32_t)tt + 1u;
                if (hbulk) {
                    gok = mbar_wait(mbh + l, (uint32_t)(tt - P.hdmin[l]) & 1u, P, 700 + l) && gok;
                    const unsigned long long* d = hst + (size_t)(((tt & 1) * L + l) * 2) * 16 * NC + plocal * NC + sub;
                    ha = d[0]; hb = d[16 * NC];
                }
                uint32_t a = (uint32_t)ha, b = (uint32_t)hb;
                if ((uint32_t)(ha >> 32) != tg || (uint32_t)(hb >> 32) != tg) {      // the helper is late: wait for it here
                    const long long tw0 = clock64();
                    const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((tt & (P.hacc_nt[l] - 1)) * twoH + my_i) * NC + sub;
                    gok = wait_gword(src, tg, a, P, 400 + l) && gok;
                    gok = wait_gword(src + H * NC, tg, b, P, 400 + l) && gok;
                    hwait += clock64() - tw0;
                }
                hz = __uint_as_float(a); hc = __uint_as_float(b);
            }
            const float4 c4 = lc[l * 16 + plocal];
            const float pz = fmaf(gain_tt, prz, c4.x), pc = fmaf(gain_tt, prc, c4.y);
            gop[(((tt & 1) * L + l) * 16 + plocal) * NC + sub] = make_float4(pz, (c4.z + rz + hz) * pz, pc, (c4.w + rc + hc) * pc);
        }
    };

    // ---- SWN_GN_TEAM: the Gumbel noise of the CTA's 16 logit rows is drawn one step ahead by (16 * NC) lanes of warps
    // 8-11 while the layers run, instead of by every warp for its own row at the top of the step (a Philox block and
    // two logarithms per warp, on every scheduler, between the input publish and the first layer).
    const bool gn_team = SWN_GN_TEAM && !is_c && D.nc <= 16;
    auto draw_gumbel = [&](int isn) {
        const int idx = (warp - 8) * 32 + lane;
        if (gn_team && warp >= 8 && idx < 16 * NC) {
            const int row = idx & 15, c = idx >> 4, bglob = min(c, P.B - 1);
            if (row < D.nc) {
                const float q = P.noise ? __ldg(P.noise + ((size_t)bglob * P.n_out + isn) * Q + D.c0 + row)
                                        : softmax_q(P.seed, P.b0 + bglob, isn, D.c0 + row);
                gnz[((isn & 1) * 16 + row) * NC + c] = -logf(q);
            }
        }
    };

    // ---- input stage: h0 rows of this CTA (row = i0 + warp), published into every peer's hx[0].
    // Scatter form again: the sample fed at time tau adds tap (K-1-d) into the partial sum of time tau+d
    // (cacc ring).  Lane d-1 handles delay d >= 1; delay 0 (the newest tap) is added at publish time from
    // shared memory.  DSWNV: the older taps are gathered from the L2-resident table with cp.async and
    // folded in one step later (they are needed one step later at the earliest).
    const int in_row = D.i0 + warp;
    const bool has_in = warp < D.ni;
    const float in_b = has_in ? __ldg(P.in_bias + in_row) : 0.f;
    auto fold_stage = [&](int tau) {                                // fold the gathers issued for the sample of time tau
        if (!has_in || is_c || lane >= K - 1 || tau < 0) return;
        const int d = lane + 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float* dst = cacc + (((tau + d) & 7) * 16 + warp) * NC + c;
            *dst += cstage[(((tau & 1) * 8 + lane) * 16 + warp) * NC + c];
        }
    };
    auto feed = [&](int tau) {                                      // the sample of time tau is known (hist)
        if (!has_in || lane >= K - 1) return;
        const int d = lane + 1, j = K - 1 - d;
        const int n = tau - P.R - 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (is_c) {
                const float a = (n >= -1) ? hist[c * HIST + ((n + HIST) & (HIST - 1))] : 0.f;
                cacc[(((tau + d) & 7) * 16 + warp) * NC + c] += sfir[(j * 2) * 16 + warp] + sfir[(j * 2 + 1) * 16 + warp] * a;
            } else {
                const int q = (n >= -1) ? __float_as_int(hist[c * HIST + ((n + HIST) & (HIST - 1))]) : Q / 2;
                cp_async4(&cstage[(((tau & 1) * 8 + lane) * 16 + warp) * NC + c], P.in_tab + ((size_t)j * Q + q) * H + in_row);
            }
        }
    };
    auto input_publish = [&](int itn) {                             // h0 of time itn (its own sample is in hist)
        if (!has_in) return;
        const uint32_t tagn = (uint32_t)itn + 1u;
        const int n = itn - P.R - 1;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            float* slot = cacc + ((itn & 7) * 16 + warp) * NC + c;
            float acc = in_b + *slot;
            if (is_c) {
                const float a = (n >= -1) ? hist[c * HIST + ((n + HIST) & (HIST - 1))] : 0.f;
                acc += sfir[((K - 1) * 2) * 16 + warp] + sfir[((K - 1) * 2 + 1) * 16 + warp] * a;
            } else {
                const int q = (n >= -1) ? __float_as_int(hist[c * HIST + ((n + HIST) & (HIST - 1))]) : Q / 2;
                acc += sTin[q * nip + warp];
            }
            __syncwarp();
            if (lane == 0) *slot = 0.f;                              // the slot is reused 8 steps later
            const float v = softsignf_(acc);
            if (lane < CORE_CL) st_async_f32(hx_peer + (uint32_t)((c * HP + in_row) * 4), v, mb_peer);
            if (lane == 16 && P.has_help[0])
                st_ll(P.hin_ll + P.hin_off[0] + ((itn & (P.hslots[0] - 1)) * NC + c) * H + in_row, v, tagn);
        }
    };
#if SWN_GOP
    __syncthreads();                                               // lc[] is complete
    for (int l = 0; l < L; ++l) {
        float prz, prc; unsigned long long ha, hb;
        gop_issue(l, 0, ptab_b, prz, prc, ha, hb);
        gop_commit(l, 0, __ldg(P.gain), prz, prc, ha, hb, 0.f, 0.f);
    }
#else
    for (int l = 0; l < L; ++l) { prefetch_gate(l, 0, ptab_b); prefetch_help(l, 0); }
#endif
    cp_async_commit();
    int fc = 0, jc = 0;                                            // frame / phase of the current time step
    input_publish(0);                                              // time 0 sees only its own (padding) sample
    feed(0);
    cp_async_commit();

    if (P.it0 == 0) draw_gumbel(0);                                // (made visible by the first step's CTA barrier)
    bool ok = true;
    for (int it = 0; it < P.n_it; ++it) {
        const bool active = it >= P.it0;
        const uint32_t tag = (uint32_t)it + 1u;
        const uint32_t par = (uint32_t)it & 1u;
        const int istep = it - P.it0;
        const uint32_t hpar = (uint32_t)istep & 1u;                // phase of the head vectors (advance only when active)
        const int t = it;
        float sk0[NC], sk1[NC];                                    // running sums of this warp's two skip rows (warps 8-15)
#pragma unroll
        for (int c = 0; c < NC; ++c) { sk0[c] = 0.f; sk1[c] = 0.f; }
        float w[TCOLS];
        if (warp < 4) tmem_ldw<TCOLS>(taddr0 + P.tap_base[0] * TCOLS, w);
        hwait = 0;
        const bool prof_on = P.prof && threadIdx.x == 0 && it >= P.prof_start && it < P.prof_start + P.prof_count;
        const float gq = __ldg(P.gain + jc);
        int fn = fc, jn = jc;                                      // frame / phase of time t+1
        if (t + 1 > P.R && ++jn == P.U) { jn = 0; ++fn; }
        const float* prow_next = ptab_b + (size_t)fn * L * twoH;
        const float gq_next = __ldg(P.gain + jn);
        // Gumbel noise of this warp's logit row: independent of the data, so it is drawn before the chain starts
        float gn[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) gn[c] = 0.f;
        if (active && !is_c && warp < D.nc && !gn_team) {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int bglob = min(c, P.B - 1);
                const float q = P.noise ? __ldg(P.noise + ((size_t)bglob * P.n_out + istep) * Q + D.c0 + warp)
                                        : softmax_q(P.seed, P.b0 + bglob, istep, D.c0 + warp);
                gn[c] = -logf(q);
            }
        }
        cp_async_wait_all();                                       // gate operands / table entries issued one step ago
        __syncwarp();
        fold_stage(it);                                            // table entries gathered for the sample of time `it`
        __syncthreads();                                           // pre[] (written by warps 4-7) is read by warps 0-3
        SWN_CPROF(0);
        if (it + 1 >= P.it0 && it + 1 < P.n_it) draw_gumbel(it + 1 - P.it0);      // for the next step
        const float4* gop_rd = gop + ((par * L) * 16 + plocal) * NC + sub;   // this lane's operands of layer 0
        for (int l = 0; l < L; ++l) {
            const bool ld1 = P.dil[l] == 1;
            // ---- chain warps: gate operands that do not depend on the layer input (one pair group = 8 lanes)
            float pz = 0.f, pc = 0.f, hz = 0.f, hcv = 0.f, rz = 0.f, rc = 0.f, bz = 0.f, bc = 0.f;
#if SWN_GOP
            float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (warp < 4 && pvalid && sub < NC) g4 = gop_rd[l * 16 * NC];
#else
            if (warp < 4 && pvalid && sub < NC && !(SWN_ABL & 1024)) {
                const PreSlot ps = pre[((par * L + l) * 16 + plocal) * NC + sub];
                const float4 c4 = lc[l * 16 + plocal];
                pz = fmaf(gq, ps.pz, c4.x); pc = fmaf(gq, ps.pc, c4.y); bz = c4.z; bc = c4.w;
                if (P.has_help[l] && t >= P.hdmin[l]) {
                    uint32_t a = (uint32_t)ps.hz, b = (uint32_t)ps.hc;
                    if ((uint32_t)(ps.hz >> 32) != tag || (uint32_t)(ps.hc >> 32) != tag) {
                        // the helper was late when the prefetch was issued: wait for it here
                        const long long tw0 = clock64();
                        const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + ((t & (P.hacc_nt[l] - 1)) * twoH + my_i) * NC + sub;
                        ok = wait_gword(src, tag, a, P, 400 + l) && ok;
                        ok = wait_gword(src + H * NC, tag, b, P, 400 + l) && ok;
                        hwait += clock64() - tw0;
                    }
                    hz = __uint_as_float(a); hcv = __uint_as_float(b);
                }
                if (ld1) {
                    const float* r0 = ring + (((l * 2 + (int)par) * 16 + plocal) * 2) * BS + sub;
                    rz = r0[0]; rc = r0[BS];
                }
            }
#endif
            // ---- sleep until the layer input has landed in local shared memory
            SWN_CCLK(14);
#if SWN_DEFER_BG
            // Background warps (delay-1 taps, prefetch, out_skip rows) start their share of stage l only after the
            // chain warps of this CTA have published: the critical path has the LDS / issue slots to itself, the
            // background work runs while the chain warps sleep on the next hop.  Not for the last layer (the head's
            // CTA barrier would wait for it).  Nobody leaves the layer loop early (named barriers need every warp).
            const bool defer = l + 1 < L;
            if (defer && warp >= 4) named_bar_sync(1 + (l % 15), SWN_NTHREADS);
            ok = ok && mbar_wait(mb + l, par, P, 500 + l);
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) {
                if (defer && warp < 4) named_bar_arrive(1 + (l % 15), SWN_NTHREADS);
                continue;
            }
#elif SWN_RELAY
            // Experiment: only the chain warps and warp 4 sleep on the exchange barrier; warp 4 re-arms it and releases
            // the other background warps.  (tools/microbench/mbar_wake.cu: 15 sleepers are released within 30 cycles
            // of each other, so there is little to gain.)
            if (warp <= 4) {
                ok = mbar_wait(mb + l, par, P, 500 + l) && ok;
                if (tid == 128) { mbar_arm(mb + l, bytes_h); mbar_arrive(mbg + l); }
            } else {
                ok = mbar_wait(mbg + l, par, P, 600 + l) && ok;
            }
            if (!__all_sync(0xffffffffu, ok)) { ok = false; break; }
#else
            ok = mbar_wait(mb + l, par, P, 500 + l) && ok;
            if (!__all_sync(0xffffffffu, ok)) { ok = false; break; }
#endif
            SWN_CCLK(15);
#if !SWN_RELAY
#if SWN_ARM_BULK
            // next phase (next time step) of every layer barrier, armed in one go once the last of them has completed:
            // a per-stage `if (tid == 128)` is predicated, and every warp pays for its operand loads
            if (l == L - 1 && warp == 4) {
                if (lane == 0)
                    for (int i = 0; i < L; ++i) mbar_arm(mb + i, bytes_h);
                __syncwarp();
            }
#else
            if (tid == (SWN_HEAD2 ? 128 : 0)) mbar_arm(mb + l, bytes_h);   // next phase (next time step); not by a chain warp
#endif
#endif
            if (prof_on && l < 6) P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + 1 + l] = gtime();
            float2 h[NC][NCH];
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (!(SWN_LANE8 && SWN_GOP && SWN_LANE8_GUARD) || warp >= 8) load_frag<NCH>(h[c], hx + (((SWN_ABL & 2048) ? 0 : l) * NC + c) * HP, lane);
            if (warp < 4) {
                // ================= chain warp: delay-0 tap of 4 gate pairs, gate, publish =================
                tmem_wait_ld();
                SWN_CCLK(16);
                float myqz = 0.f, myqc = 0.f;                      // lane (pair group, sub = column) keeps its column's sums
#if SWN_LANE8
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float qz, qc;
                    lane8_dot<NCH>(w, hx + (l * NC + c) * HP, sub, qz, qc);
                    if (sub == c) { myqz = qz; myqc = qc; }
                }
#else
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float v8[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        float a = 0.f, b = 0.f;
#pragma unroll
                        for (int cc = 0; cc < ((SWN_ABL & 1) ? 1 : NCH); ++cc) {
                            a = fmaf(w[rr * FPT + 2 * cc], h[c][cc].x, a);
                            b = fmaf(w[rr * FPT + 2 * cc + 1], h[c][cc].y, b);
                        }
                        v8[rr] = a + b;
                    }
#if SWN_ABL & 2
                    const float mine = v8[0] + v8[1] + v8[2] + v8[3] + v8[4] + v8[5] + v8[6] + v8[7];
                    const float other = mine;
#else
                    const float mine = reduce8(v8, lane);          // lane holds row (lane >> 2): even = z, odd = candidate
                    const float other = __shfl_xor_sync(0xffffffffu, mine, 4);
#endif
                    const float qz = (lane & 4) ? other : mine, qc = (lane & 4) ? mine : other;
                    if (sub == c) { myqz = qz; myqc = qc; }
                }
#endif
#if SWN_LDTM_EARLY
                // the tap registers are dead from here on: the next layer's taps stream in behind the reduction and the gate
                if (l + 1 < L) tmem_ldw<TCOLS>(taddr0 + P.tap_base[l + 1] * TCOLS, w);
#endif
                SWN_CCLK(17);
                // ---- gate (cswnv_shift1.py:281-285), all columns at once (lane sub = column sub)
                float hn = 0.f;
                if (sub < NC) {
#if SWN_GOP
                    const float gz = fmaf(myqz, g4.x, g4.y);
                    const float gc = fmaf(myqc, g4.z, g4.w);
#else
                    const float gz = (myqz + bz + rz + hz) * pz;
                    const float gc = (myqc + bc + rc + hcv) * pc;
#endif
                    // sigmoid / tanh through ex2.approx + rcp (abs. error ~1e-7, inside the 1e-4 head tolerance);
                    // both exponentials are independent, so their latencies overlap
#if SWN_ABL & 4
                    hn = gz * 1e-3f + gc * 1e-3f;
#elif SWN_GATE1
                    // (1-z) tanh(gc) + z carry over one common denominator: two ex2, one rcp on the dependent chain
                    //   z = 1/(1+ez), tanh = sign(gc) (1-ec)/(1+ec)  ->  [ez sign(gc) (1-ec) + carry (1+ec)] / [(1+ez)(1+ec)]
                    // ez is capped at 2^43 (gz >= -30: z differs from 0 by < 1e-13 either way) so the product stays finite.
                    const float carry = hx[(l * NC + sub) * HP + (pvalid ? my_i : 0)];
                    const float ez = ex2_approx(fmaxf(gz, -30.f) * -1.4426950408889634f);
                    const float ec = ex2_approx(fabsf(gc) * -2.8853900817779268f);
                    const float den = (1.f + ez) * (1.f + ec);
                    const float num = fmaf(ez, copysignf(1.f - ec, gc), carry * (1.f + ec));
                    hn = num * rcp_approx(den);
#else
                    const float ez = __expf(-gz), ec = __expf(-2.f * fabsf(gc));
                    const float z = __frcp_rn(1.f + ez);
                    const float th = copysignf((1.f - ec) * __frcp_rn(1.f + ec), gc);
                    const float carry = hx[(l * NC + sub) * HP + (pvalid ? my_i : 0)];
                    hn = (1.f - z) * th + z * carry;
#endif
                }
                SWN_CCLK(18);
                // ---- fan out: in every pair group, lane sub writes CTAs sub and sub+8; lane sub==0 feeds the helpers
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float v = __shfl_sync(0xffffffffu, hn, (lane & 24) + c);
#if SWN_PUB64
                    // two adjacent rows per message: half as many remote stores and barrier transactions per hop
                    // (tools/microbench/a2a_hop.cu).  Even pair groups serve CTAs 0-7, odd ones CTAs 8-15.
                    const float vo = __shfl_xor_sync(0xffffffffu, v, 8);
                    if (pub64) {
                        if (pvalid) {
                            const uint32_t off = (uint32_t)((((l + 1) * NC + c) * HP + (my_i & ~1)) * 4);
                            const uint32_t mbo = (uint32_t)((l + 1) * 8);
                            st_async_f32x2((pslot & 1) ? hxp1 + off : hxp0 + off, (pslot & 1) ? vo : v, (pslot & 1) ? v : vo,
                                           (pslot & 1) ? mbp1 + mbo : mbp0 + mbo);
                            if (!SWN_FWD_HELP && sub == 0 && l + 1 < L && P.has_help[l + 1])
                                st_ll(P.hin_ll + P.hin_off[l + 1] + ((it & (P.hslots[l + 1] - 1)) * NC + c) * H + my_i, v, tag);
                        }
                    } else
#endif
                    if (pvalid) {
                        const uint32_t off = (uint32_t)((((l + 1) * NC + c) * HP + my_i) * 4);
                        const uint32_t mbo = (uint32_t)((l + 1) * 8);
                        st_async_f32(hxp0 + off, v, mbp0 + mbo);
                        st_async_f32(hxp1 + off, v, mbp1 + mbo);
                        if (!SWN_FWD_HELP && sub == 0 && l + 1 < L && P.has_help[l + 1])
                            st_ll(P.hin_ll + P.hin_off[l + 1] + ((it & (P.hslots[l + 1] - 1)) * NC + c) * H + my_i, v, tag);
                    }
                }
                SWN_CCLK(19);
#if SWN_DEFER_BG
                if (defer) named_bar_arrive(1 + (l % 15), SWN_NTHREADS);
#endif
                if (!SWN_LDTM_EARLY && l + 1 < L && !(SWN_ABL & 512)) tmem_ldw<TCOLS>(taddr0 + P.tap_base[l + 1] * TCOLS, w);
                SWN_CCLK(20);
            } else if (warp < 8) {
                // ================= delay-1 tap of dilation-1 layers (for the next step) + operand prefetch =================
#if SWN_GOP
                float prz, prc, rz1 = 0.f, rc1 = 0.f; unsigned long long ha, hb;
                gop_issue(l, t + 1, prow_next, prz, prc, ha, hb);
#else
                if (SWN_PFG_W == 4 && !(SWN_ABL & 32)) prefetch_gate(l, t + 1, prow_next);
#endif
                if (ld1 && K >= 2 && !(SWN_ABL & 8)) {
                    tmem_ldw<TCOLS>(taddr0 + (P.tap_base[l] + 1) * TCOLS, w);
                    tmem_wait_ld();
#if SWN_LANE8 && SWN_GOP
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float mz, mc;
                        lane8_dot<NCH>(w, hx + (l * NC + c) * HP, sub, mz, mc);
                        if (sub == c) { rz1 = mz; rc1 = mc; }
                    }
#else
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float v8[8];
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) {
                            float a = 0.f, b = 0.f;
#pragma unroll
                            for (int cc = 0; cc < NCH; ++cc) {
                                a = fmaf(w[rr * FPT + 2 * cc], h[c][cc].x, a);
                                b = fmaf(w[rr * FPT + 2 * cc + 1], h[c][cc].y, b);
                            }
                            v8[rr] = a + b;
                        }
                        const float mine = reduce8(v8, lane);
#if SWN_GOP
                        // row (lane >> 2) = 2 * pair slot + {z, c}: hand both sums of the pair to the lane of column c
                        const float mz = __shfl_sync(0xffffffffu, mine, lane & 24), mc = __shfl_sync(0xffffffffu, mine, (lane & 24) + 4);
                        if (sub == c) { rz1 = mz; rc1 = mc; }
#else
                        if ((lane & 3) == 0 && pvalid)             // row (lane >> 2) = 2 * pair slot + {z, c}
                            ring[(((l * 2 + (int)(par ^ 1u)) * 16 + plocal) * 2 + ((lane >> 2) & 1)) * BS + c] = mine;
#endif
                    }
#endif
                }
#if SWN_GOP
                gop_commit(l, t + 1, gq_next, prz, prc, ha, hb, rz1, rc1);
#else
                if (SWN_PFH_W == 4) prefetch_help(l, t + 1);
#endif
            } else {
                if (!SWN_GOP && SWN_PFG_W != 4 && !(SWN_ABL & 32)) prefetch_gate(l, t + 1, prow_next);
#if SWN_FWD_HELP
                // the helpers of layer l get this CTA's rows of the layer input from here (one hop later than from the
                // producing chain warp; the helpers have a whole step of slack).  Layer 0 is fed by the input stage.
                if (l >= 1 && P.has_help[l] && pfh_lane)
                    st_ll(P.hin_ll + P.hin_off[l] + ((it & (P.hslots[l] - 1)) * NC + sub) * H + my_i,
                          hx[(l * NC + sub) * HP + my_i], tag);
#endif
                // ================= two rows of out_skip[l-1] applied to the input of layer l =================
                if (active && l >= 1 && !(SWN_ABL & 16)) {
                if (skr0 < D.nskip) {
                    float2 ws[NCH];
                    load_frag<NCH>(ws, sSkip + ((l - 1) * D.nskip + skr0) * HP, lane);
#pragma unroll
                    for (int c = 0; c < NC; ++c) sk0[c] += dot_frag<NCH>(ws, h[c]);
                }
                if (skr1 < D.nskip) {
                    float2 ws[NCH];
                    load_frag<NCH>(ws, sSkip + ((l - 1) * D.nskip + skr1) * HP, lane);
#pragma unroll
                    for (int c = 0; c < NC; ++c) sk1[c] += dot_frag<NCH>(ws, h[c]);
                }
                }
                if (!SWN_GOP && SWN_PFH_W != 4) prefetch_help(l, t + 1);
            }
            __syncwarp();
        }
        cp_async_commit();
#if !SWN_NOSYNC_L
        if (__syncthreads_or(!ok)) { ok = false; break; }
#endif
        // (no CTA barrier here: nothing below depends on another warp's layer work, and the background warps' last
        // L2 loads would otherwise be exposed in front of the head; a failed wait is voted on at the end of the step)
        SWN_CPROF(7);
        if (P.prof && threadIdx.x == (SWN_GOP ? 128 : 0) && it >= P.prof_start && it < P.prof_start + P.prof_count)
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + 13] = (unsigned long long)hwait;
        ok = ok && gok;
        // the output of the last layer (always consumed, so that its barrier advances every step)
        ok = mbar_wait(mb + L, par, P, 510) && ok;
        if (tid == (SWN_HEAD2 ? 128 : 0)) mbar_arm(mb + L, bytes_h);
        SWN_HCLK(31);
        if (active) {
            // ---- out_skip[L-1] on the output of the last layer; relu(sum of skips) -> x1 on every CTA
            if (warp >= 8) {
                float2 hl[NC][NCH];
#pragma unroll
                for (int c = 0; c < NC; ++c) load_frag<NCH>(hl[c], hx + (L * NC + c) * HP, lane);
#if SWN_HEAD2
                // both rows' partial sums first, then the two reductions interleaved (independent shuffle chains)
                {
                    const bool v0 = skr0 < D.nskip, v1 = skr1 < D.nskip;
                    float2 ws0[NCH], ws1[NCH];
                    load_frag<NCH>(ws0, sSkip + ((L - 1) * D.nskip + (v0 ? skr0 : 0)) * HP, lane);
                    load_frag<NCH>(ws1, sSkip + ((L - 1) * D.nskip + (v1 ? skr1 : 0)) * HP, lane);
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float a = sk0[c] + dot_frag<NCH>(ws0, hl[c]), b = sk1[c] + dot_frag<NCH>(ws1, hl[c]);
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            a += __shfl_xor_sync(0xffffffffu, a, o);
                            b += __shfl_xor_sync(0xffffffffu, b, o);
                        }
                        if (lane < CORE_CL) {
                            if (v0) st_async_f32(x1_peer + (uint32_t)((c * Sp + D.skip0 + skr0) * 4), fmaxf(a + cst0, 0.f), mb_peer + (uint32_t)((L + 1) * 8));
                            if (v1) st_async_f32(x1_peer + (uint32_t)((c * Sp + D.skip0 + skr1) * 4), fmaxf(b + cst1, 0.f), mb_peer + (uint32_t)((L + 1) * 8));
                        }
                    }
                }
#else
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int row = half ? skr1 : skr0;
                    if (row < D.nskip) {
                        float2 ws[NCH];
                        load_frag<NCH>(ws, sSkip + ((L - 1) * D.nskip + row) * HP, lane);
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
#if SWN_ABL & 256
                            const float v = (half ? sk1[c] : sk0[c]) + ws[0].x * hl[c][0].x + (half ? cst1 : cst0);
#else
                            const float v = warp_sum((half ? sk1[c] : sk0[c]) + dot_frag<NCH>(ws, hl[c])) + (half ? cst1 : cst0);
#endif
                            if (lane < CORE_CL)
                                st_async_f32(x1_peer + (uint32_t)((c * Sp + D.skip0 + row) * 4), fmaxf(v, 0.f), mb_peer + (uint32_t)((L + 1) * 8));
                        }
                    }
                }
#endif
            }
            SWN_CPROF(8);
            // ---- out_1 row of this warp (relu) -> x2 on every CTA
            if (warp < D.nb) {
                SWN_HCLK(21);
                float acc1[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) acc1[c] = 0.f;
#if SWN_HEAD2
                // the weight row is in registers before the vector lands (up to 256 columns; longer rows continue below)
                float2 wpre[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wpre[cc] = (cc * 64 < Sp) ? reinterpret_cast<const float2*>(sW1 + (size_t)warp * Sp)[cc * 32 + lane] : make_float2(0.f, 0.f);
                ok = mbar_wait(mb_x1, hpar, P, 520) && ok;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (cc * 64 < Sp) {
                            const float2 v = reinterpret_cast<const float2*>(x1 + (size_t)c * Sp)[cc * 32 + lane];
                            a = fmaf(wpre[cc].x, v.x, a); b = fmaf(wpre[cc].y, v.y, b);
                        }
                    }
                    acc1[c] = a + b;
                }
                for (int cc = 4; cc < (Sp >> 6); ++cc) {
#else
                ok = mbar_wait(mb_x1, hpar, P, 520) && ok;
                SWN_HCLK(22);
                for (int cc = 0; cc < ((SWN_ABL & 64) ? 1 : (Sp >> 6)); ++cc) {
#endif
                    const float2 wv = reinterpret_cast<const float2*>(sW1 + (size_t)warp * Sp)[cc * 32 + lane];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const float2 v = reinterpret_cast<const float2*>(x1 + (size_t)c * Sp)[cc * 32 + lane];
                        acc1[c] = fmaf(wv.x, v.x, acc1[c]);
                        acc1[c] = fmaf(wv.y, v.y, acc1[c]);
                    }
                }
                const float b1 = b1w;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float v = ((SWN_ABL & 64) ? acc1[c] : warp_sum(acc1[c])) + b1;
                    if (lane < CORE_CL)
                        st_async_f32(x2_peer + (uint32_t)((c * n1p + D.b0 + warp) * 4), fmaxf(v, 0.f), mb_peer + (uint32_t)((L + 2) * 8));
                }
            }
            SWN_CPROF(9);
            SWN_HCLK(23);
            // ---- out_2 rows
#if SWN_HEAD2
            float2 wpre2[4];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc)
                wpre2[cc] = (warp < D.nc && cc * 64 < n1p) ? reinterpret_cast<const float2*>(sW2 + warp * n1p)[cc * 32 + lane] : make_float2(0.f, 0.f);
#endif
            if (gn_team && warp < D.nc) {
#pragma unroll
                for (int c = 0; c < NC; ++c) gn[c] = gnz[((istep & 1) * 16 + warp) * NC + c];
            }
            if (warp < D.nc) ok = mbar_wait(mb_x2, hpar, P, 530) && ok;
            SWN_HCLK(24);
            if (warp < D.nc) {
                const int r = warp;
                float acc2[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) acc2[c] = 0.f;
#if SWN_HEAD2
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (cc * 64 < n1p) {
                            const float2 v = reinterpret_cast<const float2*>(x2 + c * n1p)[cc * 32 + lane];
                            a = fmaf(wpre2[cc].x, v.x, a); b = fmaf(wpre2[cc].y, v.y, b);
                        }
                    }
                    acc2[c] = a + b;
                }
                for (int cc = 4; cc < (n1p >> 6); ++cc) {
#else
                for (int cc = 0; cc < ((SWN_ABL & 128) ? 1 : (n1p >> 6)); ++cc) {
#endif
                    const float2 wv = reinterpret_cast<const float2*>(sW2 + r * n1p)[cc * 32 + lane];
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        const float2 v = reinterpret_cast<const float2*>(x2 + c * n1p)[cc * 32 + lane];
                        acc2[c] = fmaf(wv.x, v.x, acc2[c]);
                        acc2[c] = fmaf(wv.y, v.y, acc2[c]);
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float o = ((SWN_ABL & 128) ? acc2[c] : warp_sum(acc2[c])) + b2w;
                    if (is_c) {
                        if (lane == 0) o2[c * 16 + r] = o;
                    } else {
                        const int row = D.c0 + r;
                        if (lane < CORE_CL) st_async_f32(gx_peer + (uint32_t)((c * 256 + row) * 4), o + gn[c], mb_peer + (uint32_t)((L + 3) * 8));
                        if (lane == 16 && P.head_out && c < P.B) P.head_out[((size_t)c * P.n_iter + istep) * n2 + row] = o;
                    }
                }
            }
            if (is_c) {
                __syncthreads();
                if (tid < NC) {
                    const int c = tid, bglob = min(c, P.B - 1);
                    float eps[1], smp[1];
                    eps[0] = P.noise ? P.noise[(size_t)bglob * P.n_out + istep] : laplace_eps(P.seed, P.b0 + bglob, istep);
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
                SWN_HCLK(25);
                ok = mbar_wait(mb_gx, hpar, P, 540) && ok;
                SWN_HCLK(26);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    float best = -CUDART_INF_F; int bi = 0x7fffffff;
                    for (int r0 = 0; r0 < Q; r0 += 64) {
                        const float2 v = reinterpret_cast<const float2*>(gx + c * 256 + r0)[lane];
                        const int i0 = r0 + 2 * lane;
                        if (v.x > best) { best = v.x; bi = i0; }
                        if (v.y > best) { best = v.y; bi = i0 + 1; }
                    }
                    // warp arg-max with the integer reduction unit: order-preserving key, redux.max, then the
                    // smallest index among the lanes that hold the maximum (first maximum wins, as torch.argmax)
                    uint32_t key = __float_as_uint(best);
                    key = (key & 0x80000000u) ? ~key : (key | 0x80000000u);
                    const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
                    bi = (int)__reduce_min_sync(0xffffffffu, key == kmax ? (uint32_t)bi : 0xffffffffu);
                    if (lane == 0) {
                        int fed = bi;
                        if (c < P.B) {
                            if (rank == 0) ((long long*)P.out)[(size_t)c * P.n_out + istep] = bi;
                            if (P.forced) fed = clamp_idx(((const long long*)P.forced)[(size_t)c * P.n_out + istep], Q);
                        }
                        hist[c * HIST + (istep % HIST)] = __int_as_float(fed);
                    }
                }
            }
            SWN_HCLK(27);
            if (__syncthreads_or(!ok)) { ok = false; break; }
            SWN_HCLK(28);
            // every consumer of this step's head vectors is past them: arm their next phase
            if (tid == (SWN_HEAD2 ? 32 * 15 : 0)) { mbar_arm(mb_x1, bytes_x1); mbar_arm(mb_x2, bytes_x2); if (!is_c) mbar_arm(mb_gx, bytes_gx); }
        }
        SWN_CPROF(10);
        SWN_HCLK(29);
        if (it + 1 < P.n_it) {
            input_publish(it + 1);
            SWN_HCLK(30);
            SWN_CPROF(11);
            feed(it + 1);
            cp_async_commit();
        }
        SWN_CPROF(12);
        fc = fn; jc = jn;
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
