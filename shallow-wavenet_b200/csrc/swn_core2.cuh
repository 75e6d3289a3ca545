// Core CTA of the cluster ("core + helpers") latency kernel, second generation.
//
// Same work split, data layout and exchange protocol as swn_core.cuh (16-CTA core cluster, critical taps in tensor
// memory, st.async + mbarrier hops, helper CTAs for the old taps): see the header comment there.  What changed, and
// why (profiles/r01_v2c ncu source view: a chain warp was BUSY ~1300 of the ~1600 cycles of a layer stage - every
// dependent instruction in its stream costs 4-6 cycles, so the instruction count of the wake -> publish path is the
// step time):
//   * role-specialised layer loops: chain warps (0-3), delay-1 / operand warps (4-7) and skip warps (8-15) each run
//     their own loop over the layers with running pointers, instead of one loop with per-role branches, a vote and
//     per-layer constant-bank lookups in front of every stage;
//   * the chain stage is branch free between the wake-up and the publish (the gate is evaluated by every lane, the
//     carry and the gate operands are fetched before the wait, the helper hand-off and the next taps come after the
//     publish, in the shadow of the hop);
//   * DSWNV tail: every warp finishes the arg-max over the 256 published scores itself and goes straight on to its
//     row of the next input vector - no CTA barrier and no shared-memory round trip between "sampled" and "next h0";
//   * one CTA barrier per time step instead of two; the table gathers of the input stage are waited for one step
//     later (they were exposed at the top of the step);
//   * the %globaltimer stamps are compiled in only with -DSWN_STAMPS=1 (tools/gpu_prof_core.py timelines).
#pragma once
#include "swn_core.cuh"

#ifndef SWN_STAMPS
#define SWN_STAMPS 0
#endif
#if SWN_STAMPS
#define SWN_STAMP(ev)                                                                              \
    do {                                                                                           \
        if (P.prof && threadIdx.x == 0 && it >= P.prof_start && it < P.prof_start + P.prof_count) \
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + (ev)] = gtime();  \
    } while (0)
// lane 0 of warps 0, 4, 8 and 15 (one per role, plus the last) stamps into slot ev + {0, 1, 2, 3}
#define SWN_WSTAMP(ev)                                                                             \
    do {                                                                                           \
        if (P.prof && (threadIdx.x & 31) == 0 && it >= P.prof_start && it < P.prof_start + P.prof_count) { \
            const int w_ = threadIdx.x >> 5;                                                       \
            const int s_ = w_ == 0 ? 0 : (w_ == 4 ? 1 : (w_ == 8 ? 2 : (w_ == 15 ? 3 : -1)));      \
            if (s_ >= 0) P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + (ev) + s_] = gtime(); \
        }                                                                                          \
    } while (0)
#else
#define SWN_STAMP(ev) do {} while (0)
#define SWN_WSTAMP(ev) do {} while (0)
#endif

#ifndef SWN_JOINT2
#define SWN_JOINT2 1            // two utterance columns: the chain warp evaluates both at once (one 8-slot reduction, one gate, one
#endif                          // publish per stage) instead of one after the other behind their own barriers

namespace swn {

// The 8 per-lane partial sums are those of 4 gate pairs x {z, candidate} (v[2p], v[2p+1]).  Afterwards every lane of
// pair group g = lane >> 3 holds both totals of pair g: five shuffle levels (16, 8 halve the pairs; 4, 2, 1 butterfly
// both sums), one level less than reduce8 + the pair exchange.
__device__ __forceinline__ void reduce_pairs(const float (&v)[8], int lane, float& qz, float& qc) {
    const bool b4 = lane & 16, b3 = lane & 8;
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // keep pairs {0,1} (b4 = 0) or {2,3} (b4 = 1)
        const float send = b4 ? v[i] : v[i + 4], keep = b4 ? v[i + 4] : v[i];
        u[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    float z, c;
    {
        const float sz = b3 ? u[0] : u[2], kz = b3 ? u[2] : u[0];
        const float sc = b3 ? u[1] : u[3], kc = b3 ? u[3] : u[1];
        z = kz + __shfl_xor_sync(0xffffffffu, sz, 8);
        c = kc + __shfl_xor_sync(0xffffffffu, sc, 8);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
        z += __shfl_xor_sync(0xffffffffu, z, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    qz = z; qc = c;
}

// Totals of acc[c][z|c] (c < 8 column slots) over the 32 lanes: afterwards every lane of column group lane >> 2 holds
// both totals of that column.  Three halving levels (16, 8, 4) and two butterflies: 18 shuffles for all columns.
template <int SEG>
__device__ __forceinline__ void reduce_columns(const float (&a)[SEG][2], int lane, float& qz, float& qc) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float u[4][2], v[2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int zc = 0; zc < 2; ++zc) {
            const float lo = i < SEG ? a[i < SEG ? i : 0][zc] : 0.f, hi = i + 4 < SEG ? a[i + 4 < SEG ? i + 4 : 0][zc] : 0.f;
            const float send = b4 ? lo : hi, keep = b4 ? hi : lo;
            u[i][zc] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int zc = 0; zc < 2; ++zc) {
            const float send = b3 ? u[i][zc] : u[i + 2][zc], keep = b3 ? u[i + 2][zc] : u[i][zc];
            v[i][zc] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
    float z, c;
    {
        const float sz = b2 ? v[0][0] : v[1][0], kz = b2 ? v[1][0] : v[0][0];
        const float sc = b2 ? v[0][1] : v[1][1], kc = b2 ? v[1][1] : v[0][1];
        z = kz + __shfl_xor_sync(0xffffffffu, sz, 4);
        c = kc + __shfl_xor_sync(0xffffffffu, sc, 4);
    }
#pragma unroll
    for (int o = 2; o > 0; o >>= 1) {
        z += __shfl_xor_sync(0xffffffffu, z, o);
        c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    qz = z; qc = c;
}

// Sleep on a local mbarrier until the phase with `parity` has completed.  The watchdog (abort flag in global memory,
// cycle budget) is consulted on a time basis, every ~50k cycles, so that no L2 load of the flag ever sits between two
// polls of a warp that is about to be woken (measured neutral against the poll-count based check: try_wait suspends
// long enough that the count rarely got there; kept because it bounds the worst case).
static __device__ __noinline__ bool mbar_wait_slow2(unsigned long long* b, uint32_t parity, const CoreParams& P, int where) {
    const long long t_start = clock64();
    long long next_check = t_start + 50000;
    for (;;) {
        if (mbar_try(b, parity)) return true;
        const long long now = clock64();
        if (now > next_check) {
            next_check = now + 50000;
            volatile int32_t* st = P.status;
            if (st[0] != 0) return false;
            if ((unsigned long long)(now - t_start) > (P.timeout_cycles >> 1)) {
                if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)parity; st[3] = blockIdx.x; }
                return false;
            }
        }
    }
}
__device__ __forceinline__ bool mbar_wait2(unsigned long long* b, uint32_t parity, const CoreParams& P, int where) {
    if (mbar_try(b, parity)) return true;
    if (mbar_try(b, parity)) return true;
    return mbar_wait_slow2(b, parity, P, where);
}

// Several core clusters side by side in one launch (MULTI, swn_abi.cu generate_core): cluster clid works on its own
// utterances (CoreParams::cl[clid]: conditioning table, noise, outputs, Philox utterance index, input / hand-off rings); the
// single-cluster instantiation reads the same fields of CoreParams itself, as before (identical code).
#define SWN_CU(field, legacy) (MULTI ? P.cl[clid].field : P.legacy)

template <int NCH, int BS, bool MULTI>
__device__ void core_cta2(const CoreParams& P, const CoreCta& D, float* smem, const int clid) {
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
    float* hx = smem + D.sm_hx;                            // [L+1][NC][HP] exchanged layer inputs
    float* x1 = smem + D.sm_x1;                            // [NC][Sp]  relu(sum of skips)
    float* x2 = smem + D.sm_x2;                            // [NC][n1p] relu(out_1)
    float* gx = smem + D.sm_gx;                            // [NC][256] Gumbel scores
    float4* gop = reinterpret_cast<float4*>(smem + D.sm_pre);   // [2][L][16][NC] gate operands, written one step ahead
    float* misc = smem + D.sm_misc;
    float* hist = misc;                                    // [NC][HIST] fed-back samples / indices
    float* lpb = hist + NC * HIST;                         // [NC][8]
    float* o2 = lpb + NC * 8;                              // [NC][16] out_2 rows (CSWNV)
    float4* lc = reinterpret_cast<float4*>(o2 + NC * 16);  // [L][16] per (layer, pair): conditioning const z/c, bias z/c
    float* cacc = reinterpret_cast<float*>(lc + L * 16);   // [8][16][NC]  input-stage partial sums by time slot
    float* cstage = cacc + 8 * 16 * NC;                    // [2][8][16][NC] gathered table entries in flight (cp.async)
    float* sfir = cstage + 2 * 8 * 16 * NC;                // CSWNV [8][2][16] FIR taps (cb, v) of this CTA's rows
    float* gnz = sfir + 8 * 2 * 16;                        // DSWNV [2][16][NC] Gumbel noise of this CTA's logit rows
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + D.sm_tmem_slot);
    // One mbarrier per (exchanged vector, utterance column): the columns of a slot run as independent pipelines, so that
    // the stage of one column is computed while the hop of another is in flight (a joint barrier per vector made two
    // columns cost 1.75 x one, profiles/r02k_core_columns_probe.txt).  [L+1][NC] h vectors, then [NC] x1, x2, scores.
    unsigned long long* mb = reinterpret_cast<unsigned long long*>(tmem_slot + 2);
    unsigned long long* mb_x1 = mb + (L + 1) * NC, *mb_x2 = mb + (L + 2) * NC, *mb_gx = mb + (L + 3) * NC;
    const uint32_t bytes_h = (uint32_t)(H * 4), bytes_x1 = (uint32_t)(S * 4), bytes_x2 = (uint32_t)(n1 * 4), bytes_gx = (uint32_t)(Q * 4);

    auto stage = [&](float* dst, const float* src, int n) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n / 4; i += SWN_NTHREADS) d4[i] = s4[i];
    };
    stage(sSkip, P.wbuf + D.skip_off, L * D.nskip * HP);
    stage(sW1, P.wbuf + D.w1_off, D.nb * Sp);
    stage(sW2, P.wbuf + D.w2_off, D.nc * n1p);
    if (!is_c) stage(sTin, P.wbuf + D.tin_off, Q * nip);
    {
        // exchanged vectors: padded elements stay 0 forever
        const int nw = (L + 1) * NC * HP + NC * Sp + NC * n1p + NC * 256;
        for (int i = tid; i < nw; i += SWN_NTHREADS) hx[i] = 0.f;
        for (int i = tid; i < 2 * L * 16 * NC; i += SWN_NTHREADS) gop[i] = make_float4(0.f, 0.f, 0.f, 0.f);
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
        for (int i = 0; i < (L + 4) * NC; ++i) mbar_init(mb + i, 1);
        mbar_fence_init();
        for (int i = 0; i < (L + 1) * NC; ++i) mbar_arm(mb + i, bytes_h);          // phase 0 of every vector
        for (int c = 0; c < NC; ++c) { mbar_arm(mb_x1 + c, bytes_x1); mbar_arm(mb_x2 + c, bytes_x2); mbar_arm(mb_gx + c, bytes_gx); }
    }
    if (warp == 0) tmem_alloc_all(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t taddr0 = tbase + ((uint32_t)(32 * quarter) << 16);
    // ---- tensor memory <- the core taps: warp q < 4 fills its lane quarter with the taps of gate pairs 4q..4q+3
    // (thread-private columns: tap c occupies columns [c*TCOLS, (c+1)*TCOLS), row rr = 2*pair + {z,c} at rr*FPT)
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
    if (SWN_CU(seedw, seedw)) {
        for (int e = tid; e < NC; e += SWN_NTHREADS) {
            const int bglob = min(e, SWN_CU(nb, B) - 1);
            const float v = is_c ? ((const float*)SWN_CU(seedw, seedw))[bglob]
                                 : __int_as_float(pos_mod(((const long long*)SWN_CU(seedw, seedw))[bglob], Q));
            hist[e * HIST + (HIST - 1)] = v;                       // generated-index -1
            if (is_c && P.lpc > 0) lpb[e * 8 + P.lpc - 1] = v;
        }
    }
    __syncthreads();
    cluster_sync_all();                                            // every CTA's buffers and barriers are ready

    const uint32_t peer = (uint32_t)(lane & (CORE_CL - 1));
    const uint32_t hx_peer = mapa(smem_u32(hx), peer), mb_peer = mapa(smem_u32(mb), peer);
    const int twoH = 2 * H;
    // ---- roles inside the CTA (one warp group = 4 warps = one warp per SM sub-partition):
    //   warps 0-3   "chain" warps: each evaluates the delay-0 tap of 4 gate pairs (8 rows), gates, publishes
    //   warps 4-7   delay-1 taps of the dilation-1 layers + next step's gate operands
    //   warps 8-15  two rows each of every out_skip
    // A lane of a chain / delay-1 warp belongs to pair-slot (lane >> 3) of its quarter; column = (lane & 7) for
    // per-column duties (NC <= 8).
    const int pslot = lane >> 3;                                   // 0..3
    const int plocal = 4 * quarter + pslot;                        // gate pair inside this CTA
    const bool pvalid = plocal < np;
    const int my_i = D.pair0 + (pvalid ? plocal : 0);              // row index in H (a valid row for idle lanes too)
    const int sub = lane & 7;                                      // lane inside the pair group
    const int colc = sub < NC ? sub : 0;                           // column this lane gates (idle lanes shadow column 0)
    const int skr0 = 2 * (warp - 8), skr1 = skr0 + 1;              // skip rows of warps 8..15
    // ---- per-warp head record in shared memory: what a head stage needs between its wake-up and its publish (remote byte
    // offsets relative to hx, biases).  It is fetched BEFORE the stage's wait; taken from the descriptor in global memory or
    // recomputed, the same values cost an L1/L2 round trip, constant-bank loads and spill reloads after the wake-up
    // (SASS of the first version of this file: LDG D.c0 -> IMAD -> LDL -> STAS behind the out_2 reduction).
    uint4* hrec_off = reinterpret_cast<uint4*>(gnz + 2 * 16 * NC);          // [16] {x1 row a, x1 row b, x2 row, score row}
    float4* hrec_c = reinterpret_cast<float4*>(hrec_off + 16);             // [16] {skip const a, skip const b, out_1 bias, out_2 bias}
    if (lane == 0) {
        const uint32_t o_x1 = (uint32_t)((x1 - hx) * 4), o_x2 = (uint32_t)((x2 - hx) * 4), o_gx = (uint32_t)((gx - hx) * 4);
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (warp >= 8) {
            o.x = o_x1 + (uint32_t)((D.skip0 + skr0) * 4); o.y = o_x1 + (uint32_t)((D.skip0 + skr1) * 4);
            if (skr0 < D.nskip) c4.x = __ldg(P.wbuf + D.skipc_off + skr0);
            if (skr1 < D.nskip) c4.y = __ldg(P.wbuf + D.skipc_off + skr1);
        }
        o.z = o_x2 + (uint32_t)((D.b0 + warp) * 4); o.w = o_gx + (uint32_t)((D.c0 + warp) * 4);
        if (warp < D.nb) c4.z = __ldg(P.wbuf + D.b1_off + warp);
        if (warp < D.nc) c4.w = __ldg(P.wbuf + D.b2_off + warp);
        hrec_off[warp] = o; hrec_c[warp] = c4;
    }
    const bool has_b = warp < D.nb, has_c = warp < D.nc;                   // this warp owns a row of out_1 / out_2
    const bool sk_v0 = skr0 < D.nskip, sk_v1 = skr1 < D.nskip;
    const int c0row = D.c0 + warp;                                         // logit row of this warp
    const int ws_stride = D.nskip * HP;                                    // floats between the out_skip blocks of two layers
    const uint32_t mbh = mb_peer + (uint32_t)((L + 1) * NC * 8);           // x1 barriers of the peer; x2 = + NC * 8, scores = + 2 * NC * 8
    // chain warps: where lane `sub` of a pair group publishes (CTAs sub and sub + 8), layer-0 addresses
    const uint32_t pub0 = mapa(smem_u32(hx), (uint32_t)sub) + (uint32_t)(my_i * 4);
    const uint32_t pub1 = mapa(smem_u32(hx), (uint32_t)sub + 8) + (uint32_t)(my_i * 4);
    const uint32_t pmb0 = mapa(smem_u32(mb), (uint32_t)sub), pmb1 = mapa(smem_u32(mb), (uint32_t)sub + 8);

    // ---- gate operands: everything the gate of (layer l, time tt) needs besides the newest-tap product is known one
    // step ahead: conditioning row (x gain + constant), biases, helper sums, delay-1 sums.  The delay-1 warps fetch the
    // global parts (L2-scope loads, issued before their own FMA work), wait for a late helper themselves, and leave
    //   { pz, (bz + rz + hz) * pz, pc, (bc + rc + hc) * pc }
    // in gop[tt & 1][l][pair][col]; the chain warp's stage then needs one LDS.128 and two FMAs before the gate.
    const bool pf_lane = warp >= 4 && warp < 8 && pvalid && sub < NC;
    const int bg_col = min(sub, SWN_CU(nb, B) - 1);
    const float* ptab_b = SWN_CU(P, P) + (size_t)bg_col * P.Tf * L * twoH + my_i;
    bool gok = true;
    auto gop_issue = [&](int l, int tt, const float* prow, float& prz, float& prc, unsigned long long& ha, unsigned long long& hb) {
        prz = 0.f; prc = 0.f; ha = 0ull; hb = 0ull;
        if (pf_lane && tt < P.n_it) {
            prz = __ldg(prow + l * twoH); prc = __ldg(prow + l * twoH + H);
            if (P.has_help[l] && tt >= P.hdmin[l]) {
                const unsigned long long* src = SWN_CU(hacc_ll, hacc_ll) + P.hacc_off[l] + ((tt & (P.hacc_nt[l] - 1)) * twoH + my_i) * NC + sub;
                ha = ld_ll(src); hb = ld_ll(src + H * NC);
            }
        }
    };
    auto gop_commit = [&](int l, int tt, float gain_tt, float prz, float prc, unsigned long long ha, unsigned long long hb,
                          float rz, float rc) {
        if (pf_lane && tt < P.n_it) {
            float hz = 0.f, hc = 0.f;
            if (P.has_help[l] && tt >= P.hdmin[l]) {
                const uint32_t tg = (uint32_t)tt + 1u;
                uint32_t a = (uint32_t)ha, b = (uint32_t)hb;
                if ((uint32_t)(ha >> 32) != tg || (uint32_t)(hb >> 32) != tg) {      // the helper is late: wait for it here
                    const unsigned long long* src = SWN_CU(hacc_ll, hacc_ll) + P.hacc_off[l] + ((tt & (P.hacc_nt[l] - 1)) * twoH + my_i) * NC + sub;
                    gok = wait_gword(src, tg, a, P, 400 + l) && gok;
                    gok = wait_gword(src + H * NC, tg, b, P, 400 + l) && gok;
                }
                hz = __uint_as_float(a); hc = __uint_as_float(b);
            }
            const float4 c4 = lc[l * 16 + plocal];
            const float pz = fmaf(gain_tt, prz, c4.x), pc = fmaf(gain_tt, prc, c4.y);
            gop[(((tt & 1) * L + l) * 16 + plocal) * NC + sub] = make_float4(pz, (c4.z + rz + hz) * pz, pc, (c4.w + rc + hc) * pc);
        }
    };

    // ---- Gumbel noise of the CTA's 16 logit rows, drawn one step ahead by (16 * NC) lanes of warps 8-11 while the
    // layers run (a Philox block and two logarithms per row)
    const bool gn_team = !is_c && D.nc <= 16;
    auto draw_gumbel = [&](int isn) {
        const int idx = (warp - 8) * 32 + lane;
        if (gn_team && warp >= 8 && idx < 16 * NC) {
            const int row = idx & 15, c = idx >> 4, bglob = min(c, SWN_CU(nb, B) - 1);
            if (row < D.nc) {
                const float q = SWN_CU(noise, noise) ? __ldg(SWN_CU(noise, noise) + ((size_t)bglob * P.n_out + isn) * Q + D.c0 + row)
                                        : softmax_q(P.seed, SWN_CU(b0, b0) + bglob, isn, D.c0 + row);
                gnz[((isn & 1) * 16 + row) * NC + c] = -logf(q);
            }
        }
    };

    // ---- input stage: h0 rows of this CTA (row = i0 + warp), published into every peer's hx[0].
    // Scatter form again: the sample fed at time tau adds tap (K-1-d) into the partial sum of time tau+d
    // (cacc ring).  Lane d-1 handles delay d >= 1; delay 0 (the newest tap) is added at publish time from
    // shared memory.  DSWNV: the older taps are gathered from the L2-resident table with cp.async and folded in right
    // before the NEXT publish (they are needed one step later at the earliest, so their latency is never exposed).
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
    // the sample of time tau is known: CSWNV value / DSWNV class index per column in `fed` (float bits)
    auto feed = [&](int tau, const float (&fed)[NC]) {
        if (!has_in || lane >= K - 1) return;
        const int d = lane + 1, j = K - 1 - d;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            if (is_c) {
                cacc[(((tau + d) & 7) * 16 + warp) * NC + c] += sfir[(j * 2) * 16 + warp] + sfir[(j * 2 + 1) * 16 + warp] * fed[c];
            } else {
                cp_async4(&cstage[(((tau & 1) * 8 + lane) * 16 + warp) * NC + c],
                          P.in_tab + ((size_t)j * Q + __float_as_int(fed[c])) * H + in_row);
            }
        }
    };
    auto input_publish_col = [&](int itn, int c, float fv) {       // h0 of time itn, column c (its own sample is `fv`)
        if (!has_in) return;
        float* slot = cacc + ((itn & 7) * 16 + warp) * NC + c;
        float acc = in_b + *slot;
        if (is_c) acc += sfir[((K - 1) * 2) * 16 + warp] + sfir[((K - 1) * 2 + 1) * 16 + warp] * fv;
        else acc += sTin[__float_as_int(fv) * nip + warp];
        __syncwarp();
        if (lane == 0) *slot = 0.f;                                  // the slot is reused 8 steps later
        // softsign through rcp.approx (1 ulp) instead of the IEEE division sequence: it sits between the arg-max and the publish
        const float v = acc * rcp_approx(1.f + fabsf(acc));
        if (lane < CORE_CL) st_async_f32(hx_peer + (uint32_t)((c * HP + in_row) * 4), v, mb_peer + (uint32_t)(c * 8));
        if (lane == 16 && P.has_help[0])
            st_ll(SWN_CU(hin_ll, hin_ll) + P.hin_off[0] + ((itn & (P.hslots[0] - 1)) * NC + c) * H + in_row, v, (uint32_t)itn + 1u);
    };
    auto input_publish = [&](int itn, const float (&fed)[NC]) {
#pragma unroll
        for (int c = 0; c < NC; ++c) input_publish_col(itn, c, fed[c]);
    };
    // what the input stage sees as the sample of time tau (generated index n = tau - R - 1; padding before the seed)
    auto fed_from_hist = [&](int tau, float (&fed)[NC]) {
        const int n = tau - P.R - 1;
#pragma unroll
        for (int c = 0; c < NC; ++c)
            fed[c] = (n >= -1) ? hist[c * HIST + ((n + HIST) & (HIST - 1))] : (is_c ? 0.f : __int_as_float(Q / 2));
    };

    __syncthreads();                                               // lc[] is complete
    for (int l = 0; l < L; ++l) {
        float prz, prc; unsigned long long ha, hb;
        gop_issue(l, 0, ptab_b, prz, prc, ha, hb);
        gop_commit(l, 0, __ldg(P.gain), prz, prc, ha, hb, 0.f, 0.f);
    }
    int fc = 0, jc = 0;                                            // frame / phase of the current time step
    {
        float fed0[NC];
        fed_from_hist(0, fed0);
        input_publish(0, fed0);                                    // time 0 sees only its own (padding) sample
        feed(0, fed0);
        cp_async_commit();
    }
    if (P.it0 == 0) draw_gumbel(0);                                // (made visible by the first step's CTA barrier)
    bool ok = true;
    for (int it = 0; it < P.n_it; ++it) {
        const bool active = it >= P.it0;
        const uint32_t par = (uint32_t)it & 1u;
        const int istep = it - P.it0;
        const uint32_t hpar = (uint32_t)istep & 1u;                // phase of the head vectors (advance only when active)
        int fn = fc, jn = jc;                                      // frame / phase of time it+1
        if (it + 1 > P.R && ++jn == P.U) { jn = 0; ++fn; }
        // the one CTA barrier of the step: gop[] of this step (written by warps 4-7 during the previous one), the input
        // ring and the noise drawn ahead become visible; a failed wait anywhere ends the kernel here
        if (__syncthreads_or(!(ok && gok))) break;
        SWN_STAMP(0);
        if (warp < 4) {
            // ======================= chain warps: delay-0 tap of 4 gate pairs per layer, gate, publish =======================
            float w[TCOLS];
            tmem_ldw<TCOLS>(taddr0 + P.tap_base[0] * TCOLS, w);
            const float4* gp = gop + ((par * L) * 16 + plocal) * NC;            // gate operands of layer 0, column 0
            const float* hxl = hx;                                               // input vector of layer l, column 0
            uint32_t d0 = pub0 + (uint32_t)(NC * HP * 4), d1 = pub1 + (uint32_t)(NC * HP * 4);   // where hx[l+1] rows go
            uint32_t m0 = pmb0 + (uint32_t)(NC * 8), m1 = pmb1 + (uint32_t)(NC * 8);
            unsigned long long* bar = mb;
            for (int l = 0; l < L; ++l) {
                if constexpr (NC == 2 && SWN_JOINT2 != 0) {
                    // ---- two columns at once: 4 pairs x 2 columns = 8 (pair, column) slots, 4 lanes each after the reduction.
                    // One reduction (18 shuffles for both columns instead of 2 x 13), one gate, one publish; the 16 independent
                    // FMA chains fill the issue gaps a single column leaves.  12.1 -> 10.4 us per step for two utterances.
                    // (The head stages stay one column after the other: evaluated jointly they were SLOWER, 11.2 us - there the
                    // early publish of the first column is worth more than the interleaving.)
                    const int slot = lane >> 2, sp = slot >> 1, scol = slot & 1, s4 = lane & 3;
                    const int prow = 4 * quarter + sp;
                    const bool sv = prow < np;
                    const float4 g4 = gp[((sv ? prow : plocal) - plocal) * NC + scol];          // (gp points at this lane's own pair: rebase)
                    ok = mbar_wait2(bar + 0, par, P, 500 + l) && ok;
                    ok = mbar_wait2(bar + 1, par, P, 500 + l) && ok;
#if SWN_STAMPS
                    if (P.prof && tid == 0 && l < 6 && it >= P.prof_start && it < P.prof_start + P.prof_count)
                        P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + 1 + l] = gtime();
#endif
                    float2 h0[NCH], h1[NCH];
                    load_frag<NCH>(h0, hxl, lane);
                    load_frag<NCH>(h1, hxl + HP, lane);
                    const float carry = hxl[scol * HP + D.pair0 + (sv ? prow : 0)];
                    tmem_wait_ld();
                    float a8[8][2];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        float a0 = 0.f, b0 = 0.f, a1 = 0.f, b1 = 0.f;
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            a0 = fmaf(w[rr * FPT + 2 * cc], h0[cc].x, a0); b0 = fmaf(w[rr * FPT + 2 * cc + 1], h0[cc].y, b0);
                            a1 = fmaf(w[rr * FPT + 2 * cc], h1[cc].x, a1); b1 = fmaf(w[rr * FPT + 2 * cc + 1], h1[cc].y, b1);
                        }
                        a8[(rr >> 1) * 2 + 0][rr & 1] = a0 + b0;                                  // slot = pair * 2 + column
                        a8[(rr >> 1) * 2 + 1][rr & 1] = a1 + b1;
                    }
                    float qz, qc;
                    reduce_columns<8>(a8, lane, qz, qc);
                    const float hn = gate_out(fmaf(qz, g4.x, g4.y), fmaf(qc, g4.z, g4.w), carry);
                    if (sv) {
                        const uint32_t off = (uint32_t)((((l + 1) * NC + scol) * HP + D.pair0 + prow) * 4);
                        const uint32_t mo = (uint32_t)(((l + 1) * NC + scol) * 8);
                        const uint32_t hx0 = smem_u32(hx), mb0 = smem_u32(mb);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            st_async_f32(mapa(hx0 + off, (uint32_t)(s4 + 4 * k)), hn, mapa(mb0 + mo, (uint32_t)(s4 + 4 * k)));
                        if (l + 1 < L && P.has_help[l + 1] && s4 == 0)
                            st_ll(SWN_CU(hin_ll, hin_ll) + P.hin_off[l + 1] + ((it & (P.hslots[l + 1] - 1)) * NC + scol) * H + D.pair0 + prow, hn, (uint32_t)it + 1u);
                    }
                } else {
                // the columns one after the other, each behind its own barrier: while the hop of column c is in flight
                // the stage of column c + 1 is computed
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float4 g4 = gp[c];                                     // (before the wait: in the shadow of the hop)
                    ok = mbar_wait2(bar + c, par, P, 500 + l) && ok;
#if SWN_STAMPS
                    if (c == 0 && P.prof && tid == 0 && l < 6 && it >= P.prof_start && it < P.prof_start + P.prof_count)
                        P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + 1 + l] = gtime();
#endif
                    float2 h[NCH];
                    load_frag<NCH>(h, hxl + c * HP, lane);
                    const float carry = hxl[c * HP + my_i];
                    if (c == 0) tmem_wait_ld();
                    float v8[8];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        float a = 0.f, b = 0.f;
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            a = fmaf(w[rr * FPT + 2 * cc], h[cc].x, a);
                            b = fmaf(w[rr * FPT + 2 * cc + 1], h[cc].y, b);
                        }
                        v8[rr] = a + b;
                    }
                    float qz, qc;
                    reduce_pairs(v8, lane, qz, qc);
                    // ---- gate (cswnv_shift1.py:281-285) by every lane of the pair group (no branch).
                    // (1-z) tanh(gc) + z carry over one common denominator: two ex2, one rcp on the dependent chain
                    //   z = 1/(1+ez), tanh = sign(gc) (1-ec)/(1+ec)  ->  [ez sign(gc) (1-ec) + carry (1+ec)] / [(1+ez)(1+ec)]
                    // ez is capped at 2^43 (gz >= -30: z differs from 0 by < 1e-13 either way) so the product stays finite.
                    const float hn = gate_out(fmaf(qz, g4.x, g4.y), fmaf(qc, g4.z, g4.w), carry);
                    // ---- fan out: in every pair group, lane sub writes CTAs sub and sub+8
                    if (pvalid) {
                        st_async_f32(d0 + (uint32_t)(c * HP * 4), hn, m0 + (uint32_t)(c * 8));
                        st_async_f32(d1 + (uint32_t)(c * HP * 4), hn, m1 + (uint32_t)(c * 8));
                    }
                    // ---- behind the publish (in the shadow of the hop): helper hand-off
                    if (l + 1 < L && P.has_help[l + 1] && pvalid && sub == 0)
                        st_ll(SWN_CU(hin_ll, hin_ll) + P.hin_off[l + 1] + ((it & (P.hslots[l + 1] - 1)) * NC + c) * H + my_i, hn, (uint32_t)it + 1u);
                }
                }
                // next taps, running pointers
                if (l + 1 < L) tmem_ldw<TCOLS>(taddr0 + P.tap_base[l + 1] * TCOLS, w);
                gp += 16 * NC; hxl += NC * HP; bar += NC;
                d0 += (uint32_t)(NC * HP * 4); d1 += (uint32_t)(NC * HP * 4); m0 += (uint32_t)(NC * 8); m1 += (uint32_t)(NC * 8);
            }
        } else if (warp < 8) {
            // ======================= delay-1 taps of the dilation-1 layers + next step's gate operands =======================
            float w[TCOLS];
            const float* prow_next = ptab_b + (size_t)fn * L * twoH;
            const float gq_next = __ldg(P.gain + jn);
            for (int l = 0; l < L; ++l) {
#pragma unroll
                for (int c = 0; c < NC; ++c) ok = mbar_wait2(mb + l * NC + c, par, P, 600 + l) && ok;
                // next phase (next time step) of every layer barrier, armed in one go once the last of them has completed
                if (l == L - 1 && warp == 4 && lane == 0)
                    for (int i = 0; i < L * NC; ++i) mbar_arm(mb + i, bytes_h);
                float prz, prc, rz1 = 0.f, rc1 = 0.f; unsigned long long ha, hb;
                gop_issue(l, it + 1, prow_next, prz, prc, ha, hb);
                if (P.dil[l] == 1 && K >= 2) {
                    tmem_ldw<TCOLS>(taddr0 + (P.tap_base[l] + 1) * TCOLS, w);
                    float2 h[NC][NCH];
#pragma unroll
                    for (int c = 0; c < NC; ++c) load_frag<NCH>(h[c], hx + (l * NC + c) * HP, lane);
                    tmem_wait_ld();
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
                        float mz, mc;
                        reduce_pairs(v8, lane, mz, mc);
                        if (sub == c) { rz1 = mz; rc1 = mc; }
                    }
                }
                gop_commit(l, it + 1, gq_next, prz, prc, ha, hb, rz1, rc1);
                // Reconverge before the next layer.  Only the operand lanes wait for a late helper in gop_commit; without this
                // the other lanes of the warp ran on into the next dilation-1 layer's tcgen05.ld (.sync.aligned) and its
                // full-mask shuffles with a partial warp, and the delay-1 sums of that layer came out wrong - silently, and
                // only when a helper was late (profiles/r03_core2_late_helper_repro.txt).
                __syncwarp();
            }
        } else {
            // ======================= skip warps: two rows of every out_skip; the head's first vector =======================
            if (it + 1 >= P.it0 && it + 1 < P.n_it) draw_gumbel(it + 1 - P.it0);      // noise of the next step
            float sk0[NC], sk1[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) { sk0[c] = 0.f; sk1[c] = 0.f; }
            const bool v0 = sk_v0, v1 = sk_v1;
            const int nsk = ws_stride;
            const float* ws_base0 = sSkip + (size_t)(v0 ? skr0 : 0) * HP;
            const float* ws_base1 = sSkip + (size_t)(v1 ? skr1 : 0) * HP;
            for (int l = 0; l < L; ++l) {
#pragma unroll
                for (int c = 0; c < NC; ++c) ok = mbar_wait2(mb + l * NC + c, par, P, 700 + l) && ok;
                if (active && l >= 1) {                                          // out_skip[l-1] on the input of layer l
                    float2 h[NC][NCH], ws0[NCH], ws1[NCH];
#pragma unroll
                    for (int c = 0; c < NC; ++c) load_frag<NCH>(h[c], hx + (l * NC + c) * HP, lane);
                    load_frag<NCH>(ws0, ws_base0 + (size_t)(l - 1) * nsk, lane);
                    load_frag<NCH>(ws1, ws_base1 + (size_t)(l - 1) * nsk, lane);
#pragma unroll
                    for (int c = 0; c < NC; ++c) { sk0[c] += dot_frag<NCH>(ws0, h[c]); sk1[c] += dot_frag<NCH>(ws1, h[c]); }
                }
            }
            if (active) {
                // ---- out_skip[L-1] on the output of the last layer; relu(sum of skips) -> x1 on every CTA.  The weight
                // rows are in registers before the vector lands.
                float2 ws0[NCH], ws1[NCH];
                load_frag<NCH>(ws0, ws_base0 + (size_t)(L - 1) * nsk, lane);
                load_frag<NCH>(ws1, ws_base1 + (size_t)(L - 1) * nsk, lane);
                const uint4 ro = hrec_off[warp];
                const float4 rc = hrec_c[warp];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    ok = mbar_wait2(mb + L * NC + c, par, P, 510) && ok;
                    float2 hl[NCH];
                    load_frag<NCH>(hl, hx + (L * NC + c) * HP, lane);
                    float a = sk0[c] + dot_frag<NCH>(ws0, hl), b = sk1[c] + dot_frag<NCH>(ws1, hl);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        b += __shfl_xor_sync(0xffffffffu, b, o);
                    }
                    if (lane < CORE_CL) {
                        if (v0) st_async_f32(hx_peer + ro.x + (uint32_t)(c * Sp * 4), fmaxf(a + rc.x, 0.f), mbh + (uint32_t)(c * 8));
                        if (v1) st_async_f32(hx_peer + ro.y + (uint32_t)(c * Sp * 4), fmaxf(b + rc.y, 0.f), mbh + (uint32_t)(c * 8));
                    }
                }
            }
        }
        SWN_STAMP(7);
        SWN_WSTAMP(20);
        if (it + 1 < P.n_it) {
            cp_async_wait_all();                                                 // gathers issued one step ago: fold them in before
            __syncwarp();                                                        // the head, off the sampled -> next h0 path
            fold_stage(it);
            __syncwarp();
        }
        // the output of the last layer (consumed by every warp, so that its barrier advances every step)
        if (warp < 8 || !active) {
#pragma unroll
            for (int c = 0; c < NC; ++c) ok = mbar_wait2(mb + L * NC + c, par, P, 511) && ok;
        }
        if (tid == 128)
            for (int c = 0; c < NC; ++c) mbar_arm(mb + L * NC + c, bytes_h);
        float fed[NC];                                                           // what is fed back as the sample of time it+1
        if (active) {
            SWN_STAMP(8);
            // ---- out_1 row of this warp (relu) -> x2 on every CTA
            if (has_b) {
                // the weight row (up to 256 columns; longer rows continue below) and the head record are in registers
                // before the vector lands
                const float* wrow = sW1 + (size_t)warp * Sp;
                float2 wpre[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wpre[cc] = (cc * 64 < Sp) ? reinterpret_cast<const float2*>(wrow)[cc * 32 + lane] : make_float2(0.f, 0.f);
                const uint32_t dst = hx_peer + hrec_off[warp].z;
                const float bias = hrec_c[warp].z;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    ok = mbar_wait2(mb_x1 + c, hpar, P, 520) && ok;
                    if (c == 0) SWN_STAMP(13);
                    if (tid == 0) mbar_arm(mb_x1 + c, bytes_x1);                  // (its next phase only starts a step later)
                    const float2* xv = reinterpret_cast<const float2*>(x1 + (size_t)c * Sp);
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (cc * 64 < Sp) {
                            const float2 v = xv[cc * 32 + lane];
                            a = fmaf(wpre[cc].x, v.x, a); b = fmaf(wpre[cc].y, v.y, b);
                        }
                    }
                    for (int cc = 4; cc < (Sp >> 6); ++cc) {
                        const float2 wv = reinterpret_cast<const float2*>(wrow)[cc * 32 + lane], v = xv[cc * 32 + lane];
                        a = fmaf(wv.x, v.x, a); b = fmaf(wv.y, v.y, b);
                    }
                    const float v = warp_sum(a + b) + bias;
                    if (lane < CORE_CL)
                        st_async_f32(dst + (uint32_t)(c * n1p * 4), fmaxf(v, 0.f), mbh + (uint32_t)((NC + c) * 8));
                }
            } else if (tid == 0) {
                for (int c = 0; c < NC; ++c) {
                    ok = mbar_wait2(mb_x1 + c, hpar, P, 521) && ok;
                    mbar_arm(mb_x1 + c, bytes_x1);
                }
            }
            SWN_STAMP(9);
            // ---- out_2 rows
            if (has_c) {
                const int r = warp;
                const float* wrow = sW2 + r * n1p;
                float2 wpre2[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wpre2[cc] = (cc * 64 < n1p) ? reinterpret_cast<const float2*>(wrow)[cc * 32 + lane] : make_float2(0.f, 0.f);
                const uint32_t dst = hx_peer + hrec_off[warp].w;
                const float bias = hrec_c[warp].w;
                float gn[NC];                                                    // the Gumbel noise drawn a step ahead
#pragma unroll
                for (int c = 0; c < NC; ++c) gn[c] = 0.f;
                if (!is_c) {
                    if (gn_team) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) gn[c] = gnz[((istep & 1) * 16 + r) * NC + c];
                    } else {
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            const int bglob = min(c, SWN_CU(nb, B) - 1);
                            const float q = SWN_CU(noise, noise) ? __ldg(SWN_CU(noise, noise) + ((size_t)bglob * P.n_out + istep) * Q + c0row)
                                                    : softmax_q(P.seed, SWN_CU(b0, b0) + bglob, istep, c0row);
                            gn[c] = -logf(q);
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    ok = mbar_wait2(mb_x2 + c, hpar, P, 530) && ok;
                    if (c == 0) SWN_STAMP(15);
                    if (tid == 0) mbar_arm(mb_x2 + c, bytes_x2);
                    const float2* xv = reinterpret_cast<const float2*>(x2 + c * n1p);
                    float a = 0.f, b = 0.f;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (cc * 64 < n1p) {
                            const float2 v = xv[cc * 32 + lane];
                            a = fmaf(wpre2[cc].x, v.x, a); b = fmaf(wpre2[cc].y, v.y, b);
                        }
                    }
                    for (int cc = 4; cc < (n1p >> 6); ++cc) {
                        const float2 wv = reinterpret_cast<const float2*>(wrow)[cc * 32 + lane], v = xv[cc * 32 + lane];
                        a = fmaf(wv.x, v.x, a); b = fmaf(wv.y, v.y, b);
                    }
                    const float o = warp_sum(a + b) + bias;
                    if (is_c) {
                        if (lane == 0) o2[c * 16 + r] = o;
                    } else {
                        if (lane < CORE_CL) st_async_f32(dst + (uint32_t)(c * 256 * 4), o + gn[c], mbh + (uint32_t)((2 * NC + c) * 8));
                        if (lane == 16 && SWN_CU(head_out, head_out) && c < SWN_CU(nb, B)) SWN_CU(head_out, head_out)[((size_t)c * P.n_iter + istep) * n2 + c0row] = o;
                    }
                }
            } else if (tid == 0) {
                for (int c = 0; c < NC; ++c) {
                    ok = mbar_wait2(mb_x2 + c, hpar, P, 531) && ok;
                    mbar_arm(mb_x2 + c, bytes_x2);
                }
            }
            SWN_STAMP(16);
            if (is_c) {
                __syncthreads();
                if (tid < NC) {
                    const int c = tid, bglob = min(c, SWN_CU(nb, B) - 1);
                    float eps[1], smp[1];
                    eps[0] = SWN_CU(noise, noise) ? SWN_CU(noise, noise)[(size_t)bglob * P.n_out + istep] : laplace_eps(P.seed, SWN_CU(b0, b0) + bglob, istep);
                    const float* forced = SWN_CU(forced, forced) ? (const float*)SWN_CU(forced, forced) + (size_t)bglob * P.n_out + istep : nullptr;
                    laplace_step<1>(P.lpc, o2 + c * 16, lpb + c * 8, eps, forced, smp);
                    hist[c * HIST + (istep % HIST)] = forced ? forced[0] : smp[0];
                    if (rank == 0 && c < SWN_CU(nb, B)) {
                        ((float*)SWN_CU(out, out))[(size_t)c * P.n_out + istep] = smp[0];
                        if (SWN_CU(head_out, head_out))
                            for (int r = 0; r < n2; ++r) SWN_CU(head_out, head_out)[((size_t)c * P.n_iter + istep) * n2 + r] = o2[c * 16 + r];
                    }
                }
                __syncthreads();
                fed_from_hist(it + 1, fed);
            } else {
                // ---- arg-max over the 256 published Gumbel scores (first maximum wins), by EVERY warp for itself: the
                // warp goes straight on to its row of the next input vector, no CTA barrier, no shared-memory round trip
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    ok = mbar_wait2(mb_gx + c, hpar, P, 540) && ok;
                    if (c == 0) SWN_STAMP(17);
                    if (tid == 0) mbar_arm(mb_gx + c, bytes_gx);
                    // the lane's 8 scores as a compare tree (depth 3) instead of a chain of 8; ties keep the lower index
                    float bv[4]; int bx[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const float2 v = (q4 * 64 < Q) ? reinterpret_cast<const float2*>(gx + c * 256 + q4 * 64)[lane]
                                                       : make_float2(-CUDART_INF_F, -CUDART_INF_F);
                        const int i0 = q4 * 64 + 2 * lane;
                        bv[q4] = v.y > v.x ? v.y : v.x; bx[q4] = v.y > v.x ? i0 + 1 : i0;
                    }
                    const bool s01 = bv[1] > bv[0], s23 = bv[3] > bv[2];
                    const float b01 = s01 ? bv[1] : bv[0], b23 = s23 ? bv[3] : bv[2];
                    const int i01 = s01 ? bx[1] : bx[0], i23 = s23 ? bx[3] : bx[2];
                    const float best = b23 > b01 ? b23 : b01;
                    int bi = b23 > b01 ? i23 : i01;
                    // warp arg-max with the integer reduction unit: order-preserving key, redux.max, then the
                    // smallest index among the lanes that hold the maximum (first maximum wins, as torch.argmax)
                    uint32_t key = __float_as_uint(best);
                    key = (key & 0x80000000u) ? ~key : (key | 0x80000000u);
                    const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
                    bi = (int)__reduce_min_sync(0xffffffffu, key == kmax ? (uint32_t)bi : 0xffffffffu);
                    int f = bi;
                    if (c < SWN_CU(nb, B)) {
                        if (rank == 0 && tid == 0) ((long long*)SWN_CU(out, out))[(size_t)c * P.n_out + istep] = bi;
                        if (SWN_CU(forced, forced)) f = clamp_idx(__ldg((const long long*)SWN_CU(forced, forced) + (size_t)c * P.n_out + istep), Q);
                    }
                    fed[c] = __int_as_float(f);
                    if (it + 1 < P.n_it) input_publish_col(it + 1, c, fed[c]);   // straight on to this column's next input row
                }
            }
            SWN_STAMP(10);
            SWN_WSTAMP(24);
        } else {
            fed_from_hist(it + 1, fed);                                          // warm-up: the padded seed
        }
        if (it + 1 < P.n_it) {
            if (is_c || !active) input_publish(it + 1, fed);
            SWN_STAMP(11);
            feed(it + 1, fed);
            cp_async_commit();
        }
        SWN_STAMP(12);
        fc = fn; jc = jn;
    }
    // ---- teardown
    cp_async_wait_all();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_free_all(tbase);
}

#undef SWN_CU

template <int NCH, int BS, bool MULTI>
__global__ void __cluster_dims__(CORE_CL, 1, 1) __launch_bounds__(SWN_NTHREADS, 1)
swn_core_kernel2(const __grid_constant__ CoreParams P) {
    extern __shared__ __align__(16) float smem[];
    const int clid = (int)cluster_id_x(), ncl = MULTI ? P.n_cl : 1;     // clusters 0 .. ncl-1 are core clusters
    if (clid < ncl) {
        core_cta2<NCH, BS, MULTI>(P, P.cdesc[cluster_ctarank()], smem, clid);
    } else {
        const int hidx = (int)blockIdx.x - CORE_CL * ncl;
        if (hidx < P.n_help && P.hdesc[hidx].npairs > 0) helper_cta<NCH, BS, MULTI>(P, P.hdesc[hidx], smem);
    }
}

}  // namespace swn
