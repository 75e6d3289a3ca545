// Cluster ("core + helpers") latency kernel for CSWNV with several samples per loop step (seg > 1): the reference's
// primary model (src/nets/cswnv_shift1.py:348-402) at batch 1.
//
// Same principle as swn_core2.cuh - the dependent chain of a loop step stays inside ONE 16-CTA cluster, vectors hop
// through distributed shared memory (st.async + mbarrier complete_tx), the critical taps live in tensor memory, the
// old taps are evaluated by helper CTAs on the rest of the chip - re-cut for SEG time columns per step:
//
//   * a step carries SEG consecutive time columns through every layer.  The taps whose delay is shorter than SEG
//     (delta * dilation < SEG: all of them for a dilation-1 layer at seg 5, only delay 0 for the others) couple
//     columns of the SAME step (or reach back into the previous one): they are the core's, in GATHER form over a
//     double-buffered (step parity) copy of every layer's input columns.  Taps with delay >= SEG only ever feed later
//     steps: they are the helpers', in scatter form with one step of slack;
//   * tensor memory holds the core taps thread-privately: a thread of TMEM lane quarter q keeps 2*NCH rows (NCH gate
//     pairs) x 2*NCH weights per tap = 4*NCH^2 columns (36 at H=192: 14 taps of the recipe model = 504 of 512 columns);
//   * the columns of a stage are spread over TWO warps per quarter (warps q and q+4 read the same TMEM lanes): even
//     columns / odd columns.  Each warp owns its columns from the first FMA to the publish, so a stage has no
//     intra-CTA hand-off; the two warps share a scheduler and fill each other's issue gaps;
//   * the last layer only evaluates the last column (the head reads nothing else and no layer follows);
//   * warps 8-15: the out_skip rows of the last column, and the gate operands (conditioning rows x gain + constant,
//     bias + helper sums) of layer l+2 while layer l runs - late enough for the helpers' one-step slack, early enough
//     that the L2 latency is never on the chain.  A local mbarrier per layer tells the chain warps the operands are in;
//   * head as in swn_core2.cuh: x1 = relu(sum of skips) -> out_1 -> out_2 (every CTA all rows) -> Laplace / LPC sampler
//     (cswnv_shift1.py:368-391) by one thread per CTA (bit-identical in all 16) -> the SEG new input columns.
#pragma once
#include "swn_core.cuh"
#include "swn_core2.cuh"

namespace swn {

template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t a, float (&r)[N]) {
    if constexpr (N % 16 == 0) tmem_ldw<N>(a, r);
    else tmem_ld<N>(a, r);
}

// ------------------------------------------------------------------ helper CTA (time-granular targets)
// Owns the taps of one layer with delay >= SEG.  The input column of relative time tau = it*SEG + c adds W_j . h into
// the partial sum of time tau + delay_j; when the smallest delay has been applied the sum is complete and goes to the
// core (hacc word of time T, tag T + 1).
template <int NCH, int SEG>
__device__ void helper_cta_seg(const CoreParams& P, const HelpCta& D, float* smem) {
    constexpr int HP = NCH * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = D.layer, np = D.npairs, H = P.H, RL = D.ring_len, nt = D.ntaps;
    float* sW = smem + D.sm_w;
    {
        const float4* src = reinterpret_cast<const float4*>(P.wbuf + D.w_off);
        float4* dst = reinterpret_cast<float4*>(sW);
        for (int i = tid; i < nt * np * 2 * HP / 4; i += SWN_NTHREADS) dst[i] = src[i];
    }
    float* hb = smem + D.sm_h;                                 // [SEG][HP]
    for (int i = tid; i < SEG * HP; i += SWN_NTHREADS) hb[i] = 0.f;
    float* ring = smem + D.sm_ring;                            // [RL][np][2]
    for (int i = tid; i < RL * np * 2; i += SWN_NTHREADS) ring[i] = 0.f;
    __syncthreads();
    const unsigned long long* in_ll = P.hin_ll + P.hin_off[l];
    unsigned long long* acc_ll = P.hacc_ll + P.hacc_off[l];
    const int slot_mask = P.hslots[l] - 1, nt_mask = P.hacc_nt[l] - 1, rl_mask = RL - 1;   // all powers of two
    const int my_j = lane / SEG, my_c = lane - my_j * SEG;      // lane <-> (tap j, column c)
    const bool my_live = my_j < nt;
    const int my_delay = my_live ? D.delay[my_j] : 0;
    for (int it = 0; it < P.n_it; ++it) {
        const uint32_t tag = (uint32_t)it + 1u;
        bool ok = true;
        {
            const unsigned long long* src = in_ll + (size_t)(it & slot_mask) * SEG * H;
            for (int i = tid; i < SEG * H; i += SWN_NTHREADS) {
                unsigned long long v = ld_ll(src + i);
                uint32_t bits = (uint32_t)v;
                if ((uint32_t)(v >> 32) != tag) ok = wait_gword(src + i, tag, bits, P, 300 + l) && ok;
                const int c = i / H, k = i - c * H;
                hb[c * HP + k] = __uint_as_float(bits);
            }
        }
        if (__syncthreads_or(!ok)) return;
        float2 h[SEG][NCH];
#pragma unroll
        for (int c = 0; c < SEG; ++c) load_frag<NCH>(h[c], hb + c * HP, lane);
        for (int pi = warp; pi < np; pi += SWN_NWARPS) {
            float mz = 0.f, mc = 0.f;
#pragma unroll
            for (int j = 0; j < CORE_KMAX; ++j) {
                if (j < nt) {
                    const float* wr = sW + ((size_t)(j * np + pi) * 2) * HP;
                    float2 wz[NCH], wc[NCH];
                    load_frag<NCH>(wz, wr, lane);
                    load_frag<NCH>(wc, wr + HP, lane);
#pragma unroll
                    for (int c = 0; c < SEG; ++c) {
                        const float a = warp_sum(dot_frag<NCH>(wz, h[c])), b = warp_sum(dot_frag<NCH>(wc, h[c]));
                        if (lane == j * SEG + c) { mz = a; mc = b; }
                    }
                }
            }
            // ring update, one tap per round (within a round the SEG columns hit SEG different times); the smallest
            // delay comes last and finishes its target times
            for (int j = nt - 1; j >= 0; --j) {
                if (my_live && my_j == j) {
                    const int T = it * SEG + my_c + my_delay;
                    float* r0 = ring + ((T & rl_mask) * np + pi) * 2;
                    const float a = r0[0] + mz, b = r0[1] + mc;
                    if (j == 0) {
                        r0[0] = 0.f; r0[1] = 0.f;
                        unsigned long long* dst = acc_ll + (size_t)(T & nt_mask) * 2 * H;
                        st_ll(dst + D.pair0 + pi, a, (uint32_t)T + 1u);
                        st_ll(dst + H + D.pair0 + pi, b, (uint32_t)T + 1u);
                    } else {
                        r0[0] = a; r0[1] = b;
                    }
                }
                __syncwarp();
            }
        }
        __syncthreads();                                        // hb is rewritten by the next step's loads
    }
}

// ------------------------------------------------------------------ core CTA
template <int NCH, int SEG>
__device__ void core_cta_seg(const CoreParams& P, const CoreCta& D, float* smem) {
    constexpr int HP = NCH * 64, NPW = NCH, FPT = 2 * NCH, ROWS = 2 * NPW, TCOLS = ROWS * FPT;
    constexpr int HIST = 32;
    constexpr int NCI = (SEG + 1) / 2;                          // columns per team (team 0: even, team 1: odd)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = P.H, S = P.S, Sp = P.Sp, K = P.K, L = P.L, n1 = P.n1, n1p = P.n1p, n2 = P.n2;
    const uint32_t rank = cluster_ctarank();
    const int np = D.npairs;
    const int quarter = warp & 3, team = warp >> 2;
    const int twoH = 2 * H;

    // ---- shared memory carve-up (identical offsets in all 16 CTAs)
    float* sSkip = smem + D.sm_skip;                       // [L][nskip][HP]
    float* sW1 = smem + D.sm_w1;                           // [nb][Sp]
    float* sW2 = smem + D.sm_w2;                           // [n2][n1p]
    float* hx = smem + D.sm_hx;                            // [2][L+1][SEG][HP] layer inputs, by step parity
    float* x1 = smem + D.sm_x1;                            // [Sp]
    float* x2 = smem + D.sm_x2;                            // [n1p]
    float4* gop = reinterpret_cast<float4*>(smem + D.sm_pre);   // [L][16][SEG] gate operands {pz, bz*pz, pc, bc*pc}
    float* misc = smem + D.sm_misc;
    float* hist = misc;                                    // [HIST] fed-back samples
    float* lpb = hist + HIST;                              // [8]
    float* o2 = lpb + 8;                                   // [16] out_2 rows
    float* epsb = o2 + 16;                                 // [8] the step's uniform noise
    float4* lc = reinterpret_cast<float4*>(epsb + 8);      // [L][16] per (layer, pair): conditioning const z/c, bias z/c
    float* sfir = reinterpret_cast<float*>(lc + L * 16);   // [8][2][16] FIR taps (cb, v) of this CTA's input rows
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + D.sm_tmem_slot);
    unsigned long long* mb = reinterpret_cast<unsigned long long*>(tmem_slot + 2);   // [L+1] layer inputs (+ last layer's output)
    unsigned long long* mb_x1 = mb + (L + 1), *mb_x2 = mb + (L + 2);
    unsigned long long* gb = mb + (L + 3);                 // [L] "gate operands of layer l are in" (8 arrivals: warps 8-15)
    const uint32_t bytes_h = (uint32_t)(SEG * H * 4), bytes_last = (uint32_t)(H * 4);
    const uint32_t bytes_x1 = (uint32_t)(S * 4), bytes_x2 = (uint32_t)(n1 * 4);
    const int LS = (L + 1) * SEG * HP;                     // floats per parity copy of hx

    auto stage = [&](float* dst, const float* src, int n) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n / 4; i += SWN_NTHREADS) d4[i] = s4[i];
    };
    stage(sSkip, P.wbuf + D.skip_off, L * D.nskip * HP);
    stage(sW1, P.wbuf + D.w1_off, D.nb * Sp);
    stage(sW2, P.wbuf + D.w2_off, D.nc * n1p);
    for (int i = tid; i < 2 * LS + Sp + n1p; i += SWN_NTHREADS) hx[i] = 0.f;    // hx, x1, x2 are contiguous; pads stay 0
    for (int i = tid; i < L * 16 * SEG; i += SWN_NTHREADS) gop[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < HIST + 8 + 16 + 8; i += SWN_NTHREADS) misc[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < L * 16; i += SWN_NTHREADS) {
        const int l = i >> 4, wv = i & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (wv < np) {
            v.x = P.pconst[l * twoH + D.pair0 + wv]; v.y = P.pconst[l * twoH + H + D.pair0 + wv];
            v.z = P.wbuf[D.bias_off + (l * np + wv) * 2]; v.w = P.wbuf[D.bias_off + (l * np + wv) * 2 + 1];
        }
        lc[i] = v;
    }
    for (int i = tid; i < 8 * 2 * 16; i += SWN_NTHREADS) {
        const int j = i >> 5, w2 = (i >> 4) & 1, r = i & 15;
        sfir[i] = (j < K && r < D.ni) ? P.in_tab[(j * 2 + w2) * H + D.i0 + r] : 0.f;
    }
    if (tid == 0) {
        for (int i = 0; i < L + 3; ++i) mbar_init(mb + i, 1);
        for (int i = 0; i < L; ++i) mbar_init(gb + i, 8);
        mbar_fence_init();
        for (int i = 0; i < L; ++i) mbar_arm(mb + i, bytes_h);                    // phase 0 of every vector
        mbar_arm(mb + L, bytes_last); mbar_arm(mb_x1, bytes_x1); mbar_arm(mb_x2, bytes_x2);
    }
    if (warp == 0) tmem_alloc_all(tmem_slot);
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    const uint32_t tbase = *tmem_slot;
    const uint32_t taddr0 = tbase + ((uint32_t)(32 * quarter) << 16);
    // ---- tensor memory <- the core taps: warp q < 4 fills its lane quarter with the taps of gate pairs NPW*q .. +NPW-1
    // (thread-private columns: tap c occupies columns [c*TCOLS, (c+1)*TCOLS), row rr = 2*pair + {z,c} at rr*FPT)
    if (warp < 4) {
        for (int c = 0; c < P.n_core_taps; ++c) {
#pragma unroll
            for (int rr = 0; rr < ROWS; ++rr) {
                const int pl = NPW * quarter + (rr >> 1);
                float wv[FPT];
#pragma unroll
                for (int cc = 0; cc < NCH; ++cc) {
                    float2 a = make_float2(0.f, 0.f);
                    if (pl < np)
                        a = reinterpret_cast<const float2*>(P.wbuf + D.tap_off + ((size_t)(c * np + pl) * 2 + (rr & 1)) * HP)[cc * 32 + lane];
                    wv[2 * cc] = a.x; wv[2 * cc + 1] = a.y;
                }
#pragma unroll
                for (int q = 0; q < FPT; q += 2) {             // (FPT is even; two columns at a time)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr0 + c * TCOLS + rr * FPT + q),
                                 "r"(__float_as_uint(wv[q])), "r"(__float_as_uint(wv[q + 1])) : "memory");
                }
            }
        }
        tmem_wait_st();
    }
    if (P.seedw) {
        // a_full = pad(R) ++ seed(SEG) ++ generated: the seed occupies generated-indices -SEG..-1; lp_x_buffer = last
        // lpc samples of the padded seed (cswnv_shift1.py:317-319)
        if (tid < SEG) {
            const float v = ((const float*)P.seedw)[tid];
            hist[(tid - SEG + HIST) & (HIST - 1)] = v;
            const int mm = P.lpc - SEG + tid;
            if (mm >= 0) lpb[mm] = v;
        }
    }
    __syncthreads();
    cluster_sync_all();                                            // every CTA's buffers and barriers are ready

    const uint32_t peer = (uint32_t)(lane & (CORE_CL - 1));
    const uint32_t hx_peer = mapa(smem_u32(hx), peer), mb_peer = mapa(smem_u32(mb), peer);
    const int pslot = lane >> 3;                                   // pair slot inside the quarter (valid below NPW)
    const int plocal = NPW * quarter + pslot;                      // gate pair inside this CTA
    const bool pvalid = pslot < NPW && plocal < np;
    const int my_i = D.pair0 + (pvalid ? plocal : 0);              // row index in H (a valid row for idle lanes too)
    const int sub = lane & 7;                                      // lane inside the pair group: publishes to CTAs sub, sub+8
    const uint32_t pub0 = mapa(smem_u32(hx), (uint32_t)sub) + (uint32_t)(my_i * 4);
    const uint32_t pub1 = mapa(smem_u32(hx), (uint32_t)sub + 8) + (uint32_t)(my_i * 4);
    const uint32_t pmb0 = mapa(smem_u32(mb), (uint32_t)sub), pmb1 = mapa(smem_u32(mb), (uint32_t)sub + 8);
    const int skr0 = 2 * (warp - 8), skr1 = skr0 + 1;              // skip rows of warps 8..15
    const bool sk_v0 = warp >= 8 && skr0 < D.nskip, sk_v1 = warp >= 8 && skr1 < D.nskip;
    const bool has_b = warp < D.nb, has_c = warp < D.nc;
    const int in_row = D.i0 + warp;
    const bool has_in = warp < D.ni;
    const float in_b = has_in ? __ldg(P.in_bias + in_row) : 0.f;
    bool ok = true;

    // ---- gate operands of (step s, layer l), by the 256 lanes of warps 8-15: lane <-> (pair, column)
    auto prep = [&](int s, int l) {
        const int idx = (warp - 8) * 32 + lane;
        if (s < P.n_it && idx < np * SEG) {
            const int p = idx / SEG, c = idx - p * SEG;
            const int row = D.pair0 + p;
            const int t = P.t_first + s * SEG + c;
            const float4 c4 = lc[l * 16 + p];
            float az = c4.x, ac = c4.y;
            int f = 0, j = 0;
            if (t >= P.R) { f = (t - P.R) / P.U; j = (t - P.R) - f * P.U; }
            const float* pl = P.P + (size_t)l * SEG * twoH + row;
            int tau = t;
#pragma unroll
            for (int sp = 0; sp < SEG; ++sp) {
                const float g = __ldg(P.gain + j);
                const float* pr = pl + ((size_t)f * L * SEG + sp) * twoH;
                az = fmaf(g, __ldg(pr), az);
                ac = fmaf(g, __ldg(pr + H), ac);
                ++tau;
                if (tau > P.R && ++j == P.U) { j = 0; ++f; }
            }
            float hz = 0.f, hc = 0.f;
            const int T = s * SEG + c;
            if (P.has_help[l] && T >= P.hdmin[l]) {
                const unsigned long long* src = P.hacc_ll + P.hacc_off[l] + (size_t)(T & (P.hacc_nt[l] - 1)) * twoH + row;
                const uint32_t tg = (uint32_t)T + 1u;
                const unsigned long long va = ld_ll(src), vb = ld_ll(src + H);
                uint32_t a = (uint32_t)va, b = (uint32_t)vb;
                if ((uint32_t)(va >> 32) != tg) ok = wait_gword(src, tg, a, P, 400 + l) && ok;       // the helper is late
                if ((uint32_t)(vb >> 32) != tg) ok = wait_gword(src + H, tg, b, P, 400 + l) && ok;
                hz = __uint_as_float(a); hc = __uint_as_float(b);
            }
            gop[(l * 16 + p) * SEG + c] = make_float4(az, (c4.z + hz) * az, ac, (c4.w + hc) * ac);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gb + l);
    };

    // ---- input stage: the SEG columns of block itn (cswnv_shift1.py:352), rows of this CTA, published into every
    // peer's hx[parity of itn][0].  Gather form over the sample history (tap validity = the conv's zero padding).
    auto input_stage = [&](int itn) {
        if (!has_in) return;
        const int c = lane < SEG ? lane : SEG - 1;
        const int t = P.t_first + itn * SEG + c;
        float v = 0.f;
        if (t >= 0) {
            float acc = in_b;
            for (int j = 0; j < K; ++j) {
                const int tau = t - (K - 1) + j;                             // hidden time of the tap
                if (tau < 0) continue;                                       // conv zero padding
                const int n = tau - P.R - 1;                                 // newest sample seen by hidden time tau
                const float a = (n >= -SEG) ? hist[(n + HIST) & (HIST - 1)] : 0.f;
                acc += sfir[(j * 2) * 16 + warp] + sfir[(j * 2 + 1) * 16 + warp] * a;
            }
            v = acc * rcp_approx(1.f + fabsf(acc));
        }
        const uint32_t base = hx_peer + (uint32_t)((((itn & 1) * (L + 1)) * SEG * HP + in_row) * 4);
#pragma unroll
        for (int cc = 0; cc < SEG; ++cc) {
            const float vc = __shfl_sync(0xffffffffu, v, cc);
            if (lane < CORE_CL) st_async_f32(base + (uint32_t)(cc * HP * 4), vc, mb_peer);
            if (lane == 16 && P.has_help[0])
                st_ll(P.hin_ll + P.hin_off[0] + ((size_t)(itn & (P.hslots[0] - 1)) * SEG + cc) * H + in_row, vc, (uint32_t)itn + 1u);
        }
    };

    // ---- prologue: operands of the first two layers of step 0, inputs of block 0 (padding / seed)
    if (warp >= 8) { prep(0, 0); if (L > 1) prep(0, 1); }
    input_stage(0);

    for (int it = 0; it < P.n_it; ++it) {
        const bool active = it >= P.it0;
        const uint32_t par = (uint32_t)it & 1u;
        const int istep = it - P.it0;
        const uint32_t hpar = (uint32_t)istep & 1u;
        // the one CTA barrier of the step; a failed wait anywhere ends the kernel here
        if (__syncthreads_or(!ok)) break;
        SWN_STAMP(0);
        const float* hcur = hx + par * LS;                                   // this step's columns
        const float* hprev = hx + (par ^ 1u) * LS;                           // the previous step's
        if (warp < 8) {
            // ======================= column teams: core taps (gather form), gate, publish =======================
            float w[TCOLS];
            tmem_ldn<TCOLS>(taddr0 + P.tap_base[0] * TCOLS, w);
            for (int l = 0; l < L; ++l) {
                const int nco = P.ncore[l], dl = P.dil[l];
                const bool lastl = l + 1 == L;
                ok = mbar_wait2(gb + l, par, P, 480 + l) && ok;
                ok = mbar_wait2(mb + l, par, P, 500 + l) && ok;
#if SWN_STAMPS
                if (P.prof && tid == 0 && l < 6 && it >= P.prof_start && it < P.prof_start + P.prof_count)
                    P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + 1 + l] = gtime();
#endif
                float acc[NCI][ROWS];
#pragma unroll
                for (int ci = 0; ci < NCI; ++ci)
#pragma unroll
                    for (int rr = 0; rr < ROWS; ++rr) acc[ci][rr] = 0.f;
                for (int e = 0; e < nco; ++e) {
                    if (e > 0) tmem_ldn<TCOLS>(taddr0 + (P.tap_base[l] + e) * TCOLS, w);
                    tmem_wait_ld();
                    const int dly = e * dl;
#pragma unroll
                    for (int ci = 0; ci < NCI; ++ci) {
                        const int c = 2 * ci + team;
                        if (c < SEG && (!lastl || c == SEG - 1)) {
                            const int src = c - dly;
                            const float* hp = (src >= 0 ? hcur + src * HP : hprev + (SEG + src) * HP) + l * SEG * HP;
                            float2 h[NCH];
                            load_frag<NCH>(h, hp, lane);
#pragma unroll
                            for (int rr = 0; rr < ROWS; ++rr)
#pragma unroll
                                for (int cc = 0; cc < NCH; ++cc) {
                                    acc[ci][rr] = fmaf(w[rr * FPT + 2 * cc], h[cc].x, acc[ci][rr]);
                                    acc[ci][rr] = fmaf(w[rr * FPT + 2 * cc + 1], h[cc].y, acc[ci][rr]);
                                }
                        }
                    }
                }
                if (!lastl) tmem_ldn<TCOLS>(taddr0 + P.tap_base[l + 1] * TCOLS, w);    // next layer's first tap, in the shadow
#pragma unroll
                for (int ci = 0; ci < NCI; ++ci) {
                    const int c = 2 * ci + team;
                    if (c < SEG && (!lastl || c == SEG - 1)) {
                        float v8[8];
#pragma unroll
                        for (int rr = 0; rr < 8; ++rr) v8[rr] = rr < ROWS ? acc[ci][rr] : 0.f;
                        float qz, qc;
                        reduce_pairs(v8, lane, qz, qc);
                        const float4 g4 = gop[(l * 16 + (pvalid ? plocal : 0)) * SEG + c];
                        const float carry = hcur[(l * SEG + c) * HP + my_i];
                        float hn = gate_out(fmaf(qz, g4.x, g4.y), fmaf(qc, g4.z, g4.w), carry);
                        if (it == 0 && P.t_first + c < 0) hn = 0.f;              // alignment columns before time 0
                        if (pvalid) {
                            // the last layer's output is a single column (slot 0 of hx[.][L])
                            const uint32_t off = (uint32_t)(((par * (L + 1) + (l + 1)) * SEG + (lastl ? 0 : c)) * HP * 4);
                            st_async_f32(pub0 + off, hn, pmb0 + (uint32_t)((l + 1) * 8));
                            st_async_f32(pub1 + off, hn, pmb1 + (uint32_t)((l + 1) * 8));
                            if (!lastl && P.has_help[l + 1] && sub == 0)
                                st_ll(P.hin_ll + P.hin_off[l + 1] + ((size_t)(it & (P.hslots[l + 1] - 1)) * SEG + c) * H + my_i, hn,
                                      (uint32_t)it + 1u);
                        }
                    }
                }
            }
            tmem_wait_ld();
        } else {
            // ======================= warps 8-15: noise, gate operands two layers ahead, out_skip rows of the last column =======================
            if (active && warp == 15 && lane < SEG) {
                const long long n = (long long)istep * SEG + lane;
                epsb[lane] = P.noise ? __ldg(P.noise + n) : laplace_eps(P.seed, P.b0, n);
            }
            float sk0 = 0.f, sk1 = 0.f;
            const float* ws0 = sSkip + (size_t)(sk_v0 ? skr0 : 0) * HP;
            const float* ws1 = sSkip + (size_t)(sk_v1 ? skr1 : 0) * HP;
            const int nsk = D.nskip * HP;
            for (int l = 0; l < L; ++l) {
                if (l + 2 < L) prep(it, l + 2);
                ok = mbar_wait2(mb + l, par, P, 700 + l) && ok;
                if (active && l >= 1) {                                          // out_skip[l-1] on the last column of layer l's input
                    float2 h[NCH], a0[NCH], a1[NCH];
                    load_frag<NCH>(h, hcur + (l * SEG + SEG - 1) * HP, lane);
                    load_frag<NCH>(a0, ws0 + (size_t)(l - 1) * nsk, lane);
                    load_frag<NCH>(a1, ws1 + (size_t)(l - 1) * nsk, lane);
                    sk0 += dot_frag<NCH>(a0, h); sk1 += dot_frag<NCH>(a1, h);
                }
            }
            {
                float2 a0[NCH], a1[NCH];
                load_frag<NCH>(a0, ws0 + (size_t)(L - 1) * nsk, lane);
                load_frag<NCH>(a1, ws1 + (size_t)(L - 1) * nsk, lane);
                const float c0 = sk_v0 ? __ldg(P.wbuf + D.skipc_off + skr0) : 0.f, c1 = sk_v1 ? __ldg(P.wbuf + D.skipc_off + skr1) : 0.f;
                ok = mbar_wait2(mb + L, par, P, 510) && ok;
                if (active) {
                    float2 hl[NCH];
                    load_frag<NCH>(hl, hcur + (size_t)L * SEG * HP, lane);
                    float a = sk0 + dot_frag<NCH>(a0, hl), b = sk1 + dot_frag<NCH>(a1, hl);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        b += __shfl_xor_sync(0xffffffffu, b, o);
                    }
                    if (lane < CORE_CL) {
                        const uint32_t dst = hx_peer + (uint32_t)((2 * LS + D.skip0) * 4);
                        if (sk_v0) st_async_f32(dst + (uint32_t)(skr0 * 4), fmaxf(a + c0, 0.f), mb_peer + (uint32_t)((L + 1) * 8));
                        if (sk_v1) st_async_f32(dst + (uint32_t)(skr1 * 4), fmaxf(b + c1, 0.f), mb_peer + (uint32_t)((L + 1) * 8));
                    }
                }
                // every layer barrier of this step has completed (this thread waited for each): arm the next phase
                if (warp == 8 && lane == 0) {
                    for (int i = 0; i < L; ++i) mbar_arm(mb + i, bytes_h);
                    mbar_arm(mb + L, bytes_last);
                }
            }
        }
        // This thread has read the previous step's columns for the last time.  Nobody overwrites them (the next block's
        // input columns go into the same parity copy) before every thread of the cluster has said so: arrive here, wait
        // right before the next block's inputs are published - behind the head the wait is free; in the warm-up, where no
        // head couples the CTAs, it is what keeps a fast CTA from running a step ahead.
        cluster_arrive();
        SWN_STAMP(7);
        if (active) {
            // ---- out_1 row of this warp (relu) -> x2 on every CTA
            if (has_b) {
                const float* wrow = sW1 + (size_t)warp * Sp;
                float2 wpre[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wpre[cc] = (cc * 64 < Sp) ? reinterpret_cast<const float2*>(wrow)[cc * 32 + lane] : make_float2(0.f, 0.f);
                const float bias = __ldg(P.wbuf + D.b1_off + warp);
                ok = mbar_wait2(mb_x1, hpar, P, 520) && ok;
                if (tid == 0) mbar_arm(mb_x1, bytes_x1);
                const float2* xv = reinterpret_cast<const float2*>(x1);
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (cc * 64 < Sp) { const float2 v = xv[cc * 32 + lane]; a = fmaf(wpre[cc].x, v.x, a); b = fmaf(wpre[cc].y, v.y, b); }
                for (int cc = 4; cc < (Sp >> 6); ++cc) {
                    const float2 wv = reinterpret_cast<const float2*>(wrow)[cc * 32 + lane], v = xv[cc * 32 + lane];
                    a = fmaf(wv.x, v.x, a); b = fmaf(wv.y, v.y, b);
                }
                const float v = warp_sum(a + b) + bias;
                if (lane < CORE_CL)
                    st_async_f32(hx_peer + (uint32_t)((2 * LS + Sp + D.b0 + warp) * 4), fmaxf(v, 0.f), mb_peer + (uint32_t)((L + 2) * 8));
            } else if (tid == 0) {
                ok = mbar_wait2(mb_x1, hpar, P, 521) && ok;
                mbar_arm(mb_x1, bytes_x1);
            }
            SWN_STAMP(9);
            // ---- out_2 rows (every CTA computes all of them)
            if (has_c) {
                const float* wrow = sW2 + warp * n1p;
                float2 wpre2[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wpre2[cc] = (cc * 64 < n1p) ? reinterpret_cast<const float2*>(wrow)[cc * 32 + lane] : make_float2(0.f, 0.f);
                const float bias = __ldg(P.wbuf + D.b2_off + warp);
                ok = mbar_wait2(mb_x2, hpar, P, 530) && ok;
                if (tid == 0) mbar_arm(mb_x2, bytes_x2);
                const float2* xv = reinterpret_cast<const float2*>(x2);
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    if (cc * 64 < n1p) { const float2 v = xv[cc * 32 + lane]; a = fmaf(wpre2[cc].x, v.x, a); b = fmaf(wpre2[cc].y, v.y, b); }
                for (int cc = 4; cc < (n1p >> 6); ++cc) {
                    const float2 wv = reinterpret_cast<const float2*>(wrow)[cc * 32 + lane], v = xv[cc * 32 + lane];
                    a = fmaf(wv.x, v.x, a); b = fmaf(wv.y, v.y, b);
                }
                const float o = warp_sum(a + b) + bias;
                if (lane == 0) o2[warp] = o;
            } else if (tid == 0) {
                ok = mbar_wait2(mb_x2, hpar, P, 531) && ok;
                mbar_arm(mb_x2, bytes_x2);
            }
            SWN_STAMP(16);
            __syncthreads();
            if (tid == 0) {
                // Laplace / LPC sampler (cswnv_shift1.py:368-391), identically in every CTA
                const long long nbase = (long long)istep * SEG;
                float eps[SEG], smp[SEG];
#pragma unroll
                for (int j = 0; j < SEG; ++j) eps[j] = epsb[j];
                const float* forced = P.forced ? (const float*)P.forced + nbase : nullptr;
                laplace_step<SEG>(P.lpc, o2, lpb, eps, forced, smp);
#pragma unroll
                for (int j = 0; j < SEG; ++j) {
                    hist[(int)((nbase + j) & (HIST - 1))] = forced ? forced[j] : smp[j];
                    if (rank == 0) ((float*)P.out)[nbase + j] = smp[j];
                }
                if (rank == 0 && P.head_out)
                    for (int r = 0; r < n2; ++r) P.head_out[(size_t)istep * n2 + r] = o2[r];
            }
            __syncthreads();
            SWN_STAMP(10);
        }
        cluster_wait();
        if (it + 1 < P.n_it) {
            input_stage(it + 1);
            SWN_STAMP(11);
            if (warp >= 8) { prep(it + 1, 0); if (L > 1) prep(it + 1, 1); }
        }
        SWN_STAMP(12);
    }
    // ---- teardown
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_free_all(tbase);
}

template <int NCH, int SEG>
__global__ void __cluster_dims__(CORE_CL, 1, 1) __launch_bounds__(SWN_NTHREADS, 1)
swn_coreseg_kernel(const __grid_constant__ CoreParams P) {
    extern __shared__ __align__(16) float smem[];
    if (cluster_id_x() == 0) {
        core_cta_seg<NCH, SEG>(P, P.cdesc[cluster_ctarank()], smem);
    } else {
        const int hidx = (int)blockIdx.x - CORE_CL;
        if (hidx < P.n_help && P.hdesc[hidx].npairs > 0) helper_cta_seg<NCH, SEG>(P, P.hdesc[hidx], smem);
    }
}

}  // namespace swn
