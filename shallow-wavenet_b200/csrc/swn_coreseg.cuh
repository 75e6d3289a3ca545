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
//   * a chain warp owns ONE gate pair for all SEG columns (warps q, q+4, q+8 read the same TMEM lane quarter, each
//     its own pair's columns): every tap is read from tensor memory once per stage (TMEM reads are 64 B/clk per SM:
//     the first cut, columns split over two warps, read every tap twice and spent 4 us per step on it), the totals
//     of all columns come out of ONE transposed reduction (18 shuffles), one gate, one publish - no intra-CTA hand-off;
//   * the last layer only evaluates the last column (the head reads nothing else and no layer follows);
//     (Measured negative, removed: requesting them before the reduction, +0.25 us per step; requesting taps 1.. of a
//     dilation-1 layer a whole stage ahead while the single-tap layer before it runs, +0.28 us; requesting the first layer's taps
//     at the end of the step instead of behind the warp's last own stage, +0.14 us - profiles/r02z_coreseg_ab_timings.txt.)
//   * warps 12-13: the out_skip rows of the last column (eight each) and x1;  warps 14-15: nothing but the gate operands
//     (conditioning rows x gain + constant, bias + helper sums), a step ahead - operands and their "are in" mbarriers are
//     double-buffered by step parity, so these warps may block on L2 and on a late helper without holding up x1 (with
//     all four warps doing both, the blocking fetch put x1 1.2 us late once the stages had become faster);
//   * head as in swn_core2.cuh: x1 = relu(sum of skips) -> out_1 -> out_2 (every CTA all rows) -> Laplace / LPC sampler
//     (cswnv_shift1.py:368-391) by every warp for itself (bit-identical everywhere) -> the SEG new input columns.
#pragma once
#include "swn_core.cuh"
#include "swn_core2.cuh"

#ifndef SEG_FAT
#define SEG_FAT 1               // layers with a single core tap: ONE chain warp per tensor-memory quarter owns all its pairs
#endif
#if SWN_STAMPS
#define SWN_TSTAMP(ev, cond)                                                                        \
    do {                                                                                            \
        if (P.prof && (cond) && it >= P.prof_start && it < P.prof_start + P.prof_count)             \
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + (ev)] = gtime();   \
    } while (0)
#else
#define SWN_TSTAMP(ev, cond) do {} while (0)
#endif

namespace swn {

template <int N>
__device__ __forceinline__ void tmem_ldn(uint32_t a, float (&r)[N]) {
    if constexpr (N % 16 == 0) tmem_ldw<N>(a, r);
    else tmem_ld<N>(a, r);
}
// N floats (multiple of 4) into r[O .. O+N) of a larger register array
template <int N, int O, int M>
__device__ __forceinline__ void tmem_ld_at(uint32_t a, float (&r)[M]) {
    static_assert(O + N <= M && N % 4 == 0, "slice of the tap registers");
#pragma unroll
    for (int i = 0; i < N; i += 4) tmem_ld4(a + i, &r[O + i]);
}

// dot_frag with packed fp32 multiply-adds (fma.rn.f32x2: the same two accumulator chains as dot_frag, x and y parts, in half
// the issue slots - the helper's mat-vecs are issue-bound).  Bit-identical to dot_frag.
#ifndef SEG_HELP_FFMA2
#define SEG_HELP_FFMA2 1
#endif
template <int NCH>
__device__ __forceinline__ float dot_frag_packed(const float2 (&w)[NCH], const float2 (&h)[NCH]) {
#if SEG_HELP_FFMA2
    unsigned long long acc = 0ull;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        unsigned long long a, b;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(w[c].x), "f"(w[c].y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(h[c].x), "f"(h[c].y));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
    }
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc));
    return x + y;
#else
    return dot_frag<NCH>(w, h);
#endif
}

// Totals of v[slot] (32 slots) over the 32 lanes: afterwards lane i holds the total of slot i.  Five halving levels, 31 shuffles.
__device__ __forceinline__ float reduce_slots32(float (&v)[32], int lane) {
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int k = 0; k < w; ++k) {
            const float send = up ? v[k] : v[k + w];
            const float keep = up ? v[k + w] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return v[0];
}

#ifndef SEG_HELP_ZC_JOINT
#define SEG_HELP_ZC_JOINT 1
#endif
// the same for two slot arrays at once (two independent chains per level)
__device__ __forceinline__ void reduce_slots32x2(float (&a)[32], float (&b)[32], int lane, float& ra, float& rb) {
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int k = 0; k < w; ++k) {
            const float sa = up ? a[k] : a[k + w], ka = up ? a[k + w] : a[k];
            const float sb = up ? b[k] : b[k + w], kb = up ? b[k + w] : b[k];
            a[k] = ka + __shfl_xor_sync(0xffffffffu, sa, w);
            b[k] = kb + __shfl_xor_sync(0xffffffffu, sb, w);
        }
    }
    ra = a[0]; rb = b[0];
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
    float* hb_all = smem + D.sm_h;                             // [n_cl][SEG][HP]
    for (int i = tid; i < P.n_cl * SEG * HP; i += SWN_NTHREADS) hb_all[i] = 0.f;
    float* ring_all = smem + D.sm_ring;                        // [n_cl][RL][np][2]
    for (int i = tid; i < P.n_cl * RL * np * 2; i += SWN_NTHREADS) ring_all[i] = 0.f;
    __syncthreads();
    const int slot_mask = P.hslots[l] - 1, nt_mask = P.hacc_nt[l] - 1;   // powers of two (RL need not be: the rings of the
                                                                          // long dilations are what fills the helpers' shared memory)
    const int my_j = lane / SEG, my_c = lane - my_j * SEG;      // lane <-> (tap j, column c)
    const bool my_live = my_j < nt;
    const int my_delay = my_live ? D.delay[my_j] : 0;
    constexpr int WPT = (SEG * HP + SWN_NTHREADS - 1) / SWN_NTHREADS;   // input words per thread and cluster
    const int ncl = P.n_cl, n_words = SEG * H;
    // The core clusters of the launch (one utterance each) share the weights and take turns; input rings, partial sums and
    // hand-off rings are per cluster.  The input words of the NEXT turn are requested before this turn's mat-vecs (their L2
    // round trip hides behind the arithmetic when the helper is the late one; a word that has not been written yet is polled
    // as before).  Waiting for all clusters' inputs before computing any was measured slower (9.0 against 8.1 ms at two
    // clusters): the clusters are not in lock step, and a cluster's sums must not wait for another cluster's inputs.
    auto src_of = [&](int it_, int q_) { return P.cl[q_].hin_ll + P.hin_off[l] + (size_t)(it_ & slot_mask) * SEG * H; };
    unsigned long long pre[WPT];
    {
        const unsigned long long* s0 = src_of(0, 0);
#pragma unroll
        for (int r = 0; r < WPT; ++r) {
            const int i = tid + r * SWN_NTHREADS;
            pre[r] = i < n_words ? ld_ll(s0 + i) : 0ull;
        }
    }
    for (int it = 0; it < P.n_it; ++it) {
        const uint32_t tag = (uint32_t)it + 1u;
        for (int q = 0; q < ncl; ++q) {
        const unsigned long long* src = src_of(it, q);
        float* hbw = hb_all + (size_t)q * SEG * HP;
        bool ok = true;
#pragma unroll
        for (int r = 0; r < WPT; ++r) {
            const int i = tid + r * SWN_NTHREADS;
            if (i < n_words) {
                uint32_t bits = (uint32_t)pre[r];
                if ((uint32_t)(pre[r] >> 32) != tag) ok = wait_gword(src + i, tag, bits, P, 300 + l) && ok;
                const int c = i / H, k = i - c * H;
                hbw[c * HP + k] = __uint_as_float(bits);
            }
        }
        if (__syncthreads_or(!ok)) return;
        SWN_TSTAMP(2 * q, tid == 0);                                // helper timeline: inputs of cluster q are in
        if (P.help_delay_ns > 0) __nanosleep((unsigned)P.help_delay_ns);   // test hook: a late helper
        {
            int qn = q + 1, itn = it;
            if (qn == ncl) { qn = 0; ++itn; }
            if (itn < P.n_it) {
                const unsigned long long* sn = src_of(itn, qn);
#pragma unroll
                for (int r = 0; r < WPT; ++r) {
                    const int i = tid + r * SWN_NTHREADS;
                    pre[r] = i < n_words ? ld_ll(sn + i) : 0ull;
                }
            }
        }
        unsigned long long* acc_ll = P.cl[q].hacc_ll + P.hacc_off[l];
        float* ring = ring_all + (size_t)q * RL * np * 2;
        const float* hb = hb_all + (size_t)q * SEG * HP;
        float2 h[SEG][NCH];
#pragma unroll
        for (int c = 0; c < SEG; ++c) load_frag<NCH>(h[c], hb + c * HP, lane);
        for (int pi = warp; pi < np; pi += SWN_NWARPS) {
            // the lane's partial dot products of every (tap j, column c), slot j*SEG + c, then ONE transposed reduction per
            // row kind: 31 shuffles for up to 32 slots instead of five per slot (with three core clusters to serve, the
            // shuffle unit was what bounded the helpers: 11.1 ms against 8.1 ms for one utterance)
#if SEG_HELP_ZC_JOINT
            // both row kinds at once: two independent reduction chains per level (a turn is latency-bound: 3-4 warps per scheduler)
            float vz[32], vc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { vz[i] = 0.f; vc[i] = 0.f; }
#pragma unroll
            for (int j = 0; j < CORE_KMAX; ++j) {
                if ((j + 1) * SEG <= 32 && j < nt) {
                    float2 wz[NCH], wc[NCH];
                    load_frag<NCH>(wz, sW + ((size_t)(j * np + pi) * 2) * HP, lane);
                    load_frag<NCH>(wc, sW + ((size_t)(j * np + pi) * 2 + 1) * HP, lane);
#pragma unroll
                    for (int c = 0; c < SEG; ++c) {
                        vz[(j * SEG + c) & 31] = dot_frag_packed<NCH>(wz, h[c]);
                        vc[(j * SEG + c) & 31] = dot_frag_packed<NCH>(wc, h[c]);
                    }
                }
            }
            float mzc[2];
            reduce_slots32x2(vz, vc, lane, mzc[0], mzc[1]);
#else
            float mzc[2];
#pragma unroll
            for (int zc = 0; zc < 2; ++zc) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
                for (int j = 0; j < CORE_KMAX; ++j) {
                    if ((j + 1) * SEG <= 32 && j < nt) {
                        float2 w[NCH];
                        load_frag<NCH>(w, sW + ((size_t)(j * np + pi) * 2 + zc) * HP, lane);
#pragma unroll
                        for (int c = 0; c < SEG; ++c) v[(j * SEG + c) & 31] = dot_frag_packed<NCH>(w, h[c]);
                    }
                }
                mzc[zc] = reduce_slots32(v, lane);
            }
#endif
            const float mz = mzc[0], mc = mzc[1];
            // ring update, one tap per round (within a round the SEG columns hit SEG different times); the smallest
            // delay comes last and finishes its target times
            // (a dilation of at least SEG keeps all (tap, column) targets of a step distinct: one round)
            const bool one_round = D.dil >= SEG;
            for (int j = one_round ? 0 : nt - 1; j >= 0; --j) {
                if (my_live && (one_round || my_j == j)) {
                    const int T = it * SEG + my_c + my_delay;
                    float* r0 = ring + ((T % RL) * np + pi) * 2;
                    const float a = r0[0] + mz, b = r0[1] + mc;
                    if (my_j == 0) {
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
        SWN_TSTAMP(2 * q + 1, warp == SWN_NWARPS - 1 && lane == 0);  // ... the last warp has handed its sums over
        }
        __syncthreads();                                        // hb is rewritten by the next step's loads
    }
}

// ------------------------------------------------------------------ core CTA
// Totals of a[slot][z|c] (16 slots) over the 32 lanes: afterwards the two lanes of slot lane >> 1 hold both totals of that
// slot.  Four halving levels (16, 8, 4, 2) and one butterfly: 32 shuffles for 16 (pair, column) slots.
__device__ __forceinline__ void reduce_slots16(const float (&a)[16][2], int lane, float& qz, float& qc) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float u[8][2], v[4][2], w2[2][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int zc = 0; zc < 2; ++zc) {
            const float send = b4 ? a[i][zc] : a[i + 8][zc], keep = b4 ? a[i + 8][zc] : a[i][zc];
            u[i][zc] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int zc = 0; zc < 2; ++zc) {
            const float send = b3 ? u[i][zc] : u[i + 4][zc], keep = b3 ? u[i + 4][zc] : u[i][zc];
            v[i][zc] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int zc = 0; zc < 2; ++zc) {
            const float send = b2 ? v[i][zc] : v[i + 2][zc], keep = b2 ? v[i + 2][zc] : v[i][zc];
            w2[i][zc] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    float z, c;
    {
        const float sz = b1 ? w2[0][0] : w2[1][0], kz = b1 ? w2[1][0] : w2[0][0];
        const float sc = b1 ? w2[0][1] : w2[1][1], kc = b1 ? w2[1][1] : w2[0][1];
        z = kz + __shfl_xor_sync(0xffffffffu, sz, 2);
        c = kc + __shfl_xor_sync(0xffffffffu, sc, 2);
    }
    qz = z + __shfl_xor_sync(0xffffffffu, z, 1);
    qc = c + __shfl_xor_sync(0xffffffffu, c, 1);
}

template <int NCH, int SEG>
__device__ void core_cta_seg(const CoreParams& P, const CoreCta& D, float* smem, const int clid) {
    const CoreUtt& U = P.cl[clid];                              // this cluster's utterance
    static_assert(NCH <= 3 && SEG <= 8, "chain warps 0 .. 4*NCH-1, service warps 12-15; 8 column slots in the reduction");
    constexpr int HP = NCH * 64, NPW = NCH, FPT = 2 * NCH, ROWS = 2 * NPW, TCOLS = ROWS * FPT;
    constexpr int WT = 2 * FPT;                                 // tensor-memory columns of one warp per tap: 2 rows x FPT weights
    constexpr int HIST = 32;
    constexpr int NCHAIN = 4 * NCH;                             // chain warps
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
    float4* gop = reinterpret_cast<float4*>(smem + D.sm_pre);   // [2][L][16][SEG] gate operands {pz, bz*pz, pc, bc*pc}, by step parity
    float* misc = smem + D.sm_misc;
    float* hist = misc;                                    // [HIST] fed-back samples
    float* lpb = hist + HIST;                              // [2][8] LPC history, by head-step parity
    float* o2 = lpb + 16;                                  // [16] out_2 rows
    float* epsb = o2 + 16;                                 // [8] sign(e) log1p(-2|e|) of the step's noise
    float4* lc = reinterpret_cast<float4*>(epsb + 8);      // [L][16] per (layer, pair): conditioning const z/c, bias z/c
    float* sfir = reinterpret_cast<float*>(lc + L * 16);   // [8][2][16] FIR taps (cb, v) of this CTA's input rows
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + D.sm_tmem_slot);
    unsigned long long* mb = reinterpret_cast<unsigned long long*>(tmem_slot + 2);   // [L+1] layer inputs (+ last layer's output)
    unsigned long long* mb_x1 = mb + (L + 1), *mb_x2 = mb + (L + 2);
    unsigned long long* gb = mb + (L + 3);                 // [2][L] "gate operands of layer l are in", by step parity (2 arrivals: warps 14, 15)
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
    for (int i = tid; i < 2 * L * 16 * SEG; i += SWN_NTHREADS) gop[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < HIST + 16 + 16 + 8; i += SWN_NTHREADS) misc[i] = 0.f;
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
        for (int i = 0; i < 2 * L; ++i) mbar_init(gb + i, 2);
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
    // (thread-private columns: tap c occupies columns [c*TCOLS, (c+1)*TCOLS), row rr = 2*pair + {z,c} at rr*FPT); in the
    // loop the chain warp of team t reads the 2*FPT columns of pair slot t
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
    if (U.seedw) {
        // a_full = pad(R) ++ seed(SEG) ++ generated: the seed occupies generated-indices -SEG..-1; lp_x_buffer = last
        // lpc samples of the padded seed (cswnv_shift1.py:317-319)
        if (tid < SEG) {
            const float v = ((const float*)U.seedw)[tid];
            hist[(tid - SEG + HIST) & (HIST - 1)] = v;
            if (8 - SEG + tid >= 0) lpb[8 - SEG + tid] = v;            // newest last (slot 7); slots beyond lpc have zero coefficients
        }
    }
    __syncthreads();
    cluster_sync_all();                                            // every CTA's buffers and barriers are ready

    const uint32_t peer = (uint32_t)(lane & (CORE_CL - 1));
    const uint32_t hx_peer = mapa(smem_u32(hx), peer), mb_peer = mapa(smem_u32(mb), peer);
    // ---- chain warps: warp (quarter q, team t) owns gate pair NPW*q + t of this CTA for all SEG columns
    const int plocal = NPW * quarter + team;
    const bool is_chain = warp < NCHAIN;
    const bool pvalid = is_chain && plocal < np;
    const int my_i = D.pair0 + (pvalid ? plocal : 0);              // row index in H (a valid row for idle warps too)
    const int cg = lane >> 2, sub4 = lane & 3;                     // column group of the lane after the reduction; its 4 lanes
    const int ccl = cg < SEG ? cg : SEG - 1;                       // publish to CTAs sub4, sub4+4, sub4+8, sub4+12
    const uint32_t hx_mine = smem_u32(hx) + (uint32_t)(my_i * 4), mb_mine = smem_u32(mb);
    // ---- service warps 12-15: four out_skip rows each (last column only), gate operands
    const int srow0 = 8 * (warp - 12);                              // skip warps 12, 13: eight out_skip rows each
    const bool has_b = warp < D.nb, has_c = warp < D.nc;
    // per-layer record of the chain stage in shared memory: indexed kernel-parameter loads (LDC c[0][R+imm]) between the
    // wake-up and the publish showed up as the largest single stall of the stage (profiles/r02u: short scoreboard on the
    // LDC.64 of P.hin_off[l+1])
    // fat: this layer has a single core tap and runs on team 0 alone (one warp per quarter, all its pairs).  nxa_*: the next
    // layer (wrapping into the next step) = what team 0 requests behind its publish; nxb_*: the next layer in the one-pair-per-
    // warp layout (-1: none) = what teams 1, 2 request behind theirs.
    struct LRec { int8_t nco, nco_next, nxb_nco, fat, nxa_fat, nxb_wrap; int16_t dil; int16_t tcol, tcol_next, nxb_tcol, r0; int32_t hmask, r1;
                  unsigned long long hin; };
    static_assert(sizeof(LRec) == 32, "per-layer record: two 16-byte loads");
    LRec* lrec = reinterpret_cast<LRec*>(reinterpret_cast<float4*>(sfir + 8 * 2 * 16) + 20);   // [L] behind wconst
    if (tid < L) {
        const int l = tid;
        LRec r;
        r.nco = (int8_t)P.ncore[l]; r.dil = (int16_t)(P.dil[l] < 32767 ? P.dil[l] : 32767); r.tcol = (int16_t)(P.tap_base[l] * TCOLS);
        const bool hn_ = l + 1 < L && P.has_help[l + 1];
        r.hmask = hn_ ? P.hslots[l + 1] - 1 : -1;                  // -1: no helper behind this layer's output
        r.hin = hn_ ? (unsigned long long)(U.hin_ll + P.hin_off[l + 1]) : 0ull;
        r.nco_next = (int8_t)P.ncore[(l + 1) % L]; r.tcol_next = (int16_t)(P.tap_base[(l + 1) % L] * TCOLS);     // (the last layer: layer 0 of the next step)
        r.fat = SEG_FAT && P.ncore[l] == 1 && NPW * SEG <= 16;
        r.nxa_fat = SEG_FAT && P.ncore[(l + 1) % L] == 1 && NPW * SEG <= 16;
        r.nxb_tcol = 0; r.nxb_nco = -1; r.nxb_wrap = 0; r.r0 = 0; r.r1 = 0;
        for (int k = 1; k <= L; ++k) {
            const int l2 = (l + k) % L;
            if (!(SEG_FAT && P.ncore[l2] == 1 && NPW * SEG <= 16)) { r.nxb_tcol = (int16_t)(P.tap_base[l2] * TCOLS); r.nxb_nco = (int8_t)P.ncore[l2]; r.nxb_wrap = l + k >= L; break; }
        }
        lrec[l] = r;
    }
    __syncthreads();
    // per-warp constants of the head (out_1 / out_2 bias, input bias, skip constants) live in shared memory and are read
    // right before the waits they follow: as loop-long registers they pushed the chain section into spills
    float4* wconst = reinterpret_cast<float4*>(sfir + 8 * 2 * 16);      // [16] {out_1 bias, out_2 bias, input bias, -}, [4] skip constants
    const int in_row = D.i0 + warp;
    const bool has_in = warp < D.ni;
    if (lane == 0) {
        wconst[warp] = make_float4(has_b ? __ldg(P.wbuf + D.b1_off + warp) : 0.f, has_c ? __ldg(P.wbuf + D.b2_off + warp) : 0.f,
                                   has_in ? __ldg(P.in_bias + in_row) : 0.f, 0.f);
        if (warp == 12 || warp == 13) {
            float k8[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) k8[r] = (srow0 + r < D.nskip) ? __ldg(P.wbuf + D.skipc_off + srow0 + r) : 0.f;
            wconst[16 + 2 * (warp - 12)] = make_float4(k8[0], k8[1], k8[2], k8[3]);
            wconst[17 + 2 * (warp - 12)] = make_float4(k8[4], k8[5], k8[6], k8[7]);
        }
    }
    __syncwarp();
    // the cluster-wide "previous columns are no longer read" barrier (below) is only needed where the head does not couple
    // every chain warp of every CTA to the next step: in the warm-up, and when some chain warp owns no out_1 row
    const bool cb_always = P.audio_in != 0;                        // (field reused as a flag by the host plan)
    bool ok = true;

    // ---- gate operands of (step s_, layer l), by the 64 lanes of the operand warps 14, 15: lane <-> (pair, column).  These
    // warps do nothing else in the layer section and run up to a step ahead (operands and their barriers are double-buffered by
    // step parity), so they are free to block on L2 and on a late helper.
    const int pidx = (warp - 14) * 32 + lane;
    auto prep = [&](int s_, int l) {
        for (int idx = pidx; idx < np * SEG; idx += 64) {
            const int p = idx / SEG, c = idx - p * SEG;
            const int row = D.pair0 + p;
            const int t = P.t_first + s_ * SEG + c;
            const float4 c4 = lc[l * 16 + p];
            float az = c4.x, ac = c4.y;
            int f = 0, j = 0;
            if (t >= P.R) { f = (t - P.R) / P.U; j = (t - P.R) - f * P.U; }
            const float* pl = U.P + (size_t)l * SEG * twoH + row;
            float hz = 0.f, hc = 0.f;
            const int T = s_ * SEG + c;
            const bool hh = P.has_help[l] && T >= P.hdmin[l];
            const unsigned long long* src = U.hacc_ll + P.hacc_off[l] + (size_t)(T & (P.hacc_nt[l] - 1)) * twoH + row;
            unsigned long long va = 0ull, vb = 0ull;
            if (hh) { va = ld_ll(src); vb = ld_ll(src + H); }
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
            if (hh) {
                const uint32_t tg = (uint32_t)T + 1u;
                uint32_t a = (uint32_t)va, b = (uint32_t)vb;
                if ((uint32_t)(va >> 32) != tg) ok = wait_gword(src, tg, a, P, 400 + l) && ok;       // the helper is late
                if ((uint32_t)(vb >> 32) != tg) ok = wait_gword(src + H, tg, b, P, 400 + l) && ok;
                hz = __uint_as_float(a); hc = __uint_as_float(b);
            }
            gop[(((s_ & 1) * L + l) * 16 + p) * SEG + c] = make_float4(az, (c4.z + hz) * az, ac, (c4.w + hc) * ac);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(gb + (s_ & 1) * L + l);
    };
    // a layer is "late" when its helper sums for step s only exist once step s-1 is (nearly) over: one step of helper slack and
    // not among the first two layers.  Late layers of step s are fetched at the top of step s, everything else a step ahead.
    auto late = [&](int l) { return l >= 2 && P.has_help[l] && P.hdmin[l] < 2 * SEG; };

    // ---- input stage: the SEG columns of block itn (cswnv_shift1.py:352), rows of this CTA, published into every
    // peer's hx[parity of itn][0].  Gather form over the sample history (tap validity = the conv's zero padding).
    float cbsum = 0.f;                                             // sum of the taps' constant terms (all taps valid)
    for (int j = 0; j < K; ++j) cbsum += sfir[(j * 2) * 16 + (warp < 16 ? warp : 0)];
    auto input_stage = [&](int itn) {
        if (!has_in) return;
        const int c = (lane & 15) < SEG ? (lane & 15) : SEG - 1;          // lanes 16.. mirror lanes 0..: they hand the columns to the helper
        const int t = P.t_first + itn * SEG + c;
        const int t0b = P.t_first + itn * SEG;                     // first column of the block
        float v = 0.f;
        if (t0b - (K - 1) >= 0 && t0b - (K - 1) - P.R - 1 >= -SEG) {
            // steady state (warp-uniform): every tap of every column is valid and inside the sample history
            float acc = wconst[warp].z + cbsum;
            const int n0 = t - (K - 1) - P.R - 1 + HIST;
#pragma unroll
            for (int j = 0; j < CORE_KMAX; ++j)
                if (j < K) acc = fmaf(sfir[(j * 2 + 1) * 16 + warp], hist[(n0 + j) & (HIST - 1)], acc);
            v = acc * rcp_approx(1.f + fabsf(acc));
        } else if (t >= 0) {
            float acc = wconst[warp].z;
#pragma unroll
            for (int j = 0; j < CORE_KMAX; ++j) {
                const int tau = t - (K - 1) + j;                             // hidden time of the tap
                if (j < K && tau >= 0) {                                     // (tau < 0: conv zero padding)
                    const int n = tau - P.R - 1;                             // newest sample seen by hidden time tau
                    const float a = (n >= -SEG) ? hist[(n + HIST) & (HIST - 1)] : 0.f;
                    acc += sfir[(j * 2) * 16 + warp] + sfir[(j * 2 + 1) * 16 + warp] * a;
                }
            }
            v = acc * rcp_approx(1.f + fabsf(acc));
        }
        const uint32_t base = mapa(smem_u32(hx), (uint32_t)(lane & (CORE_CL - 1))) + (uint32_t)((((itn & 1) * (L + 1)) * SEG * HP + in_row) * 4);
        const uint32_t mbp = mapa(smem_u32(mb), (uint32_t)(lane & (CORE_CL - 1)));
        const uint32_t lo16 = lane < CORE_CL ? 1u : 0u;
#pragma unroll
        for (int cc = 0; cc < SEG; ++cc) {
            const float vc = __shfl_sync(0xffffffffu, v, cc);
            st_async_f32_if(lo16, base + (uint32_t)(cc * HP * 4), vc, mbp);
        }
        if (lane >= 16 && lane < 16 + SEG && P.has_help[0])
            st_ll(U.hin_ll + P.hin_off[0] + ((size_t)(itn & (P.hslots[0] - 1)) * SEG + (lane - 16)) * H + in_row, v, (uint32_t)itn + 1u);
    };

    // ---- prologue: operands of the first two layers of step 0, inputs of block 0 (padding / seed)
    if (warp >= 14)
        for (int l = 0; l < L; ++l) prep(0, l);                    // every layer of step 0 (nothing to wait for yet)
    input_stage(0);

    // The tap registers of a chain warp: [SEG taps][WT] of its pair for a one-pair-per-warp layer, or (team 0, `fat` layer)
    // the single tap of all NPW pairs of the quarter [TCOLS] - the same registers: a warp only ever holds its next own stage.
    constexpr int WREGS = SEG * WT > TCOLS ? SEG * WT : TCOLS;
    float wr[WREGS];
#define W_(e, q) wr[(e) * WT + (q)]
    auto request_w = [&](int tcol, int nco_) {
        const uint32_t ta = taddr0 + tcol + team * WT;
        tmem_ld_at<WT, 0>(ta, wr);
        if (nco_ > 1) {
            if (1 < SEG && 1 < nco_) tmem_ld_at<WT, (1 < SEG ? 1 : 0) * WT>(ta + 1 * TCOLS, wr);
            if (2 < SEG && 2 < nco_) tmem_ld_at<WT, (2 < SEG ? 2 : 0) * WT>(ta + 2 * TCOLS, wr);
            if (3 < SEG && 3 < nco_) tmem_ld_at<WT, (3 < SEG ? 3 : 0) * WT>(ta + 3 * TCOLS, wr);
            if (4 < SEG && 4 < nco_) tmem_ld_at<WT, (4 < SEG ? 4 : 0) * WT>(ta + 4 * TCOLS, wr);
            static_assert(SEG <= 5, "tap requests are written out for up to five taps");
        }
    };
    auto request_fat = [&](int tcol) { tmem_ld_at<TCOLS, 0>(taddr0 + tcol, wr); };
    auto request_first = [&]() {                                   // the taps of this warp's first own stage of a step
        const LRec l0 = lrec[0];
        if (l0.fat) {
            if (team == 0) request_fat(l0.tcol);
            else if (l0.nxb_nco > 0) request_w(l0.nxb_tcol, l0.nxb_nco);
        } else request_w(l0.tcol, l0.nco);
    };
    if (is_chain) request_first();
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
        if (is_chain) {
            // ======================= chain warps: one gate pair, all columns: core taps (gather form), gate, publish =======================
            const bool t_neg = it == 0 && P.t_first + ccl < 0;               // alignment columns before time 0
            for (int l = 0; l < L; ++l) {
                const LRec lr = lrec[l];
                const int nco = lr.nco, dl = lr.dil;
                const bool lastl = l + 1 == L;
                if (lr.fat) {
                    // ---------- single core tap: team 0 alone, one warp per quarter with all NPW pairs: one 16-slot reduction,
                    // one gate, one publish per quarter instead of NPW of each on the same scheduler at the same time
                    if (team != 0) continue;
                    tmem_wait_ld();
                    ok = mbar_wait2(gb + par * L + l, (uint32_t)(it >> 1) & 1u, P, 480 + l) && ok;
                    ok = mbar_wait2(mb + l, par, P, 500 + l) && ok;
#if SWN_STAMPS
                    if (P.prof && tid == 0 && l < 6 && it >= P.prof_start && it < P.prof_start + P.prof_count)
                        P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + 1 + l] = gtime();
#endif
                    const int slot = lane >> 1, sj = slot / SEG, sc = slot - sj * SEG;       // (pair of the quarter, column) of this lane
                    const int prow = NPW * quarter + sj;
                    const bool svalid = slot < NPW * SEG && prow < np;
                    const float4 g4 = gop[((par * L + l) * 16 + (svalid ? prow : 0)) * SEG + (svalid ? sc : 0)];
                    const float carry = hcur[(l * SEG + (svalid ? sc : 0)) * HP + D.pair0 + (svalid ? prow : 0)];
                    float a16[16][2];
#pragma unroll
                    for (int i = 0; i < 16; ++i) { a16[i][0] = 0.f; a16[i][1] = 0.f; }
#pragma unroll
                    for (int c = 0; c < SEG; ++c) {
                        if (!lastl || c == SEG - 1) {
                            float2 h[NCH];
                            load_frag<NCH>(h, hcur + (l * SEG + c) * HP, lane);
#pragma unroll
                            for (int j = 0; j < NPW; ++j)
#pragma unroll
                                for (int zc = 0; zc < 2; ++zc)
#pragma unroll
                                    for (int cc = 0; cc < NCH; ++cc) {
                                        a16[j * SEG + c][zc] = fmaf(wr[(2 * j + zc) * FPT + 2 * cc], h[cc].x, a16[j * SEG + c][zc]);
                                        a16[j * SEG + c][zc] = fmaf(wr[(2 * j + zc) * FPT + 2 * cc + 1], h[cc].y, a16[j * SEG + c][zc]);
                                    }
                        }
                    }
                    float qz, qc;
                    reduce_slots16(a16, lane, qz, qc);
                    float hn = gate_out(fmaf(qz, g4.x, g4.y), fmaf(qc, g4.z, g4.w), carry);
                    if (it == 0 && P.t_first + sc < 0) hn = 0.f;
                    if (svalid && (!lastl || sc == SEG - 1)) {
                        const uint32_t off = (uint32_t)(((par * (L + 1) + (l + 1)) * SEG + (lastl ? 0 : sc)) * HP * 4 + (D.pair0 + prow) * 4);
                        const uint32_t mo = (uint32_t)((l + 1) * 8);
                        const uint32_t hx0 = smem_u32(hx), s1 = (uint32_t)(lane & 1);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            st_async_f32(mapa(hx0 + off, s1 + 2 * k), hn, mapa(mb_mine + mo, s1 + 2 * k));
                        if (lr.hmask >= 0 && s1 == 0)
                            st_ll(reinterpret_cast<unsigned long long*>(lr.hin) + ((size_t)(it & lr.hmask) * SEG + sc) * H + D.pair0 + prow, hn,
                                  (uint32_t)it + 1u);
                    }
                    if (!lastl || it + 1 < P.n_it) {        // the next layer's taps (team 0 takes part in every layer)
                        if (lr.nxa_fat) request_fat(lr.tcol_next);
                        else request_w(lr.tcol_next, lr.nco_next);
                    }
                    continue;
                }
                float acc[SEG][2];
#pragma unroll
                for (int c = 0; c < SEG; ++c) { acc[c][0] = 0.f; acc[c][1] = 0.f; }
                tmem_wait_ld();
                const bool more = !lastl || it + 1 < P.n_it;                   // a next layer exists (the last layer wraps to the next step)
                if (dl == 1) {
                    // dilation 1: source columns -(SEG-1) .. SEG-1 (negative = the previous step's), each loaded once; tap e
                    // of source s feeds target column s + e (taps beyond ncore are zero-filled).  The previous step's
                    // sources do not depend on this step's input: they are done BEFORE the wait, in the shadow of the hop.
#pragma unroll
                    for (int e = 1; e < SEG; ++e)
                        if (e >= nco)
#pragma unroll
                            for (int q = 0; q < WT; ++q) W_(e, q) = 0.f;
#pragma unroll
                    for (int sidx = 0; sidx < SEG - 1; ++sidx) {
                        const int s = sidx - (SEG - 1);
                        float2 h[NCH];
                        load_frag<NCH>(h, hprev + (SEG + s) * HP + l * SEG * HP, lane);
#pragma unroll
                        for (int e = 1; e < SEG; ++e) {
                            const int c = s + e;
                            if (c >= 0 && c < SEG) {
#pragma unroll
                                for (int cc = 0; cc < NCH; ++cc) {
                                    acc[c][0] = fmaf(W_(e, 2 * cc), h[cc].x, acc[c][0]);
                                    acc[c][0] = fmaf(W_(e, 2 * cc + 1), h[cc].y, acc[c][0]);
                                    acc[c][1] = fmaf(W_(e, FPT + 2 * cc), h[cc].x, acc[c][1]);
                                    acc[c][1] = fmaf(W_(e, FPT + 2 * cc + 1), h[cc].y, acc[c][1]);
                                }
                            }
                        }
                    }
                }
                SWN_TSTAMP(17, tid == 0 && l == 1);
                ok = mbar_wait2(gb + par * L + l, (uint32_t)(it >> 1) & 1u, P, 480 + l) && ok;
                SWN_TSTAMP(18, tid == 0 && l == 1);
                ok = mbar_wait2(mb + l, par, P, 500 + l) && ok;
#if SWN_STAMPS
                if (P.prof && tid == 0 && l < 6 && it >= P.prof_start && it < P.prof_start + P.prof_count)
                    P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + 1 + l] = gtime();
#endif
                const float4 g4 = gop[((par * L + l) * 16 + (pvalid ? plocal : 0)) * SEG + ccl];
                const float carry = hcur[(l * SEG + ccl) * HP + my_i];
                if (dl == 1) {
#pragma unroll
                    for (int s = 0; s < SEG; ++s) {
                        float2 h[NCH];
                        load_frag<NCH>(h, hcur + s * HP + l * SEG * HP, lane);
#pragma unroll
                        for (int e = 0; e < SEG; ++e) {
                            const int c = s + e;
                            if (c < SEG) {
#pragma unroll
                                for (int cc = 0; cc < NCH; ++cc) {
                                    acc[c][0] = fmaf(W_(e, 2 * cc), h[cc].x, acc[c][0]);
                                    acc[c][0] = fmaf(W_(e, 2 * cc + 1), h[cc].y, acc[c][0]);
                                    acc[c][1] = fmaf(W_(e, FPT + 2 * cc), h[cc].x, acc[c][1]);
                                    acc[c][1] = fmaf(W_(e, FPT + 2 * cc + 1), h[cc].y, acc[c][1]);
                                }
                            }
                        }
                    }
                } else if (nco == 1 && !lastl) {
                    // dilation >= SEG (the usual case): delay 0 only, column c from column c - fixed addresses, all loads first
                    float2 h[SEG][NCH];
#pragma unroll
                    for (int c = 0; c < SEG; ++c) load_frag<NCH>(h[c], hcur + c * HP + l * SEG * HP, lane);
#pragma unroll
                    for (int c = 0; c < SEG; ++c)
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            acc[c][0] = fmaf(W_(0, 2 * cc), h[c][cc].x, acc[c][0]);
                            acc[c][0] = fmaf(W_(0, 2 * cc + 1), h[c][cc].y, acc[c][0]);
                            acc[c][1] = fmaf(W_(0, FPT + 2 * cc), h[c][cc].x, acc[c][1]);
                            acc[c][1] = fmaf(W_(0, FPT + 2 * cc + 1), h[c][cc].y, acc[c][1]);
                        }
                } else {
                    // 2 <= dilation < SEG, or the last layer: at most (SEG+1)/2 core taps; one fragment load per (tap, target column)
#pragma unroll
                    for (int e = 0; e < (SEG + 1) / 2; ++e) {
                        if (e < nco) {
                            const int dly = e * dl;
#pragma unroll
                            for (int c = 0; c < SEG; ++c) {
                                if (!lastl || c == SEG - 1) {
                                    const int src = c - dly;
                                    const float* hp = (src >= 0 ? hcur + src * HP : hprev + (SEG + src) * HP) + l * SEG * HP;
                                    float2 h[NCH];
                                    load_frag<NCH>(h, hp, lane);
#pragma unroll
                                    for (int cc = 0; cc < NCH; ++cc) {
                                        acc[c][0] = fmaf(W_(e, 2 * cc), h[cc].x, acc[c][0]);
                                        acc[c][0] = fmaf(W_(e, 2 * cc + 1), h[cc].y, acc[c][0]);
                                        acc[c][1] = fmaf(W_(e, FPT + 2 * cc), h[cc].x, acc[c][1]);
                                        acc[c][1] = fmaf(W_(e, FPT + 2 * cc + 1), h[cc].y, acc[c][1]);
                                    }
                                }
                            }
                        }
                    }
                }
                SWN_TSTAMP(19, tid == 0 && l == 1);
                // next layer's taps: requested as soon as the registers are free (tensor-memory reads are 64 B/clk per SM: the
                // five taps of a dilation-1 layer take 0.7 us per CTA, more than the hop alone would hide)
                float qz, qc;
                reduce_columns<SEG>(acc, lane, qz, qc);
                SWN_TSTAMP(20, tid == 0 && l == 1);
                float hn = gate_out(fmaf(qz, g4.x, g4.y), fmaf(qc, g4.z, g4.w), carry);
                if (t_neg) hn = 0.f;
                if (pvalid && cg < SEG && (!lastl || cg == SEG - 1)) {
                    // the last layer's output is a single column (slot 0 of hx[.][L])
                    const uint32_t off = (uint32_t)(((par * (L + 1) + (l + 1)) * SEG + (lastl ? 0 : cg)) * HP * 4);
                    const uint32_t mo = (uint32_t)((l + 1) * 8);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        st_async_f32(mapa(hx_mine + off, (uint32_t)(sub4 + 4 * k)), hn, mapa(mb_mine + mo, (uint32_t)(sub4 + 4 * k)));
                    if (lr.hmask >= 0 && sub4 == 0)
                        st_ll(reinterpret_cast<unsigned long long*>(lr.hin) + ((size_t)(it & lr.hmask) * SEG + cg) * H + my_i, hn,
                              (uint32_t)it + 1u);
                }
                SWN_TSTAMP(21, tid == 0 && l == 1);
                {
                    // behind the publish: the taps of this warp's next own stage
                    if (team == 0) {
                        if (more) {
                            if (lr.nxa_fat) request_fat(lr.tcol_next);
                            else request_w(lr.tcol_next, lr.nco_next);
                        }
                    } else if (lr.nxb_nco > 0 && (!lr.nxb_wrap || it + 1 < P.n_it)) request_w(lr.nxb_tcol, lr.nxb_nco);
                }
            }
        } else if (warp >= 14) {
            // ======================= operand warps: noise, then gate operands - the late layers of this step, everything else of the next =======================
            if (active && warp == 15 && lane < SEG) {
                const long long n = (long long)istep * SEG + lane;
                const float e = U.noise ? __ldg(U.noise + n) : laplace_eps(P.seed, U.b0, n);
                const float sgn = (e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : 0.f);
                epsb[lane] = sgn * log1pf(-2.f * fabsf(e));
            }
            if (it > 0)
                for (int l = 0; l < L; ++l) if (late(l)) prep(it, l);
            if (it + 1 < P.n_it) {
                for (int l = 2; l < L; ++l) if (!late(l)) prep(it + 1, l);
                for (int l = 0; l < 2 && l < L; ++l) prep(it + 1, l);          // (their helper inputs are the youngest: last)
            }
        } else if (warp >= 12) {
            // ======================= skip warps: eight out_skip rows each, last column only; x1 =======================
            float sk[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) sk[r] = 0.f;
            const int nsk = D.nskip * HP;
            for (int l = 0; l < L; ++l) {
                ok = mbar_wait2(mb + l, par, P, 700 + l) && ok;
                if (active && l >= 1) {                                          // out_skip[l-1] on the last column of layer l's input
                    float2 h[NCH];
                    load_frag<NCH>(h, hcur + (l * SEG + SEG - 1) * HP, lane);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        float2 a[NCH];
                        load_frag<NCH>(a, sSkip + (size_t)(l - 1) * nsk + (size_t)(srow0 + r < D.nskip ? srow0 + r : 0) * HP, lane);
                        sk[r] += dot_frag<NCH>(a, h);
                    }
                }
            }
            {
                const float4 ka = wconst[16 + 2 * (warp - 12)], kb = wconst[17 + 2 * (warp - 12)];
                const float skc[8] = {ka.x, ka.y, ka.z, ka.w, kb.x, kb.y, kb.z, kb.w};
                ok = mbar_wait2(mb + L, par, P, 510) && ok;
                if (active) {
                    float2 hl[NCH];
                    load_frag<NCH>(hl, hcur + (size_t)L * SEG * HP, lane);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        float2 a[NCH];
                        load_frag<NCH>(a, sSkip + (size_t)(L - 1) * nsk + (size_t)(srow0 + r < D.nskip ? srow0 + r : 0) * HP, lane);
                        sk[r] += dot_frag<NCH>(a, hl);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                        for (int r = 0; r < 8; ++r) sk[r] += __shfl_xor_sync(0xffffffffu, sk[r], o);
                    if (lane < CORE_CL) {
                        const uint32_t dst = hx_peer + (uint32_t)((2 * LS + D.skip0 + srow0) * 4);
#pragma unroll
                        for (int r = 0; r < 8; ++r)
                            if (srow0 + r < D.nskip)
                                st_async_f32(dst + (uint32_t)(r * 4), fmaxf(sk[r] + skc[r], 0.f), mb_peer + (uint32_t)((L + 1) * 8));
                    }
                }
                // every layer barrier of this step has completed (this thread waited for each): arm the next phase
                if (warp == 12 && lane == 0) {
                    for (int i = 0; i < L; ++i) mbar_arm(mb + i, bytes_h);
                    mbar_arm(mb + L, bytes_last);
                }
            }
        }
        // The previous step's columns have been read for the last time.  Where the head does not couple every chain warp
        // of every CTA to the next block's inputs (warm-up; models with fewer out_1 rows than chain warps) a fast CTA could
        // overwrite them a step early: cluster barrier, arrive here, wait right before the next inputs are published.
        // (Its acquire side flushes L1, so it is not used where the head already gives the ordering.)
        const bool cb = !active || cb_always;
        if (cb) cluster_arrive();
        SWN_STAMP(7);
        float fed[SEG];
#pragma unroll
        for (int j = 0; j < SEG; ++j) fed[j] = 0.f;
        if (active) {
            // ---- out_1 row of this warp (relu) -> x2 on every CTA
            if (has_b) {
                const float* wrow = sW1 + (size_t)warp * Sp;
                float2 wpre[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    wpre[cc] = (cc * 64 < Sp) ? reinterpret_cast<const float2*>(wrow)[cc * 32 + lane] : make_float2(0.f, 0.f);
                const float bias1 = wconst[warp].x;
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
                const float v = warp_sum(a + b) + bias1;
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
                const float bias2 = wconst[warp].y;
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
                const float o = warp_sum(a + b) + bias2;
                if (lane == 0) o2[warp] = o;
            } else if (tid == 0) {
                ok = mbar_wait2(mb_x2, hpar, P, 531) && ok;
                mbar_arm(mb_x2, bytes_x2);
            }
            SWN_STAMP(16);
            __syncthreads();                                                     // o2[] is complete
            // ---- Laplace / LPC sampler (cswnv_shift1.py:368-391), by EVERY warp for itself (bit-identical): lane j evaluates
            // the scale and the noise term of sample j, then all lanes run the LPC recursion on the broadcast values; no
            // second CTA barrier and no shared-memory round trip between "sampled" and the next block's input columns
            {
                const long long nbase = (long long)istep * SEG;
                const int jj = lane < SEG ? lane : 0;
                const float bj = sigmoidf_(o2[SEG + jj]);                         // exp(logsigmoid(.))
                const float bt = bj * epsb[jj];
                const float mu = o2[jj];
                // LPC history in 8 slots, newest LAST (slot 7), coefficients aligned to it and zero beyond lpc: the recursion
                // does not depend on lpc (as code it compiled to a jump table over local memory)
                const float* lpr = lpb + hpar * 8;
                float lh[8], aa[8];
#pragma unroll
                for (int mm = 0; mm < 8; ++mm) {
                    lh[mm] = lpr[mm];
                    // ass = a.flip(-1) times the buffer ordered oldest -> newest: a[0] multiplies the newest
                    aa[mm] = (7 - mm) < P.lpc ? o2[2 * SEG + (7 - mm)] : 0.f;
                }
                const float* forced = U.forced ? (const float*)U.forced + nbase : nullptr;
                float smp[SEG];
#pragma unroll
                for (int j = 0; j < SEG; ++j) {
                    float lpv = 0.f;
                    // (summation order of the layer-group kernel: newest-last buffer, oldest tap first)
#pragma unroll
                    for (int mm = 0; mm < 8; ++mm) lpv += aa[mm] * lh[mm];
                    float v = lpv + __shfl_sync(0xffffffffu, mu, j) - __shfl_sync(0xffffffffu, bt, j);
                    v = fminf(fmaxf(v, -1.f), 1.f);
                    smp[j] = v;
                    fed[j] = forced ? __ldg(forced + j) : v;
#pragma unroll
                    for (int mm = 0; mm < 7; ++mm) lh[mm] = lh[mm + 1];
                    lh[7] = v;
                }
                if (forced) {
                    // teacher forcing: inside the segment the recursion ran on its own samples (as the reference's sampler
                    // would); the history handed to the next step is the forced waveform
#pragma unroll
                    for (int mm = 0; mm < 8; ++mm) lh[mm] = lpr[mm];
#pragma unroll
                    for (int j = 0; j < SEG; ++j) {
#pragma unroll
                        for (int mm = 0; mm < 7; ++mm) lh[mm] = lh[mm + 1];
                        lh[7] = fed[j];
                    }
                }
                // every warp files the fed-back samples itself (identical values): its own input rows read them right away
                if (lane < SEG) {
                    float f = fed[0];
#pragma unroll
                    for (int j = 1; j < SEG; ++j) f = lane == j ? fed[j] : f;
                    hist[(int)((nbase + lane) & (HIST - 1))] = f;
                }
                if (warp == 0) {
                    if (lane < 8) {
                        float f = lh[0];
#pragma unroll
                        for (int mm = 1; mm < 8; ++mm) f = lane == mm ? lh[mm] : f;
                        lpb[(hpar ^ 1u) * 8 + lane] = f;
                    }
                    if (rank == 0) {
                        if (lane < SEG) {
                            float f = smp[0];
#pragma unroll
                            for (int j = 1; j < SEG; ++j) f = lane == j ? smp[j] : f;
                            ((float*)U.out)[nbase + lane] = f;
                        }
                        if (U.head_out && lane < n2) U.head_out[(size_t)istep * n2 + lane] = o2[lane];
                    }
                }
                __syncwarp();
            }
            SWN_STAMP(10);
        }
        if (cb) cluster_wait();
        if (it + 1 < P.n_it) {
            input_stage(it + 1);
            SWN_STAMP(11);
        }
        SWN_STAMP(12);
    }
    // ---- teardown
    tmem_wait_ld();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) tmem_free_all(tbase);
}

template <int NCH, int SEG>
__global__ void __cluster_dims__(CORE_CL, 1, 1) __launch_bounds__(SWN_NTHREADS, 1)
swn_coreseg_kernel(const __grid_constant__ CoreParams P) {
    extern __shared__ __align__(16) float smem[];
    const int clid = (int)cluster_id_x();
    if (clid < P.n_cl) {
        core_cta_seg<NCH, SEG>(P, P.cdesc[cluster_ctarank()], smem, clid);
    } else {
        const int hidx = (int)blockIdx.x - CORE_CL * P.n_cl;
        if (hidx < P.n_help && P.hdesc[hidx].npairs > 0) helper_cta_seg<NCH, SEG>(P, P.hdesc[hidx], smem);
    }
}

}  // namespace swn
