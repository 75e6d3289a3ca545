// Persistent generation kernel of the B200-native shallow-wavenet path.
//
// Replaces the per-sample Python loop of CSWNV.batch_fast_generate
// (reference src/nets/cswnv_shift1.py:314-402) and DSWNV.batch_fast_generate
// (src/nets/dswnv.py:312-374): warm-up over the padded seed, the six gated cells, the
// output head and the sampler run as ONE cooperative launch, one CTA per SM.
//
// Work split:
//   * layer l is owned by a group of G CTAs; each CTA keeps its rows of dil_h[l]
//     (all K taps) and of out_skip[l-1] resident in shared memory for the whole utterance;
//   * the dilated-conv history is kept in "scatter" form: when the input vector of time t
//     arrives, its contribution to every future output t+j*d is accumulated into a ring of
//     partial sums (SMEM, or L2-backed scratch when the batch is large); only the tap that
//     feeds the current output is on the critical path;
//   * the head group (NH CTAs) owns out_skip[L-1], out_1, out_2, the sampler and the
//     causal input stage of the next step;
//   * vectors travel between groups through L2 as 64-bit {fp32, tag} words written with one
//     relaxed store and polled with relaxed loads (no fences, no grid barrier).
//
// Template parameters: NCH = padded hidden size / 64 (register blocking of the input
// vector), SEG = time steps per loop iteration (the reference's `seg`), BS = utterances per
// pipeline slot.  NC = SEG*BS columns are processed together against one pass over the
// SMEM-resident weights.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include "swn_common.h"

namespace swn {

// ------------------------------------------------------------------ primitives
__device__ __forceinline__ unsigned long long ld_ll(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_ll(unsigned long long* p, float val, uint32_t tag) {
    unsigned long long v = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(val);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_ll_u32(unsigned long long* p, uint32_t val, uint32_t tag) {
    unsigned long long v = ((unsigned long long)tag << 32) | (unsigned long long)val;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Spin until the word carries `tag`; returns false when the watchdog fired or another CTA
// aborted (every caller funnels the result through __syncthreads_or and leaves).  The slow path is
// kept out of line: the persistent loops are latency-bound and must stay instruction-cache resident.
static __device__ __noinline__ bool ll_wait_slow(const unsigned long long* p, uint32_t tag, uint32_t& payload,
                                          const GenParams& P, int where) {
    long long t_start = clock64();
    uint32_t spins = 0;
    for (;;) {
        const unsigned long long v = ld_ll(p);
        if ((uint32_t)(v >> 32) == tag) { payload = (uint32_t)v; return true; }
        if (((++spins) & 255u) == 0) {
            volatile int32_t* st = P.status;
            if (st[0] != 0) return false;
            if ((unsigned long long)(clock64() - t_start) > P.timeout_cycles) {
                if (atomicCAS((int*)P.status, 0, 1) == 0) { st[1] = where; st[2] = (int)tag; st[3] = blockIdx.x; }
                return false;
            }
        }
    }
}
__device__ __forceinline__ bool ll_wait(const unsigned long long* p, uint32_t tag, uint32_t& payload,
                                        const GenParams& P, int where) {
    const unsigned long long v = ld_ll(p);
    if ((uint32_t)(v >> 32) == tag) { payload = (uint32_t)v; return true; }
    return ll_wait_slow(p, tag, payload, P, where);
}

// Cheap first phase of a wait: lane 0 of the warp watches ONE word of the vector with a sleep
// between polls, so that the ~140 CTAs that are idle at any time do not flood L2 with polls;
// the per-word LL polls start only when the vector is about to be complete.
__device__ __forceinline__ void sentinel_wait(const unsigned long long* p, uint32_t tag, const GenParams& P) {
    if (P.poll_mode == 0) return;
    if ((threadIdx.x & 31) == 0) {
        uint32_t spins = 0;
        while ((uint32_t)(ld_ll(p) >> 32) != tag) {
            __nanosleep(P.poll_sleep_ns);
            if (((++spins) & 1023u) == 0 && ((volatile int32_t*)P.status)[0] != 0) break;   // ll_wait reports the abort
        }
    }
    __syncwarp();
}

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define SWN_PROF(ev)                                                                              \
    do {                                                                                          \
        if (P.prof && threadIdx.x == 0 && it >= P.prof_start && it < P.prof_start + P.prof_count) \
            P.prof[((size_t)(it - P.prof_start) * gridDim.x + blockIdx.x) * 32 + (ev)] = gtime();  \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NCH>
__device__ __forceinline__ float dot_frag(const float2 (&w)[NCH], const float2 (&h)[NCH]) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) { a = fmaf(w[c].x, h[c].x, a); b = fmaf(w[c].y, h[c].y, b); }
    return a + b;
}

template <int NCH>
__device__ __forceinline__ void load_frag(float2 (&w)[NCH], const float* row, int lane) {
    const float2* r = reinterpret_cast<const float2*>(row);
#pragma unroll
    for (int c = 0; c < NCH; ++c) w[c] = r[c * 32 + lane];
}

// generic (runtime length) dot product of a weight row with a shared-memory vector, both
// padded to a multiple of 64 floats; lanes split k.
__device__ __forceinline__ float dot_row(const float* __restrict__ wrow, const float* __restrict__ x, int kp, int lane) {
    const float2* w2 = reinterpret_cast<const float2*>(wrow);
    const float2* x2 = reinterpret_cast<const float2*>(x);
    float a = 0.f, b = 0.f;
    for (int c = lane; c < (kp >> 1); c += 32) {
        float2 w = w2[c], v = x2[c];
        a = fmaf(w.x, v.x, a); b = fmaf(w.y, v.y, b);
    }
    return a + b;
}

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Gated-cell output (1 - z) tanh(gc) + z carry with z = sigmoid(gz) (cswnv_shift1.py:276-278), over one common
// denominator: [ez sgn(gc) (1 - ec) + carry (1 + ec)] / [(1 + ez)(1 + ec)], ez = exp(-gz) (capped: gz >= -30),
// ec = exp(-2|gc|).  Two ex2.approx and one rcp.approx on the dependent chain (abs. error ~1e-7).
__device__ __forceinline__ float gate_out(float gz, float gc, float carry) {
    const float ez = ex2_approx(fmaxf(gz, -30.f) * -1.4426950408889634f);
    const float ec = ex2_approx(fabsf(gc) * -2.8853900817779268f);
    const float den = (1.f + ez) * (1.f + ec);
    const float num = fmaf(ez, copysignf(1.f - ec, gc), carry * (1.f + ec));
    return num * rcp_approx(den);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float softsignf_(float x) { return x / (1.f + fabsf(x)); }

// ------------------------------------------------------------------ Philox4x32-10
__device__ __host__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// uniform in [0,1) with 24 bits
__device__ __host__ __forceinline__ float u01(uint32_t x) { return (float)(x >> 8) * (1.0f / 16777216.0f); }
// the uniform number the sampler draws for (utterance b, generated sample n, lane r)
static __device__ __host__ __noinline__ float philox_uniform(unsigned long long seed, long long b, long long n, int r) {
    uint32_t o[4];
    philox4x32_10((uint32_t)n, (uint32_t)((unsigned long long)n >> 32), (uint32_t)b, (uint32_t)(r >> 2),
                  (uint32_t)seed, (uint32_t)(seed >> 32), o);
    return u01(o[r & 3]);
}
// CSWNV: eps ~ U(-0.4999, 0.5)  (cswnv_shift1.py:373)
__device__ __host__ __forceinline__ float laplace_eps(unsigned long long seed, long long b, long long n) {
    return -0.4999f + 0.9999f * philox_uniform(seed, b, n, 0);
}
// DSWNV: q ~ Exp(1)  (the variate torch.multinomial divides by, dswnv.py:365)
__device__ __host__ __forceinline__ float softmax_q(unsigned long long seed, long long b, long long n, int r) {
    float u = philox_uniform(seed, b, n, r);
    // 1-u in (0,1]: u == 0 (probability 2^-24 per draw) would give q = 0 and a Gumbel score of +inf, i.e. one
    // uniformly random class every ~2^16 samples.  torch's exponential_ keeps its variate away from 0 the same way
    // (it returns eps/2 for the largest uniform); q >= 2^-24 caps the score at logit + 16.6.
    return fmaxf(-logf(1.0f - u), 5.9604645e-8f);
}
// non-negative remainder (the reference's OneHot does x % depth with Python semantics, dswnv.py:77)
__device__ __host__ __forceinline__ int pos_mod(long long x, int q) { const int r = (int)(x % q); return r < 0 ? r + q : r; }
// fed-back class index: the reference's OneHot would raise on an out-of-range index; the Python API validates
// `forced`, the kernels clamp so that a bad pointer can never index the gather tables out of bounds
__device__ __forceinline__ int clamp_idx(long long x, int q) { return x < 0 ? 0 : (x >= q ? q - 1 : (int)x); }

// ------------------------------------------------------------------ layer CTA
template <int NCH, int SEG, int BS>
__device__ void layer_cta(const GenParams& P, const CtaDesc& D, float* smem) {
    constexpr int NC = SEG * BS;
    constexpr int HP = NCH * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int l = D.layer, np = D.npairs, K = P.K, H = P.H, dil = D.dil, RL = D.ring_len;

    // ---- one-time: stage the resident weights
    float* sW = smem + D.sm_w;
    const float* gW = P.wbuf + D.w_off;
    const int jmin = K - D.smem_taps;                         // taps >= jmin live in SMEM
    {
        const float4* src = reinterpret_cast<const float4*>(gW + (size_t)jmin * np * 2 * HP);
        float4* dst = reinterpret_cast<float4*>(sW);
        const int n4 = D.smem_taps * np * 2 * HP / 4;
        for (int i = tid; i < n4; i += SWN_NTHREADS) dst[i] = src[i];
    }
    float* sWs = smem + D.sm_ws;
    const float* gWs = P.wbuf + D.ws_off;
    if (D.skip_in_smem) {
        const float4* src = reinterpret_cast<const float4*>(gWs);
        float4* dst = reinterpret_cast<float4*>(sWs);
        for (int i = tid; i < D.nskip * HP / 4; i += SWN_NTHREADS) dst[i] = src[i];
    }
    const float* Ws = D.skip_in_smem ? sWs : gWs;
    float* hb = smem + D.sm_h;                                // [2][NC][HP]
    for (int i = tid; i < 2 * NC * HP; i += SWN_NTHREADS) hb[i] = 0.f;
    float* ring;
    const size_t ring_slot = (size_t)RL * np * 2 * BS;        // floats per pipeline slot
    if (D.ring_in_smem) {
        ring = smem + D.sm_ring;
        for (size_t i = tid; i < ring_slot * P.M; i += SWN_NTHREADS) ring[i] = 0.f;
    } else {
        ring = P.ring_scratch + D.ring_off;                   // zeroed by the host
    }
    __syncthreads();

    const unsigned long long* in_ll = P.ring_ll + (size_t)l * P.M * NC * H;
    unsigned long long* out_ll = P.ring_ll + (size_t)(l + 1) * P.M * NC * H;
    const float* bias = P.wbuf + D.b_off;
    const int twoH = 2 * H;
    int hsel = 0;

    for (int it = 0; it < P.n_it; ++it) {
        const int t0 = P.t_first + it * SEG;
        const bool head_active = it >= P.it0;
        const uint32_t tag = (uint32_t)it + 1u;
        for (int m = 0; m < P.M; ++m) {
            float* hbuf = hb + hsel * NC * HP;
            hsel ^= 1;
            float* rg = ring + ring_slot * m;
            // ---- prefetch what the gate needs that does not depend on the input vector
            const int col = lane;                            // gate lane <-> column
            const int bl = col / SEG, s = col - bl * SEG;
            const int t = t0 + s;
            int bglob = m * BS + bl; if (bglob >= P.B) bglob = P.B - 1;
            float pz[2] = {0.f, 0.f}, pc[2] = {0.f, 0.f}, rz[2] = {0.f, 0.f}, rc[2] = {0.f, 0.f};
            bool ok = true;
            __syncwarp();
            // a warp may own up to two pairs (npairs <= 2*NWARPS)
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int pi = warp + pp * SWN_NWARPS;
                if (pi < np && col < NC && t >= 0) {
                    const int i = D.pair0 + pi;
                    float az = P.pconst[l * twoH + i], ac = P.pconst[l * twoH + H + i];
#pragma unroll
                    for (int sp = 0; sp < SEG; ++sp) {
                        const int tau = t + sp;
                        int f = 0, j = 0;
                        if (tau >= P.R) { f = (tau - P.R) / P.U; j = (tau - P.R) - f * P.U; }
                        const float g = P.gain[j];
                        const float* pr = P.P + ((((size_t)bglob * P.Tf + f) * P.L + l) * SEG + sp) * twoH;
                        az = fmaf(g, __ldg(pr + i), az);
                        ac = fmaf(g, __ldg(pr + H + i), ac);
                    }
                    if (P.audio_in) {
                        uint32_t idx;
                        ok = ll_wait(P.idx_ll + m * BS + bl, tag, idx, P, 10 + l) && ok;
                        if (ok) {
                            const float* xr = P.xa_tab + ((size_t)l * P.Q + idx) * twoH;
                            az += __ldg(xr + i); ac += __ldg(xr + H + i);
                        }
                    }
                    pz[pp] = az; pc[pp] = ac;
                    const int slot = (t - P.t_first) % RL;
                    float* r0 = rg + ((size_t)(slot * np + pi) * 2) * BS + bl;
                    // read-and-clear: the slot is reused RL steps later.  A ring that lives in global memory is
                    // updated with fire-and-forget reductions (below), so it is read at L2 (not through L1)
                    if (D.ring_in_smem) { rz[pp] = r0[0]; rc[pp] = r0[BS]; }
                    else { rz[pp] = __ldcg(r0); rc[pp] = __ldcg(r0 + BS); }
                    r0[0] = 0.f; r0[BS] = 0.f;
                }
            }
            // ---- wait for the input vectors of this block (H x NC words)
            {
                const unsigned long long* src = in_ll + (size_t)m * NC * H;
                sentinel_wait(src, tag, P);
                for (int i = tid; i < NC * H; i += SWN_NTHREADS) {
                    uint32_t bits;
                    ok = ll_wait(src + i, tag, bits, P, 100 + l) && ok;
                    const int c = i / H, k = i - c * H;
                    hbuf[c * HP + k] = __uint_as_float(bits);
                }
            }
            if (__syncthreads_or(!ok)) return;
            SWN_PROF(0);

            // ---- register-resident slice of the inputs: lane owns k = c*64 + 2*lane, +1
            float2 h[NC][NCH];
#pragma unroll
            for (int c = 0; c < NC; ++c) load_frag<NCH>(h[c], hbuf + c * HP, lane);

#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int pi = warp + pp * SWN_NWARPS;
                if (pi >= np) break;
                // ---- pass A: every (tap, source column) pair whose target lies inside this block
                float qz[NC], qc[NC];
#pragma unroll
                for (int c = 0; c < NC; ++c) { qz[c] = 0.f; qc[c] = 0.f; }
#pragma unroll
                for (int e = 0; e < SEG; ++e) {
                    if (e % dil == 0 && e / dil <= K - 1) {
                        const int j = K - 1 - e / dil;
                        const float* wr = (j >= jmin) ? sW + ((size_t)((j - jmin) * np + pi) * 2) * HP
                                                      : gW + ((size_t)(j * np + pi) * 2) * HP;
                        float2 wz[NCH], wc[NCH];
                        load_frag<NCH>(wz, wr, lane);
                        load_frag<NCH>(wc, wr + HP, lane);
#pragma unroll
                        for (int b = 0; b < BS; ++b)
#pragma unroll
                            for (int ss = 0; ss + e < SEG; ++ss) {
                                qz[b * SEG + ss + e] += dot_frag<NCH>(wz, h[b * SEG + ss]);
                                qc[b * SEG + ss + e] += dot_frag<NCH>(wc, h[b * SEG + ss]);
                            }
                    }
                }
                float myz = 0.f, myc = 0.f;
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float vz = warp_sum(qz[c]), vc = warp_sum(qc[c]);
                    if (lane == c) { myz = vz; myc = vc; }
                }
                // ---- gate (cswnv_shift1.py:281-285) and publish
                if (col < NC) {
                    const int i = D.pair0 + pi;
                    float hn = 0.f;
                    if (t >= 0) {
                        const float gz = (myz + bias[pi * 2] + rz[pp]) * pz[pp];
                        const float gc = (myc + bias[pi * 2 + 1] + rc[pp]) * pc[pp];
                        const float z = sigmoidf_(gz);
                        hn = (1.f - z) * tanhf(gc) + z * hbuf[col * HP + i];
                    }
                    st_ll(out_ll + ((size_t)m * NC + col) * H + i, hn, tag);
                }
            }

            SWN_PROF(1);
            // ---- pass B (off the critical path): contributions to future blocks
#pragma unroll
            for (int pp = 0; pp < 2; ++pp) {
                const int pi = warp + pp * SWN_NWARPS;
                if (pi >= np) break;
                for (int j = 0; j < K; ++j) {
                    const int delta = (K - 1 - j) * dil;
                    if (delta == 0) continue;
                    const float* wr = (j >= jmin) ? sW + ((size_t)((j - jmin) * np + pi) * 2) * HP
                                                  : gW + ((size_t)(j * np + pi) * 2) * HP;
                    float2 wz[NCH], wc[NCH];
                    load_frag<NCH>(wz, wr, lane);
                    load_frag<NCH>(wc, wr + HP, lane);
#pragma unroll
                    for (int b = 0; b < BS; ++b)
#pragma unroll
                        for (int ss = 0; ss < SEG; ++ss) {
                            if (ss + delta >= SEG) {           // warp-uniform
                                const float vz = warp_sum(dot_frag<NCH>(wz, h[b * SEG + ss]));
                                const float vc = warp_sum(dot_frag<NCH>(wc, h[b * SEG + ss]));
                                if (lane == b) {
                                    const int slot = (t0 + ss + delta - P.t_first) % RL;
                                    float* r0 = rg + ((size_t)(slot * np + pi) * 2) * BS + b;
                                    if (D.ring_in_smem) { r0[0] += vz; r0[BS] += vc; }
                                    else { atomicAdd(r0, vz); atomicAdd(r0 + BS, vc); }   // RED: no round trip to wait for
                                }
                            }
                        }
                }
            }
            // ---- rows of out_skip[l-1] applied to the last time column of every utterance
            if (l >= 1 && head_active) {
                unsigned long long* sk = P.skip_ll + ((size_t)(l - 1) * P.M + m) * BS * P.S;
                for (int r = warp; r < D.nskip; r += SWN_NWARPS) {
                    float2 w[NCH];
                    load_frag<NCH>(w, Ws + (size_t)r * HP, lane);
#pragma unroll
                    for (int b = 0; b < BS; ++b) {
                        const float v = warp_sum(dot_frag<NCH>(w, h[b * SEG + SEG - 1]));
                        if (lane == 0) st_ll(sk + (size_t)b * P.S + D.skip0 + r, v, tag);
                    }
                }
            }
            SWN_PROF(2);
        }
    }
}

// ------------------------------------------------------------------ head CTA
// Laplace / LPC sampler of one utterance for one loop step (cswnv_shift1.py:368-391).
template <int SEG>
static __device__ __noinline__ void laplace_step(const int lpc, const float* o, float* lp, const float* eps,
                                             const float* forced, float* smp) {
    float l[8];
#pragma unroll
    for (int mm = 0; mm < 8; ++mm) l[mm] = (mm < lpc) ? lp[mm] : 0.f;
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
        float lpv = 0.f;
        // ass = a.flip(-1) times the buffer ordered oldest -> newest: a[0] multiplies the newest
#pragma unroll
        for (int mm = 0; mm < 8; ++mm)
            if (mm < lpc) lpv += o[2 * SEG + lpc - 1 - mm] * l[mm];
        const float e = eps[j];
        const float sgn = (e > 0.f) ? 1.f : ((e < 0.f) ? -1.f : 0.f);
        const float b = sigmoidf_(o[SEG + j]);               // exp(logsigmoid(.))
        float v = lpv + o[j] - b * sgn * log1pf(-2.f * fabsf(e));
        v = fminf(fmaxf(v, -1.f), 1.f);
        smp[j] = v;
#pragma unroll
        for (int mm = 0; mm < 7; ++mm)
            if (mm + 1 < lpc) l[mm] = l[mm + 1];
#pragma unroll
        for (int mm = 0; mm < 8; ++mm)
            if (mm == lpc - 1) l[mm] = v;
    }
    if (forced) {
        // teacher forcing: inside the segment the recursion ran on its own samples (as the reference's
        // sampler would); the history handed to the next step is the forced waveform
        for (int j = 0; j < SEG; ++j) {
            for (int mm = 0; mm + 1 < lpc; ++mm) lp[mm] = lp[mm + 1];
            if (lpc > 0) lp[lpc - 1] = forced[j];
        }
    } else {
#pragma unroll
        for (int mm = 0; mm < 8; ++mm)
            if (mm < lpc) lp[mm] = l[mm];
    }
}

template <int NCH, int SEG, int BS>
__device__ void head_cta(const GenParams& P, const CtaDesc& D, float* smem) {
    constexpr int NC = SEG * BS;
    constexpr int HP = NCH * 64;
    constexpr int HIST = 32;                                  // per-utterance input history (>= K + SEG)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = P.H, S = P.S, Sp = P.Sp, K = P.K, L = P.L, n1 = P.n1, n1p = P.n1p, n2 = P.n2;
    const int hc = D.layer;
    const bool is_c = P.kind == 0;

    // ---- one-time weight staging
    auto stage = [&](float* dst, const float* src, int n) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        float4* d4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n / 4; i += SWN_NTHREADS) d4[i] = s4[i];
    };
    const float* WA = P.wbuf + D.wa_off;
    const float* WB = P.wbuf + D.wb_off;
    const float* WC = P.wbuf + D.wc_off;
    if (D.wa_in_smem) { stage(smem + D.sm_wa, WA, D.na * HP); WA = smem + D.sm_wa; }
    if (D.wb_in_smem) { stage(smem + D.sm_wb, WB, D.nb * Sp); WB = smem + D.sm_wb; }
    if (D.wc_in_smem) { stage(smem + D.sm_wc, WC, D.nc * n1p); WC = smem + D.sm_wc; }
    const float* CA = P.wbuf + D.ca_off;
    const float* BB = P.wbuf + D.bb_off;
    const float* BC = P.wbuf + D.bc_off;
    float* x0 = smem + D.sm_x0;       // [BS][HP]   last-column output of layer L-1
    float* x1 = smem + D.sm_x1;       // [BS][Sp]   relu(sum of skips)
    float* x2 = smem + D.sm_x2;       // [BS][n1p]  relu(out_1)
    float* misc = smem + D.sm_misc;
    // misc layout
    float* sacc = misc;                               // [BS][na]      partial skip sums from the layer groups
    float* oc = sacc + BS * ((D.na + 3) & ~3);        // [BS][max(nc,4)] out_2 rows of this CTA
    float* hist = oc + BS * ((D.nc + 3) & ~3);        // [M][BS][HIST] input history (samples or indices as float bits)
    float* lpb = hist + (size_t)P.M * BS * HIST;      // [M][BS][8]    LPC feedback buffers
    float* snew = lpb + (size_t)P.M * BS * 8;         // [BS][SEG]     samples of the current step
    int* inew = reinterpret_cast<int*>(snew + BS * SEG);   // [BS]     sampled index
    for (int i = tid; i < BS * HP; i += SWN_NTHREADS) x0[i] = 0.f;
    for (int i = tid; i < BS * Sp; i += SWN_NTHREADS) x1[i] = 0.f;
    for (int i = tid; i < BS * n1p; i += SWN_NTHREADS) x2[i] = 0.f;
    for (int i = tid; i < P.M * BS * HIST; i += SWN_NTHREADS)
        hist[i] = is_c ? 0.f : __int_as_float(P.Q / 2);       // pad value (cswnv:302 / dswnv:308)
    for (int i = tid; i < P.M * BS * 8; i += SWN_NTHREADS) lpb[i] = 0.f;
    __syncthreads();
    if (P.seedw) {
        // a_full = pad(R) ++ seed(SEG) ++ generated: the seed occupies generated-indices -SEG..-1
        for (int e = tid; e < P.M * BS * SEG; e += SWN_NTHREADS) {
            const int mb = e / SEG, j = e - mb * SEG;
            const int bglob = min(mb, P.B - 1);
            const float v = is_c ? ((const float*)P.seedw)[(size_t)bglob * SEG + j]
                                 : __int_as_float(pos_mod(((const long long*)P.seedw)[bglob], P.Q));
            hist[(size_t)mb * HIST + ((j - SEG + HIST) % HIST)] = v;
            // lp_x_buffer = last lpc samples of the padded seed (cswnv_shift1.py:317-319)
            const int mm = P.lpc - SEG + j;
            if (is_c && mm >= 0) lpb[(size_t)mb * 8 + mm] = v;
        }
        __syncthreads();
    }

    const unsigned long long* lastl = P.ring_ll + (size_t)L * P.M * NC * H;

    // Input stage for block `itn` of slot m: h0 = softsign(causal(...)) for the rows this CTA
    // owns (cswnv_shift1.py:352 / dswnv.py:345), published as the input of layer 0.
    auto input_stage = [&](int itn, int m) {
        const int t0n = P.t_first + itn * SEG;
        const uint32_t tagn = (uint32_t)itn + 1u;
        const float* hs = hist + (size_t)m * BS * HIST;
        unsigned long long* dst = P.ring_ll + (size_t)m * NC * H;       // layer 0 input
        for (int e = tid; e < D.ni * NC; e += SWN_NTHREADS) {
            const int r = D.i0 + e % D.ni, c = e / D.ni;
            const int b = c / SEG, s = c - b * SEG;
            const int t = t0n + s;
            float v = 0.f;
            if (t >= 0) {
                float acc = P.in_bias[r];
                for (int j = 0; j < K; ++j) {
                    const int tau = t - (K - 1) + j;                     // hidden time of the tap
                    if (tau < 0) continue;                               // conv zero padding
                    if (is_c) {
                        // newest sample seen by hidden time tau is a_full[tau+SEG-1]; gen index = tau-R-1
                        const int n = tau - P.R - 1;
                        const float a = (n >= -SEG) ? hs[b * HIST + ((n + HIST) % HIST)] : 0.f;
                        acc += P.in_tab[(j * 2) * H + r] + P.in_tab[(j * 2 + 1) * H + r] * a;
                    } else {
                        const int n = tau - P.R - 1;
                        const int q = (n >= -SEG) ? __float_as_int(hs[b * HIST + ((n + HIST) % HIST)]) : P.Q / 2;
                        acc += __ldg(P.in_tab + ((size_t)j * P.Q + q) * H + r);
                    }
                }
                v = softsignf_(acc);
            }
            st_ll(dst + (size_t)c * H + r, v, tagn);
        }
        if (P.audio_in && hc == 0) {
            // index seen at the (single) column of block itn: a_full[t]
            for (int b = tid; b < BS; b += SWN_NTHREADS) {
                const int n = t0n - P.R - 1;
                const int q = (n >= -SEG) ? __float_as_int(hs[b * HIST + ((n + HIST) % HIST)]) : P.Q / 2;
                st_ll_u32(P.idx_ll + m * BS + b, (uint32_t)q, tagn);
            }
        }
    };

    // prologue: inputs of block 0 (all padding); a head group only serves its own pipeline slots
    const int hg = D.hgroup;
    for (int m = hg; m < P.M; m += P.NHG) input_stage(0, m);

    for (int it = 0; it < P.n_it; ++it) {
        const bool active = it >= P.it0;
        const uint32_t tag = (uint32_t)it + 1u;
        const int istep = it - P.it0;                                    // generation step index
        for (int m = hg; m < P.M; m += P.NHG) {
            bool ok = true;
            // ---- wait for layer L-1 (flow control in warm-up, data when active)
            {
                const unsigned long long* src = lastl + (size_t)m * NC * H;
                sentinel_wait(src, tag, P);
                for (int i = tid; i < NC * H; i += SWN_NTHREADS) {
                    uint32_t bits;
                    ok = ll_wait(src + i, tag, bits, P, 200) && ok;
                    const int c = i / H, k = i - c * H;
                    const int b = c / SEG;
                    if (c - b * SEG == SEG - 1) x0[b * HP + k] = __uint_as_float(bits);
                }
            }
            if (active) {
                // partial skip sums published by layer groups 1..L-1 (rows of this CTA only)
                for (int e = tid; e < BS * D.na; e += SWN_NTHREADS) {
                    const int b = e / D.na, r = e - b * D.na;
                    float acc = 0.f;
                    for (int l2 = 0; l2 + 1 < L; ++l2) {
                        uint32_t bits;
                        ok = ll_wait(P.skip_ll + (((size_t)l2 * P.M + m) * BS + b) * S + D.a0 + r, tag, bits, P, 210) && ok;
                        acc += __uint_as_float(bits);
                    }
                    sacc[b * D.na + r] = acc;
                }
            }
            if (__syncthreads_or(!ok)) return;
            SWN_PROF(0);
            if (active) {
                // ---- stage A: relu(sum_l skip_l)   (cswnv_shift1.py:367 inner sum)
                for (int r = warp; r < D.na; r += SWN_NWARPS) {
                    for (int b = 0; b < BS; ++b) {
                        const float v = warp_sum(dot_row(WA + (size_t)r * HP, x0 + b * HP, HP, lane));
                        if (lane == 0) {
                            const float sv = v + CA[r] + sacc[b * D.na + r];
                            st_ll(P.hx1 + ((size_t)m * BS + b) * S + D.a0 + r, fmaxf(sv, 0.f), tag);
                        }
                    }
                }
                for (int i = tid; i < BS * S; i += SWN_NTHREADS) {
                    uint32_t bits;
                    ok = ll_wait(P.hx1 + (size_t)m * BS * S + i, tag, bits, P, 220) && ok;
                    const int b = i / S, k = i - b * S;
                    x1[b * Sp + k] = __uint_as_float(bits);
                }
                if (__syncthreads_or(!ok)) return;
                SWN_PROF(1);
                // ---- stage B: relu(out_1(.))
                for (int r = warp; r < D.nb; r += SWN_NWARPS) {
                    for (int b = 0; b < BS; ++b) {
                        const float v = warp_sum(dot_row(WB + (size_t)r * Sp, x1 + b * Sp, Sp, lane));
                        if (lane == 0) st_ll(P.hx2 + ((size_t)m * BS + b) * n1 + D.b0 + r, fmaxf(v + BB[r], 0.f), tag);
                    }
                }
                for (int i = tid; i < BS * n1; i += SWN_NTHREADS) {
                    uint32_t bits;
                    ok = ll_wait(P.hx2 + (size_t)m * BS * n1 + i, tag, bits, P, 230) && ok;
                    const int b = i / n1, k = i - b * n1;
                    x2[b * n1p + k] = __uint_as_float(bits);
                }
                if (__syncthreads_or(!ok)) return;
                SWN_PROF(2);
                // ---- stage C: out_2 rows
                const int ncp = (D.nc + 3) & ~3;
                for (int r = warp; r < D.nc; r += SWN_NWARPS) {
                    for (int b = 0; b < BS; ++b) {
                        const float v = warp_sum(dot_row(WC + (size_t)r * n1p, x2 + b * n1p, n1p, lane));
                        if (lane == 0) oc[b * ncp + r] = v + BC[r];
                    }
                }
                __syncthreads();
                float* hs = hist + (size_t)m * BS * HIST;
                if (is_c) {
                    // every head CTA holds all n2 rows and samples redundantly (bitwise identical)
                    if (tid < BS) {
                        const int b = tid;
                        int bglob = m * BS + b;
                        const bool real = bglob < P.B;
                        if (!real) bglob = P.B - 1;
                        const long long nbase = (long long)istep * SEG;
                        float eps[SEG], smp[SEG];
#pragma unroll
                        for (int j = 0; j < SEG; ++j)
                            eps[j] = P.noise ? P.noise[(size_t)bglob * P.n_out + nbase + j]
                                             : laplace_eps(P.seed, bglob, nbase + j);
                        const float* forced = P.forced ? (const float*)P.forced + (size_t)bglob * P.n_out + nbase : nullptr;
                        laplace_step<SEG>(P.lpc, oc + b * ncp, lpb + ((size_t)m * BS + b) * 8, eps, forced, smp);
#pragma unroll
                        for (int j = 0; j < SEG; ++j) {
                            const float fed = forced ? forced[j] : smp[j];
                            hs[b * HIST + (int)((nbase + j) % HIST)] = fed;
                            if (hc == 0 && real) ((float*)P.out)[(size_t)bglob * P.n_out + nbase + j] = smp[j];
                        }
                    }
                    if (hc == 0 && P.head_out) {
                        for (int e = tid; e < BS * n2; e += SWN_NTHREADS) {
                            const int b = e / n2, r = e - b * n2;
                            const int bglob = m * BS + b;
                            if (bglob < P.B) P.head_out[((size_t)bglob * P.n_iter + istep) * n2 + r] = oc[b * ncp + r];
                        }
                    }
                    __syncthreads();
                } else {
                    // Gumbel-max over this CTA's rows: argmax_k(logit_k - log q_k)  (dswnv.py:363-369)
                    for (int b = warp; b < BS; b += SWN_NWARPS) {
                        int bglob = m * BS + b;
                        const bool real = bglob < P.B;
                        if (!real) bglob = P.B - 1;
                        float best = -CUDART_INF_F; int bi = 0x7fffffff;
                        for (int r = lane; r < D.nc; r += 32) {
                            const float o = oc[b * ncp + r];
                            const int row = D.c0 + r;
                            if (P.head_out && real) P.head_out[((size_t)bglob * P.n_iter + istep) * n2 + row] = o;
                            const float q = P.noise ? P.noise[((size_t)bglob * P.n_out + istep) * P.Q + row]
                                                    : softmax_q(P.seed, bglob, istep, row);
                            const float g = o - logf(q);
                            if (g > best || (g == best && row < bi)) { best = g; bi = row; }
                        }
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) {
                            const float ob = __shfl_xor_sync(0xffffffffu, best, o2);
                            const int oi = __shfl_xor_sync(0xffffffffu, bi, o2);
                            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
                        }
                        if (lane == 0) {
                            unsigned long long* d3 = P.hx3 + (((size_t)m * P.NH + hc) * BS + b) * 2;
                            st_ll(d3, best, tag);
                            st_ll_u32(d3 + 1, (uint32_t)bi, tag);
                        }
                    }
                    // combine the NH candidates (every head CTA does it, identically)
                    if (tid < BS) {
                        const int b = tid;
                        float best = -CUDART_INF_F; int bi = 0x7fffffff;
                        for (int c = 0; c < P.NH; ++c) {
                            const unsigned long long* s3 = P.hx3 + (((size_t)m * P.NH + c) * BS + b) * 2;
                            uint32_t gb, ib;
                            ok = ll_wait(s3, tag, gb, P, 240) && ok;
                            ok = ll_wait(s3 + 1, tag, ib, P, 241) && ok;
                            const float g = __uint_as_float(gb);
                            if (g > best || (g == best && (int)ib < bi)) { best = g; bi = (int)ib; }
                        }
                        const int bglob = m * BS + b;
                        int fed = bi;
                        if (bglob < P.B) {
                            if (hc == 0) ((long long*)P.out)[(size_t)bglob * P.n_out + istep] = bi;
                            if (P.forced) fed = clamp_idx(((const long long*)P.forced)[(size_t)bglob * P.n_out + istep], P.Q);
                        }
                        hs[b * HIST + (istep % HIST)] = __int_as_float(fed);
                    }
                    if (__syncthreads_or(!ok)) return;
                }
            }
            // ---- inputs of the next block
            if (it + 1 < P.n_it) input_stage(it + 1, m);
            SWN_PROF(3);
        }
    }
}

// ------------------------------------------------------------------ kernel
template <int NCH, int SEG, int BS>
__global__ void __launch_bounds__(SWN_NTHREADS, 1) swn_generate_kernel(const __grid_constant__ GenParams P) {
    extern __shared__ __align__(16) float smem[];
    const CtaDesc& D = P.desc[blockIdx.x];
    if (D.role == ROLE_LAYER) {
        if (D.npairs > 0) layer_cta<NCH, SEG, BS>(P, D, smem);
    } else if (D.role == ROLE_HEAD) {
        head_cta<NCH, SEG, BS>(P, D, smem);
    }
}

}  // namespace swn
