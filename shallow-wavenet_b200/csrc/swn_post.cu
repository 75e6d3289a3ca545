// Noise-shaping restore filter on the GPU: what src/bin/noise_shaping.py:52-87 (synthesis_diff:
// pysptk mc2b + MLSADF + Synthesizer) and src/bin/feature_extract.py:57-76 (low_cut_filter: 255-tap FIR) do to every
// generated waveform in recipe stages 6 / 9, as a post-filter over the whole decoded batch.
//
//   mlsa_kernel   one warp per utterance.  The MLSA filter is a time recursion (no parallelism along time), but inside a
//                 sample its `pd` Pade stages only read the PREVIOUS sample's stage outputs (SPTK mlsadf2 walks i = pd..1
//                 and stage i reads pt[i-1] before it is overwritten): lanes 1..pd run the pd order-m all-pass chains
//                 (mlsafir) side by side, lane 0 combines them in SPTK's order.  The per-sample shift of each delay line
//                 is a ring-buffer offset, not a copy.  Filter coefficients move linearly inside a frame by repeated
//                 addition, exactly as pysptk's Synthesizer does.  All arithmetic is float64 with explicitly rounded
//                 add / mul (no FMA contraction), so that the result equals the CPU restatement up to exp()'s last bit.
//   fir_kernel    y[n] = sum_k h[k] x[n-k], one thread per output sample, taps in shared memory.
//
// Utterances are independent: the batch shards over GPUs like the generation itself (no collective).
#include <cuda_runtime.h>
#include <stdint.h>

#include "swn_post.h"

namespace swn {

namespace {

// SPTK 3.11 mlsadf.c: Pade approximants of exp(x), rows of order 1..5 back to back (row pd at pd*(pd+1)/2)
__constant__ double c_pade[21] = {1.0,
                                  1.0, 0.0,
                                  1.0, 0.0, 0.0,
                                  1.0, 0.0, 0.0, 0.0,
                                  1.0, 0.4999273, 0.1067005, 0.01170221, 0.0005656279,
                                  1.0, 0.4999391, 0.1107098, 0.01369984, 0.0009564853, 0.00003041721};

constexpr int NS_WARPS = 4;          // utterances per block
constexpr int NS_MAXM = 62;          // mel-cepstrum order (two coefficients per lane)
constexpr int NS_RING = 64;          // ring slots of one delay line (>= order + 1, power of two)

struct MlsaParams {
    const double* x; double* y; long long stride;
    const long long* n_samples;      // [B]
    const int* n_frames;             // [B]
    const double* b;                 // [b_utts][b_frames][m+1] (mc2b applied)
    int b_utts, b_frames, m, pd, hop, B;
    double alpha;
};

__global__ void __launch_bounds__(NS_WARPS * 32) mlsa_kernel(const MlsaParams P) {
    __shared__ double s_coef[NS_WARPS][NS_MAXM + 2];                 // interpolated b of the current sample
    __shared__ double s_ring[NS_WARPS][5][NS_RING];                  // delay lines d[2..m+1] of the pd stages
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int u = blockIdx.x * NS_WARPS + wib;
    if (u >= P.B) return;
    const int m = P.m, pd = P.pd, hop = P.hop;
    const double a = P.alpha, aa = 1.0 - a * a;
    const double* ppade = c_pade + pd * (pd + 1) / 2;
    const long long N = P.n_samples[u];
    const double* x = P.x + (size_t)u * P.stride;
    double* y = P.y + (size_t)u * P.stride;
    const double* bt = P.b + (size_t)(P.b_utts > 1 ? u : 0) * P.b_frames * (m + 1);
    double* coef = s_coef[wib];
    for (int i = lane; i < 5 * NS_RING; i += 32) (&s_ring[wib][0][0])[i] = 0.0;
    // frames that are synthesised: i < n_frames and (i + 1) * hop < N (pysptk Synthesizer.synthesis)
    long long nf = P.n_frames ? P.n_frames[u] : (P.b_frames > 1 ? P.b_frames : (1LL << 40));
    long long full = N > 0 ? (N - 1) / hop : 0;                     // largest i + 1 with (i + 1) * hop < N
    if (nf > full) nf = full;
    const long long n_syn = nf * hop;
    // per-lane coefficient slots: indices lane and lane + 32
    double prev0 = 0.0, prev1 = 0.0, c0 = 0.0, c1 = 0.0, sl0 = 0.0, sl1 = 0.0;
    const int i0 = lane, i1 = lane + 32;
    if (i0 <= m) prev0 = bt[i0];
    if (i1 <= m) prev1 = bt[i1];
    // mlsadf1 state (every lane keeps a copy: uniform arithmetic), mlsadf2 state: lane i holds pt2[i], lanes 1..pd their d0/d1
    double d1s[6] = {0, 0, 0, 0, 0, 0}, pt1[6] = {0, 0, 0, 0, 0, 0};
    double pt2 = 0.0, sd1 = 0.0;
    int base = 0;                                                    // ring offset: logical d[i] (i >= 2) = ring[(base + i) & 63]
    double* ring = s_ring[wib][lane >= 1 && lane <= pd ? lane - 1 : 0];
    double xin = 0.0, yout = 0.0;
    __syncwarp();
    for (long long n = 0; n < n_syn; ++n) {
        const int k = (int)(n & 31);
        if (k == 0) xin = (n + lane < N) ? x[n + lane] : 0.0;        // 32 samples per coalesced load
        const long long f = n / hop;
        if (n - f * hop == 0) {
            // frame start: coefficients run from the previous frame's vector to this frame's
            const double* cur = bt + (size_t)(P.b_frames > 1 ? f : 0) * (m + 1);
            const double q0 = i0 <= m ? cur[i0] : 0.0, q1 = i1 <= m ? cur[i1] : 0.0;
            sl0 = __ddiv_rn(__dsub_rn(q0, prev0), (double)hop); sl1 = __ddiv_rn(__dsub_rn(q1, prev1), (double)hop);
            c0 = prev0; c1 = prev1;
            prev0 = q0; prev1 = q1;
        }
        if (i0 <= m) coef[i0] = c0;
        if (i1 <= m) coef[i1] = c1;
        __syncwarp();
        const double xs = __shfl_sync(0xffffffffu, xin, k);
        double xv = __dmul_rn(xs, exp(coef[0]));                     // source[i] * exp(interpolated_coef[0])
        // ---- mlsadf1 (first-order section on b[1])
        {
            const double b1 = coef[1];
            double out = 0.0;
#pragma unroll
            for (int i = 5; i >= 1; --i) {
                if (i <= pd) {
                    d1s[i] = __dadd_rn(__dmul_rn(aa, pt1[i - 1]), __dmul_rn(a, d1s[i]));
                    pt1[i] = __dmul_rn(d1s[i], b1);
                    const double v = __dmul_rn(pt1[i], ppade[i]);
                    xv = (i & 1) ? __dadd_rn(xv, v) : __dsub_rn(xv, v);
                    out = __dadd_rn(out, v);
                }
            }
            pt1[0] = xv;
            xv = __dadd_rn(out, xv);
        }
        // ---- mlsadf2: stage i (lane i) filters the PREVIOUS sample's pt2[i-1]
        const double pin = __shfl_up_sync(0xffffffffu, pt2, 1);
        double ystage = 0.0;
        if (lane >= 1 && lane <= pd) {
            // mlsafir: d[0] = x; d[1] = aa d[0] + a d[1]; d[i] += a (d[i+1] - d[i-1]), y += d[i] b[i]; then d shifts up by one
            sd1 = __dadd_rn(__dmul_rn(aa, pin), __dmul_rn(a, sd1));
            // (the old delay-line values and the coefficients of four taps are fetched ahead of the dependent chain:
            // issued one by one behind each store, their shared-memory latency was on it - 6.7 k cycles per sample)
            double prevd = sd1;
            double di = ring[(base + 2) & (NS_RING - 1)];
            for (int i = 2; i <= m; i += 4) {
                double o[4], cf[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    o[j] = ring[(base + i + j + 1) & (NS_RING - 1)];          // old d[i+j+1] (slots up to m+4 exist: 64-slot ring, m <= 62 ... masked)
                    cf[j] = coef[i + j <= m ? i + j : m];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (i + j <= m) {
                        const double cur = __dadd_rn(di, __dmul_rn(a, __dsub_rn(o[j], prevd)));
                        ystage = __dadd_rn(ystage, __dmul_rn(cur, cf[j]));
                        ring[(base + i + j) & (NS_RING - 1)] = cur;
                        prevd = cur; di = o[j];
                    }
                }
            }
            // shift: new d[i] = d[i-1] (i = m+1 .. 2) = one step of the ring offset, plus d[2] <- d[1]
            ring[(base - 1 + 2) & (NS_RING - 1)] = sd1;
        }
        base = (base - 1) & (NS_RING - 1);
        if (lane >= 1 && lane <= pd) pt2 = ystage;
        // ---- lane 0 combines the stages in SPTK's order (i = pd .. 1)
        double out2 = 0.0;
#pragma unroll
        for (int i = 5; i >= 1; --i) {
            const double pti = __shfl_sync(0xffffffffu, pt2, i);
            if (i <= pd) {
                const double v = __dmul_rn(pti, ppade[i]);
                xv = (i & 1) ? __dadd_rn(xv, v) : __dsub_rn(xv, v);
                out2 = __dadd_rn(out2, v);
            }
        }
        if (lane == 0) pt2 = xv;
        out2 = __dadd_rn(out2, xv);
        if (lane == k) yout = out2;
        if (k == 31 || n + 1 == n_syn) {
            const long long n0 = n - k;
            if (lane <= k) y[n0 + lane] = yout;                      // 32 samples per coalesced store
        }
        c0 = __dadd_rn(c0, sl0); c1 = __dadd_rn(c1, sl1);           // interpolated_coef += slope
        __syncwarp();
    }
    // the trailing partial frame is not synthesised (pysptk leaves it uninitialised): zeros
    for (long long n = n_syn + lane; n < N; n += 32) y[n] = 0.0;
}

__global__ void fir_kernel(const double* __restrict__ x, double* __restrict__ y, long long stride, const long long* __restrict__ n_samples,
                           const double* __restrict__ h, int taps) {
    extern __shared__ double s_h[];
    for (int i = threadIdx.x; i < taps; i += blockDim.x) s_h[i] = h[i];
    __syncthreads();
    const int u = blockIdx.y;
    const long long N = n_samples[u];
    const double* xu = x + (size_t)u * stride;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += (long long)gridDim.x * blockDim.x) {
        double acc = 0.0;
        const int kmax = n + 1 < taps ? (int)(n + 1) : taps;
        for (int k = 0; k < kmax; ++k) acc = __dadd_rn(acc, __dmul_rn(s_h[k], xu[n - k]));
        y[(size_t)u * stride + n] = acc;
    }
}

}  // namespace

cudaError_t launch_mlsa(const double* x, double* y, long long stride, const long long* n_samples_dev, const int* n_frames_dev,
                        const double* b_dev, int b_utts, int b_frames, int order, int pd, double alpha, int hop, int B,
                        cudaStream_t st) {
    if (order < 2 || order > NS_MAXM || (pd != 4 && pd != 5) || hop < 1 || B < 1) return cudaErrorInvalidValue;
    MlsaParams P;
    P.x = x; P.y = y; P.stride = stride; P.n_samples = n_samples_dev; P.n_frames = n_frames_dev; P.b = b_dev;
    P.b_utts = b_utts; P.b_frames = b_frames; P.m = order; P.pd = pd; P.hop = hop; P.B = B; P.alpha = alpha;
    mlsa_kernel<<<(B + NS_WARPS - 1) / NS_WARPS, NS_WARPS * 32, 0, st>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_fir(const double* x, double* y, long long stride, const long long* n_samples_dev, long long n_max,
                       const double* h_dev, int taps, int B, cudaStream_t st) {
    if (taps < 1 || taps > 4096 || B < 1) return cudaErrorInvalidValue;
    const int threads = 256;
    long long blocks = (n_max + threads - 1) / threads;
    if (blocks < 1) blocks = 1;
    if (blocks > 4096) blocks = 4096;
    fir_kernel<<<dim3((unsigned)blocks, (unsigned)B), threads, taps * sizeof(double), st>>>(x, y, stride, n_samples_dev, h_dev, taps);
    return cudaGetLastError();
}

}  // namespace swn
