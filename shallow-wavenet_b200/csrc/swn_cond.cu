// Conditioning kernels, run once per batch of utterances (frame rate, not sample rate).
//
// Reference: x = upsampling(conv_aux(scale_in(aux))), replicate left pad, seg-window stacking and
// the per-layer in_x[l] 1x1 convs evaluated at every time step (src/nets/cswnv_shift1.py:297-311
// and :282; src/nets/dswnv.py:302-309 and :291).  All of that is affine in the frame-rate
// features y = conv_aux(scale_in(aux)), and the upsampler only multiplies by a per-phase gain,
// so the whole chain collapses to
//     in_x[l](x')[t] = const_l + sum_s gain[phase(t+s)] * Pt[frame(t+s)][l][s][:]
// with Pt = y^T . Wp computed here once per frame (DESIGN.md "conditioning hoist").
#include <cuda_runtime.h>
#include <algorithm>
#include "swn_launch.h"

namespace swn {

// ---------------------------------------------------------------- two-sided dilated conv (+ scale_in as ks=1)
// in [B][Ci][T], w_t [Ci][ks][Co] (transposed on the host for coalescing), zero padding of the
// INPUT outside [0,T) exactly like nn.Conv1d(padding=pad) (TwoSidedDilConv1d, cswnv_shift1.py:95-127).
// out: [B][Co][T], or frame-major [B][T][ldo] when transpose_out (ldo >= Co, padding columns zeroed).
constexpr int CONV_FT = 16;
constexpr int CONV_THREADS = 256;

__global__ void __launch_bounds__(CONV_THREADS) aux_conv_kernel(const float* __restrict__ in, const float* __restrict__ w_t,
                                                               const float* __restrict__ bias, float* __restrict__ out,
                                                               int Ci, int Co, int T, int ks, int dil, int pad,
                                                               int transpose_out, int ldo) {
    extern __shared__ float tile[];                 // [Ci][W], W = CONV_FT + (ks-1)*dil
    const int W = CONV_FT + (ks - 1) * dil;
    const int b = blockIdx.y, f0 = blockIdx.x * CONV_FT;
    const float* inb = in + (size_t)b * Ci * T;
    for (int i = threadIdx.x; i < Ci * W; i += CONV_THREADS) {
        const int ci = i / W, x = i - ci * W;
        const int f = f0 + x - pad;
        tile[i] = (f >= 0 && f < T) ? inb[(size_t)ci * T + f] : 0.f;
    }
    __syncthreads();
    for (int co = threadIdx.x; co < (transpose_out ? ldo : Co); co += CONV_THREADS) {
        float acc[CONV_FT];
        const bool real = co < Co;
        const float bv = real ? bias[co] : 0.f;
#pragma unroll
        for (int f = 0; f < CONV_FT; ++f) acc[f] = bv;
        if (real) {
            for (int ci = 0; ci < Ci; ++ci) {
                const float* trow = tile + ci * W;
                for (int kk = 0; kk < ks; ++kk) {
                    const float w = __ldg(w_t + ((size_t)ci * ks + kk) * Co + co);
                    const float* tp = trow + kk * dil;
#pragma unroll
                    for (int f = 0; f < CONV_FT; ++f) acc[f] = fmaf(w, tp[f], acc[f]);
                }
            }
        }
        if (transpose_out) {
#pragma unroll
            for (int f = 0; f < CONV_FT; ++f)
                if (f0 + f < T) out[((size_t)b * T + f0 + f) * ldo + co] = acc[f];
        } else {
#pragma unroll
            for (int f = 0; f < CONV_FT; ++f)
                if (f0 + f < T) out[((size_t)b * Co + co) * T + f0 + f] = acc[f];
        }
    }
}

cudaError_t launch_aux_conv(const float* in, const float* w_t, const float* bias, float* out, int B, int Ci, int Co,
                            int T, int ks, int dil, int pad, int transpose_out, cudaStream_t st) {
    const int W = CONV_FT + (ks - 1) * dil;
    const size_t smem = (size_t)Ci * W * sizeof(float);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(aux_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    const int ldo = transpose_out ? ((Co + 15) / 16) * 16 : Co;
    dim3 grid((T + CONV_FT - 1) / CONV_FT, B);
    aux_conv_kernel<<<grid, CONV_THREADS, smem, st>>>(in, w_t, bias, out, Ci, Co, T, ks, dil, pad, transpose_out, ldo);
    return cudaGetLastError();
}

// ---------------------------------------------------------------- projection GEMM  C[M][N] = A[M][K] . B[K][N]
// A = frame-major features (row stride K, K a multiple of 16, zero padded), B = stacked in_x
// weights of all layers / seg offsets (zero padded rows), C = the conditioning table.
// fp32 FFMA, 128x128x16 tiles, 8x8 register tile per thread, float4 global and shared accesses.
constexpr int GB_M = 128, GB_N = 128, GB_K = 16, GB_THREADS = 256;

__global__ void __launch_bounds__(GB_THREADS) proj_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Bm,
                                                              float* __restrict__ C, long long M, int N, int K) {
    __shared__ __align__(16) float As[GB_K][GB_M + 4];
    __shared__ __align__(16) float Bs[GB_K][GB_N];
    const int tid = threadIdx.x;
    const long long m0 = (long long)blockIdx.y * GB_M;
    const int n0 = blockIdx.x * GB_N;
    const int ty = tid / 16, tx = tid % 16;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // A tile: 128 rows x 16 k = 512 float4, two per thread; B tile: 16 k x 128 n = 512 float4
    const int a_row = tid / 4, a_k4 = (tid % 4) * 4;            // rows a_row and a_row+64
    const int b_k = tid / 32, b_n4 = (tid % 32) * 4;            // k rows b_k and b_k+8
    for (int k0 = 0; k0 < K; k0 += GB_K) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long m = m0 + a_row + h * 64;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M) v = *reinterpret_cast<const float4*>(A + m * K + k0 + a_k4);
            As[a_k4 + 0][a_row + h * 64] = v.x; As[a_k4 + 1][a_row + h * 64] = v.y;
            As[a_k4 + 2][a_row + h * 64] = v.z; As[a_k4 + 3][a_row + h * 64] = v.w;
            const int kk = b_k + h * 8;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + b_n4 < N) w = *reinterpret_cast<const float4*>(Bm + (size_t)(k0 + kk) * N + n0 + b_n4);
            *reinterpret_cast<float4*>(&Bs[kk][b_n4]) = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GB_K; ++kk) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (m >= M) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int n = n0 + h * 64 + tx * 4;
            if (n < N)
                *reinterpret_cast<float4*>(C + m * N + n) =
                    make_float4(acc[i][h * 4 + 0], acc[i][h * 4 + 1], acc[i][h * 4 + 2], acc[i][h * 4 + 3]);
        }
    }
}

cudaError_t launch_proj_gemm(const float* A, const float* Bm, float* C, long long M, int N, int K, cudaStream_t st) {
    if (K % GB_K != 0 || N % 4 != 0) return cudaErrorInvalidValue;
    dim3 grid((N + GB_N - 1) / GB_N, (unsigned)((M + GB_M - 1) / GB_M));
    proj_gemm_kernel<<<grid, GB_THREADS, 0, st>>>(A, Bm, C, M, N, K);
    return cudaGetLastError();
}


// ---------------------------------------------------------------- shared-memory bandwidth probe
// Measures the aggregate LDS bandwidth of the chip (one CTA per SM streaming a 192 KB buffer with
// 128-bit loads).  bench.py uses it as the denominator of the batch-1 step roofline: at batch 1
// every time step has to read all SMEM-resident weights once.
__global__ void __launch_bounds__(1024, 1) smem_bw_kernel(float* __restrict__ out, int iters, int n4) {
    extern __shared__ float4 sbuf[];
    for (int i = threadIdx.x; i < n4; i += blockDim.x) sbuf[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    float4 a0 = make_float4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
    for (int it = 0; it < iters; ++it) {
        for (int i = threadIdx.x; i + 3 * 1024 < n4; i += 4 * 1024) {
            const float4 v0 = sbuf[i], v1 = sbuf[i + 1024], v2 = sbuf[i + 2048], v3 = sbuf[i + 3072];
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
            a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a0.z + a0.w + a1.x + a1.y + a1.z + a1.w +
                                                 a2.x + a2.y + a2.z + a2.w + a3.x + a3.y + a3.z + a3.w;
}

cudaError_t measure_smem_bandwidth(int n_sm, double* gb_per_s, double* sm_mhz) {
    const int n4 = 12 * 1024;                       // 192 KB
    const size_t smem = (size_t)n4 * 16;
    cudaError_t e = cudaFuncSetAttribute(smem_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    float* out = nullptr;
    e = cudaMalloc((void**)&out, (size_t)n_sm * 1024 * sizeof(float));
    if (e != cudaSuccess) return e;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4000;
    smem_bw_kernel<<<n_sm, 1024, smem>>>(out, 200, n4);      // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        smem_bw_kernel<<<n_sm, 1024, smem>>>(out, iters, n4);
        cudaEventRecord(b);
        e = cudaEventSynchronize(b);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double bytes = (double)n_sm * iters * (double)(n4 / 4096 * 4096) * 16.0;
        best = std::max(best, bytes / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(out);
    *gb_per_s = best;
    if (sm_mhz) *sm_mhz = best * 1e9 / ((double)n_sm * 128.0) / 1e6;   // implied clock at 128 B/clk/SM
    return e;
}

}  // namespace swn
