// Host/device interface of the time-parallel forward kernels (swn_fwd.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace swn {

enum { FWD_GATE = 0, FWD_RELU = 1, FWD_LINEAR = 2 };

// One launch of fwd_gemm_kernel: D[rows x (nbatch * T)] = W . X with a fused epilogue.
struct FwdGemm {
    const float* A;            // packed weights: [n_mt][n_kb][hi | lo][BK/4][16][8][4]  (pack_fwd_weights)
    const float* Bhi;          // activation planes, time-major [nbatch * T][ldb]
    const float* Blo;
    int32_t ldb, b_col0;       // first column of tap 0
    int32_t tap_col_stride;    // column offset between taps (0: same channels, shifted in time)
    int32_t C, Cpad;           // channels per tap, padded to a multiple of the k block
    int32_t ntaps, n_kb, n_mt;
    int32_t delay[16];         // rows (time steps) a tap looks back
    int32_t T, nbatch;
    int32_t mode, rows, H;
    const float* bias;         // GATE: dil_h bias [2H]; else [rows]
    const float* p;            // GATE: sample-rate conditioning of this layer [nbatch * T][2H]
    int32_t carry_col0;        // GATE: column block of the layer input inside Bhi / Blo
    float* out_hi;             // GATE / RELU: output planes [nbatch * T][ldo], written at out_col0 + row
    float* out_lo;
    int32_t ldo, out_col0;
    float* out_f;              // LINEAR: fp32 [nbatch * T][ldf]
    int32_t ldf;
    int32_t* status;           // [0] != 0: a barrier wait timed out
};

struct FwdInput {
    int32_t kind, H, K, Q, seg, T, Ta, nbatch, ld;
    const float* in_tab; const float* in_bias;
    const float* audio_f; const int32_t* audio_i;
    float* out_hi; float* out_lo;
};

struct FwdCond {
    int32_t H, L, l, seg, U, T, Tf, Ta, Q, nbatch;
    const float* P; const float* pconst; const float* gain;
    const float* xa_tab; const int32_t* audio_i;
    float* out;
};

cudaError_t launch_fwd_gemm(const FwdGemm& g, cudaStream_t st);
void pack_fwd_weights(const float* W, int ktot, int n_mt, const int* row_map, float* out);
int fwd_bk();
int fwd_bm();
cudaError_t launch_fwd_onehot(const float* x, int B, int Q, int T, int32_t* idx, int32_t* bad, cudaStream_t st);
cudaError_t launch_fwd_index(const long long* x, int Q, long long n, int32_t* idx, cudaStream_t st);
cudaError_t launch_fwd_input(const FwdInput& I, cudaStream_t st);
cudaError_t launch_fwd_cond(const FwdCond& C, cudaStream_t st);
cudaError_t launch_fwd_laplace(const float* out, int n2, int seg, long long rows, float* b, float* log_b, float* b_clip,
                               float* log_b_clip, cudaStream_t st);

}  // namespace swn
