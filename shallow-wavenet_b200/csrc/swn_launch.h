// Host-side lookup of the instantiated persistent kernels.
#pragma once
#include <cuda_runtime.h>
#include "swn_common.h"

namespace swn { struct CoreParams; struct BatchParams; }

namespace swn {

typedef cudaError_t (*launch_fn)(const GenParams& p, int grid, size_t smem, cudaStream_t st);

launch_fn find_kernel_nch1(int seg, int bs);
launch_fn find_kernel_nch2(int seg, int bs);
launch_fn find_kernel_nch3(int seg, int bs);
launch_fn find_kernel_nch4(int seg, int bs);

struct core_fns {
    cudaError_t (*prepare)(size_t smem, int* max_clusters);
    cudaError_t (*launch)(const CoreParams& p, int grid, size_t smem, cudaStream_t st);
};
core_fns find_core_kernel_nch1(int bs, bool multi = false);
core_fns find_core_kernel_nch2(int bs, bool multi = false);
core_fns find_core_kernel_nch3(int bs, bool multi = false);
core_fns find_core_kernel_nch4(int bs, bool multi = false);
// seg > 1 cluster kernel (swn_coreseg.cuh), one utterance per launch; instantiated for seg 2 .. 5
core_fns find_coreseg_kernel_nch1(int seg);
core_fns find_coreseg_kernel_nch2(int seg);
core_fns find_coreseg_kernel_nch3(int seg);

// ---- variant 3: pipelined batch kernel (swn_batch.cuh), one object file per seg
typedef cudaError_t (*batch_launch_fn)(const BatchParams& p, int grid, size_t smem, cudaStream_t st);
batch_launch_fn find_batch_kernel_seg1(int bs);
batch_launch_fn find_batch_kernel_seg2(int bs);
batch_launch_fn find_batch_kernel_seg3(int bs);
batch_launch_fn find_batch_kernel_seg4(int bs);
batch_launch_fn find_batch_kernel_seg5(int bs);
batch_launch_fn find_batch_kernel_seg10(int bs);
inline batch_launch_fn find_batch_kernel(int seg, int bs) {
    switch (seg) {
        case 1: return find_batch_kernel_seg1(bs);
        case 2: return find_batch_kernel_seg2(bs);
        case 3: return find_batch_kernel_seg3(bs);
        case 4: return find_batch_kernel_seg4(bs);
        case 5: return find_batch_kernel_seg5(bs);
        case 10: return find_batch_kernel_seg10(bs);
        default: return nullptr;
    }
}
// utterances per pipeline slot instantiated for a given seg (descending), terminated by 0
inline const int* batch3_tiles_for_seg(int seg) {
    static const int a[] = {16, 8, 4, 2, 1, 0};
    static const int b[] = {8, 4, 2, 1, 0};
    static const int c[] = {4, 2, 1, 0};
    static const int none[] = {0};
    switch (seg) {
        case 1: case 5: return a;
        case 2: return b;
        case 3: case 4: case 10: return c;
        default: return none;
    }
}

// batch tiles instantiated for a given seg (descending), terminated by 0
inline const int* batch_tiles_for_seg(int seg) {
    static const int s1[] = {8, 4, 2, 1, 0};
    static const int s2[] = {4, 2, 1, 0};
    static const int s5[] = {2, 1, 0};
    static const int sx[] = {1, 0};
    static const int none[] = {0};
    switch (seg) {
        case 1: return s1;
        case 2: return s2;
        case 5: return s5;
        case 3: case 4: case 10: return sx;
        default: return none;
    }
}

// conditioning kernels (swn_cond.cu)
cudaError_t launch_aux_conv(const float* in, const float* w_t, const float* bias, float* out, int B, int Ci, int Co,
                            int T, int ks, int dil, int pad, int transpose_out, cudaStream_t st);
cudaError_t launch_proj_gemm(const float* A, const float* Bm, float* C, long long M, int N, int K, cudaStream_t st);

cudaError_t measure_smem_bandwidth(int n_sm, double* gb_per_s, double* sm_mhz);

}  // namespace swn
