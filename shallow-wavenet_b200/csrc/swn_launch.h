// Host-side lookup of the instantiated persistent kernels.
#pragma once
#include <cuda_runtime.h>
#include "swn_common.h"

namespace swn { struct CoreParams; }

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
core_fns find_core_kernel_nch1(int bs);
core_fns find_core_kernel_nch2(int bs);
core_fns find_core_kernel_nch3(int bs);
core_fns find_core_kernel_nch4(int bs);

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
