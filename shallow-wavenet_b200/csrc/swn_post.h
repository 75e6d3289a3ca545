// Launchers of the noise-shaping post-filter kernels (swn_post.cu).
#pragma once
#include <cuda_runtime.h>

namespace swn {

cudaError_t launch_mlsa(const double* x, double* y, long long stride, const long long* n_samples_dev, const int* n_frames_dev,
                        const double* b_dev, int b_utts, int b_frames, int order, int pd, double alpha, int hop, int B,
                        cudaStream_t st);
cudaError_t launch_fir(const double* x, double* y, long long stride, const long long* n_samples_dev, long long n_max,
                       const double* h_dev, int taps, int B, cudaStream_t st);

}  // namespace swn
