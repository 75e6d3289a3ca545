// Instantiations of the pipelined batch kernel (variant 3) for one `seg`
// (compiled once per SWN_SEG, see Makefile).
#include "swn_batch.cuh"
#include "swn_launch.h"

#ifndef SWN_SEG
#error "compile with -DSWN_SEG=<1,2,3,4,5,10>"
#endif

namespace swn {

template <int SEG, int BS>
static cudaError_t launch_batch_inst(const BatchParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = swn_batch_kernel<SEG, BS>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, SWN_NTHREADS, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    void* args[] = {(void*)&p};
    return cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(SWN_NTHREADS), args, smem, st);
}

#define SWN_CAT2(a, b) a##b
#define SWN_CAT(a, b) SWN_CAT2(a, b)
#define B3_CASE(B) \
    if (bs == B) return &launch_batch_inst<SWN_SEG, B>;

batch_launch_fn SWN_CAT(find_batch_kernel_seg, SWN_SEG)(int bs) {
    B3_CASE(1) B3_CASE(2) B3_CASE(4)
#if SWN_SEG == 1 || SWN_SEG == 2 || SWN_SEG == 5
    B3_CASE(8)
#endif
#if SWN_SEG == 1 || SWN_SEG == 5
    B3_CASE(16)
#endif
    return nullptr;
}

}  // namespace swn
