// Instantiations of the persistent generation kernel for one hidden-size class
// (compiled once per SWN_NCH = padded hidden size / 64, see Makefile).
#include "swn_gen.cuh"
#include "swn_launch.h"

#ifndef SWN_NCH
#error "compile with -DSWN_NCH=<1..4>"
#endif

namespace swn {

template <int NCH, int SEG, int BS>
static cudaError_t launch_inst(const GenParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = swn_generate_kernel<NCH, SEG, BS>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int nb = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, SWN_NTHREADS, smem);
    if (e != cudaSuccess) return e;
    if (nb < 1) return cudaErrorLaunchOutOfResources;
    void* args[] = {(void*)&p};
    return cudaLaunchCooperativeKernel((const void*)k, dim3(grid), dim3(SWN_NTHREADS), args, smem, st);
}

#define SWN_CASE(S, B) \
    if (seg == S && bs == B) return &launch_inst<SWN_NCH, S, B>;

#define SWN_CAT2(a, b) a##b
#define SWN_CAT(a, b) SWN_CAT2(a, b)

launch_fn SWN_CAT(find_kernel_nch, SWN_NCH)(int seg, int bs) {
    SWN_CASE(1, 1) SWN_CASE(1, 2) SWN_CASE(1, 4) SWN_CASE(1, 8)
    SWN_CASE(2, 1) SWN_CASE(2, 2) SWN_CASE(2, 4)
    SWN_CASE(3, 1)
    SWN_CASE(4, 1)
    SWN_CASE(5, 1) SWN_CASE(5, 2)
    SWN_CASE(10, 1)
    return nullptr;
}

}  // namespace swn
