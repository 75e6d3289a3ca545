// Instantiations of the seg > 1 cluster ("core + helpers") latency kernel (swn_coreseg.cuh) for one hidden-size class.
#include <stdio.h>
#include <stdlib.h>
#include "swn_coreseg.cuh"
#include "swn_launch.h"

#ifndef SWN_NCH
#error "compile with -DSWN_NCH=<1..3>"
#endif

namespace swn {

template <int NCH, int SEG>
static cudaError_t coreseg_prepare(size_t smem, int* max_clusters) {
    auto k = swn_coreseg_kernel<NCH, SEG>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CORE_CL * 8);
    cfg.blockDim = dim3(SWN_NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CORE_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaOccupancyMaxActiveClusters(max_clusters, (const void*)k, &cfg);
}

// (same launch rules as swn_core_inst.cu: cooperative, so that all clusters are co-resident; resource-class errors go
// back to the caller, which falls back to the layer-group kernel)
template <int NCH, int SEG>
static cudaError_t coreseg_launch(const CoreParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = swn_coreseg_kernel<NCH, SEG>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(SWN_NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, p);
    if (e == cudaSuccess) return e;
    (void)cudaGetLastError();
    if (e != cudaErrorInvalidValue && e != cudaErrorNotSupported && e != cudaErrorInvalidConfiguration) return e;
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, k, p);
}

#define SWN_CAT2(a, b) a##b
#define SWN_CAT(a, b) SWN_CAT2(a, b)

core_fns SWN_CAT(find_coreseg_kernel_nch, SWN_NCH)(int seg) {
    core_fns f = {nullptr, nullptr};
    switch (seg) {
        case 2: f.prepare = &coreseg_prepare<SWN_NCH, 2>; f.launch = &coreseg_launch<SWN_NCH, 2>; break;
        case 3: f.prepare = &coreseg_prepare<SWN_NCH, 3>; f.launch = &coreseg_launch<SWN_NCH, 3>; break;
        case 4: f.prepare = &coreseg_prepare<SWN_NCH, 4>; f.launch = &coreseg_launch<SWN_NCH, 4>; break;
        case 5: f.prepare = &coreseg_prepare<SWN_NCH, 5>; f.launch = &coreseg_launch<SWN_NCH, 5>; break;
        default: break;
    }
    return f;
}

}  // namespace swn
