// Instantiations of the cluster ("core + helpers") latency kernel for one hidden-size class.
#include <stdio.h>
#include <stdlib.h>
#include "swn_core.cuh"
#include "swn_core2.cuh"
#include "swn_launch.h"

#ifndef SWN_NCH
#error "compile with -DSWN_NCH=<1..4>"
#endif

namespace swn {

template <int NCH, int BS, bool MULTI>
static cudaError_t core_prepare(size_t smem, int* max_clusters) {
    auto k = swn_core_kernel2<NCH, BS, MULTI>;
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

template <int NCH, int BS, bool MULTI>
static cudaError_t core_launch(const CoreParams& p, int grid, size_t smem, cudaStream_t st) {
    auto k = swn_core_kernel2<NCH, BS, MULTI>;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(SWN_NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;          // all clusters co-resident: they spin on each other
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, p);
    static const bool dbg = getenv("SWN_DEBUG_LAUNCH") != nullptr;
    if (dbg) fprintf(stderr, "[swn] cluster kernel: cooperative launch of %d CTAs -> %s\n", grid, cudaGetErrorName(e));
    if (e == cudaSuccess) return e;
    (void)cudaGetLastError();
    // The clusters spin on each other (DSMEM, mbarriers, LL words), so they must all be co-resident: a resource-class
    // error (too large / out of resources - SMs held by other streams, MPS clients or processes) is returned to the
    // caller, which falls back to the cooperatively launched layer-group kernel.  Only when the driver rejects the
    // ATTRIBUTE COMBINATION cooperative + compile-time cluster dims (an argument-class error) is the launch repeated
    // without the cooperative attribute; the caller has checked the grid against cudaOccupancyMaxActiveClusters, and
    // a cluster that still fails to become resident trips the in-kernel watchdog instead of hanging.
    if (e != cudaErrorInvalidValue && e != cudaErrorNotSupported && e != cudaErrorInvalidConfiguration) return e;
    cfg.numAttrs = 0;
    return cudaLaunchKernelEx(&cfg, k, p);
}

#define SWN_CAT2(a, b) a##b
#define SWN_CAT(a, b) SWN_CAT2(a, b)

// multi = several core clusters side by side (one or two utterances each) sharing the helper CTAs: instantiated where the
// helpers' shared memory can hold more than one utterance group's rings (hid_chn <= 192)
core_fns SWN_CAT(find_core_kernel_nch, SWN_NCH)(int bs, bool multi) {
    core_fns f = {nullptr, nullptr};
    if (multi) {
#if SWN_NCH <= 3
        switch (bs) {
            case 1: f.prepare = &core_prepare<SWN_NCH, 1, true>; f.launch = &core_launch<SWN_NCH, 1, true>; break;
            case 2: f.prepare = &core_prepare<SWN_NCH, 2, true>; f.launch = &core_launch<SWN_NCH, 2, true>; break;
            default: break;
        }
#endif
        return f;
    }
    switch (bs) {
        case 1: f.prepare = &core_prepare<SWN_NCH, 1, false>; f.launch = &core_launch<SWN_NCH, 1, false>; break;
        case 2: f.prepare = &core_prepare<SWN_NCH, 2, false>; f.launch = &core_launch<SWN_NCH, 2, false>; break;
        case 4: f.prepare = &core_prepare<SWN_NCH, 4, false>; f.launch = &core_launch<SWN_NCH, 4, false>; break;
        default: break;
    }
    return f;
}

}  // namespace swn
