// Platform glue: real CUDA (nvcc, sm_100a) in the product build; tests/emul/cuda_emul.h when the
// CPU-only test tier compiles the same kernel sources with -DQR_HOST_EMUL (never shipped).
#pragma once
#ifdef QR_HOST_EMUL
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#define QR_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// launch through cudaLaunchKernelEx: optional thread-block cluster of `cluster` CTAs along x (grid must be a
// multiple of it) and optional programmatic dependent launch (PDL): the CTAs of this grid may be scheduled while
// the previous kernel on the stream drains; the kernel itself orders its memory accesses with griddepcontrol.wait.
template <class... KArgs, class... Args>
static inline cudaError_t qr_launch_ex(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                       unsigned cluster, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    unsigned na = 0;
    if (cluster > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cluster;
        at[na].val.clusterDim.y = 1;
        at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#define QR_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cluster, ...) \
    qr_launch_ex(kernel, (grid), (block), (smem), (stream), (cluster), false, __VA_ARGS__)
#define QR_LAUNCH_EX(kernel, grid, block, smem, stream, cluster, pdl, ...) \
    qr_launch_ex(kernel, (grid), (block), (smem), (stream), (cluster), (pdl), __VA_ARGS__)
#define QR_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char qr_dyn_smem_[]; \
    type* name = reinterpret_cast<type*>(qr_dyn_smem_)
#endif

#include <cstdint>
#include <cstddef>

typedef unsigned long long u64;
typedef long long i64;

// complex128 helpers on double2 (x = re, y = im)
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
// Im(conj(l) * p)
__device__ __forceinline__ double im_conj_mul(double2 l, double2 p) { return l.x * p.y - l.y * p.x; }
// Re(conj(l) * p)
__device__ __forceinline__ double re_conj_mul(double2 l, double2 p) { return l.x * p.x + l.y * p.y; }

// CNOT-ladder index map (state.py:229-241): j' = j ^ ((j>>1)&M1) ^ ((j>>2)&M2)
__host__ __device__ __forceinline__ u64 ladder_map(u64 j, u64 m1, u64 m2) {
    return j ^ ((j >> 1) & m1) ^ ((j >> 2) & m2);
}
