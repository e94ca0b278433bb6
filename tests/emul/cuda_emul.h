// Host emulation of the small CUDA subset the qradient_b200 kernels use.
//
// TEST INFRASTRUCTURE ONLY.  The product library (libqradient_b200.so) is compiled by nvcc
// for sm_100a and never sees this header.  `tests/emul/build_emul.py` compiles the SAME
// kernel sources with g++ -DQR_HOST_EMUL against this shim into tests/emul/libqr_emul.so so
// that the CPU-only test tier (`pytest -m "not gpu"`) can execute every kernel's thread
// program (index arithmetic, shared-memory exchanges, barriers, reductions, host planner)
// without a GPU.  Nothing under qradient_b200/ loads that library.
//
// Execution model: blocks run one after another; the threads of a block are ucontext fibers
// scheduled round-robin, __syncthreads() is a cooperative barrier, warp shuffles are
// emulated with a per-block exchange buffer and two barriers (valid because every kernel
// here calls them in block-uniform control flow).
#pragma once
#include <ucontext.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __grid_constant__

struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint3_ { unsigned x, y, z; };

namespace emul {
extern uint3_ g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char* g_smem;
void sync();
double shfl_xor(double v, int lane_mask);
long long shfl_xor_ll(long long v, int lane_mask);
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
}  // namespace emul

#define threadIdx (emul::g_threadIdx)
#define blockIdx (emul::g_blockIdx)
#define blockDim (emul::g_blockDim)
#define gridDim (emul::g_gridDim)
static inline void __syncthreads() { emul::sync(); }
static inline void __syncwarp() { emul::sync(); }   // fibers of a warp run one after another: a warp barrier must yield too
static inline double __shfl_xor_sync(unsigned, double v, int m) { return emul::shfl_xor(v, m); }
static inline double __shfl_down_sync(unsigned, double v, int d) {
    // only used in full-warp tree reductions where lane 0's result matters: xor gives the same sum
    return emul::shfl_xor(v, d);
}
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double atomicAdd(double* p, double v) { double o = *p; *p = o + v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { auto o = *p; *p = o + v; return o; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline void __threadfence() {}
#include <sched.h>
static inline void emul_yield_cpu() { sched_yield(); }
static inline int __double2int_rn(double x) { return (int)std::nearbyint(x); }

// ---- runtime subset -------------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct emul_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; size_t sharedMemPerBlockOptin; int major, minor; char name[64]; int l2CacheSize; };
static inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emulated CUDA error" : "no error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p)); p->multiProcessorCount = 4; p->totalGlobalMem = (size_t)8 << 30;
    p->sharedMemPerBlockOptin = 227 * 1024; p->major = 10; p->minor = 0; strcpy(p->name, "host-emulation");
    p->l2CacheSize = 1 << 20; return cudaSuccess; }
// "device" allocations are POSIX shared memory so that the IPC entry points work across the
// processes of a gloo world_size-2 test (emulates cudaIpcGetMemHandle / cudaIpcOpenMemHandle)
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
cudaError_t emul_shm_malloc(void** p, size_t n);
cudaError_t emul_shm_free(void* p);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
static inline cudaError_t cudaMalloc(void** p, size_t n) { return emul_shm_malloc(p, n); }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return emul_shm_malloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { return emul_shm_free(p); }
static inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMallocHost((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emul_event{0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
double emul_now_ms();
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = 0) { e->t = emul_now_ms(); return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
static inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = *t = (size_t)8 << 30; return cudaSuccess; }

#define QR_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emul::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#define QR_LAUNCH_CLUSTER(kernel, grid, block, smem, stream, cluster, ...) \
    (emul::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); }), cudaSuccess)   /* clusters: blocks run one after another */
#define QR_LAUNCH_EX(kernel, grid, block, smem, stream, cluster, pdl, ...) \
    (emul::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); }), cudaSuccess)   /* PDL: kernels run one after another */
#define QR_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emul::g_smem)
