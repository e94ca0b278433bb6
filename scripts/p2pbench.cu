// NVLink P2P microbenchmark (diagnostic, not part of the library): device 0 and device 1 each run a kernel that moves
// `bytes` between its own memory and the peer's, both at the same time (as the exchange passes of the swap engine do).
//   mode 0: pull, 16 B loads (what k_tile12_x does)      mode 1: pull, 32 B loads (ld.global.v4.f64)
//   mode 2: pull, 16 B ld.global.cg                      mode 3: pull, 16 B ld.global.nc (read-only path)
//   mode 4: push, 16 B stores to the peer                mode 5: push, 32 B stores
//   mode 6: cudaMemcpyPeerAsync (copy engines)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/bin/p2pbench scripts/p2pbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int MODE, int UNROLL>
__global__ void __launch_bounds__(512) k_move(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (MODE == 1 || MODE == 5) {   // 32-byte accesses
        const double4* s4 = reinterpret_cast<const double4*>(src);
        double4* d4 = reinterpret_cast<double4*>(dst);
        const size_t n4 = n / 2;
        for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
            double4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[u].x), "=d"(v[u].y), "=d"(v[u].z), "=d"(v[u].w) : "l"(s4 + i + u * stride));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(d4 + i + u * stride), "d"(v[u].x), "d"(v[u].y), "d"(v[u].z), "d"(v[u].w) : "memory");
        }
        return;
    }
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        double2 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const double2* p = src + i + u * stride;
            if (MODE == 2) asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(p));
            else if (MODE == 3) asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(v[u].x), "=d"(v[u].y) : "l"(p));
            else v[u] = *p;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) dst[i + u * stride] = v[u];
    }
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int ctas = argc > 2 ? atoi(argv[2]) : 148;
    const int unroll = argc > 3 ? atoi(argv[3]) : 8;
    const size_t bytes = (size_t)(argc > 4 ? atoi(argv[4]) : 4096) << 20;
    int nd = 0;
    cudaGetDeviceCount(&nd);
    if (nd < 2) { printf("needs 2 GPUs\n"); return 1; }
    double2 *buf[2], *out[2];
    cudaStream_t st[2];
    cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; ++d) {
        cudaSetDevice(d);
        cudaDeviceEnablePeerAccess(1 - d, 0);
        cudaMalloc(&buf[d], bytes);
        cudaMalloc(&out[d], bytes);
        cudaMemset(buf[d], 1, bytes);
        cudaStreamCreate(&st[d]);
        cudaEventCreate(&e0[d]);
        cudaEventCreate(&e1[d]);
    }
    const size_t n = bytes / sizeof(double2);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        for (int d = 0; d < 2; ++d) {
            cudaSetDevice(d);
            cudaEventRecord(e0[d], st[d]);
            const bool push = mode == 4 || mode == 5;
            const double2* s = push ? buf[d] : buf[1 - d];     // pull: read the peer, write locally; push: read locally, write the peer
            double2* t = push ? out[1 - d] : out[d];
            if (mode == 6) cudaMemcpyPeerAsync(out[d], d, buf[1 - d], 1 - d, bytes, st[d]);
            else {
#define LAUNCH(M, U) k_move<M, U><<<ctas, 512, 0, st[d]>>>(s, t, n)
#define BYU(M) do { if (unroll == 4) LAUNCH(M, 4); else if (unroll == 16) LAUNCH(M, 16); else LAUNCH(M, 8); } while (0)
                switch (mode) { case 0: BYU(0); break; case 1: BYU(1); break; case 2: BYU(2); break; case 3: BYU(3); break; case 4: BYU(4); break; default: BYU(5); break; }
            }
            cudaEventRecord(e1[d], st[d]);
        }
        float worst = 0.f;
        for (int d = 0; d < 2; ++d) {
            cudaSetDevice(d);
            cudaEventSynchronize(e1[d]);
            float ms;
            cudaEventElapsedTime(&ms, e0[d], e1[d]);
            worst = ms > worst ? ms : worst;
        }
        if (rep > 0 && worst < best) best = worst;
    }
    cudaError_t err = cudaGetLastError();
    printf("mode=%d ctas=%d unroll=%d MiB=%zu  ms=%.3f  GB/s per direction=%.0f %s\n", mode, ctas, unroll, bytes >> 20, best, bytes / best / 1e6,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    return 0;
}
