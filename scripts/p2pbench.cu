// NVLink P2P microbenchmark (diagnostic, not part of the library): device 0 and device 1 each run a kernel that moves
// `bytes` between its own memory and the peer's, both at the same time (as the exchange passes of the swap engine do).
//   mode 0: pull, 16 B loads (what k_tile12_x does)      mode 1: pull, 32 B loads (ld.global.v4.f64)
//   mode 2: pull, 16 B ld.global.cg                      mode 3: pull, 16 B ld.global.nc (read-only path)
//   mode 4: push, 16 B stores to the peer                mode 5: push, 32 B stores
//   mode 6: cudaMemcpyPeerAsync (copy engines)
//   mode 7: pull with bulk asynchronous copies (cp.async.bulk global(peer) -> shared, mbarrier; shared -> local global), one
//           issuing thread per CTA, ring of 4 buffers; argv[3] = chunk size in KiB (4, 8, 16)
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

// mode 7: bulk-copy pull.  One thread per CTA drives a ring of NB shared-memory buffers.
template <int NB>
__global__ void __launch_bounds__(32) k_bulk(const char* __restrict__ src, char* __restrict__ dst, size_t bytes, int chunk) {
    extern __shared__ __align__(128) char ring[];
    __shared__ __align__(8) unsigned long long bar[NB];
    if (threadIdx.x != 0) return;
    const size_t nchunks = bytes / (size_t)chunk;
    for (int b = 0; b < NB; ++b) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(a));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    unsigned phase[NB];
    for (int b = 0; b < NB; ++b) phase[b] = 0;
    auto issue = [&](size_t c, int b) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]);
        const unsigned d = (unsigned)__cvta_generic_to_shared(ring + (size_t)b * chunk);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(src + c * (size_t)chunk),
                     "r"(chunk), "r"(a)
                     : "memory");
    };
    size_t next = blockIdx.x, cur = blockIdx.x;
    int nissued = 0;
    for (; nissued < NB && next < nchunks; ++nissued, next += gridDim.x) issue(next, nissued);
    int b = 0;
    for (; cur < nchunks; cur += gridDim.x) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(&bar[b]);
        unsigned ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(a), "r"(phase[b]) : "memory");
        phase[b] ^= 1;
        const unsigned sm = (unsigned)__cvta_generic_to_shared(ring + (size_t)b * chunk);
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + cur * (size_t)chunk), "r"(sm), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (next < nchunks) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the buffer has been read by the store
            issue(next, b);
            next += gridDim.x;
        }
        b = (b + 1) % NB;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
            else if (mode == 7) {
                const int chunk = (unroll == 4 || unroll == 8 || unroll == 16 || unroll == 32) ? unroll << 10 : 16 << 10;
                cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * chunk);
                k_bulk<4><<<ctas, 32, 4 * chunk, st[d]>>>((const char*)buf[1 - d], (char*)out[d], bytes, chunk);
            } else {
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
