// Memory-system microbenchmark for the strided tile passes (diagnostic, not part of the library).
//
// Each CTA streams "tiles" of 4096 amplitudes (16 B each) of NV vectors: 512 rows of 128 B whose row index bits
// (local bits 3..11) are mapped to an arbitrary list of global index bits; the data is loaded into registers and
// stored back in place (no arithmetic), with the same thread <-> address mapping as k_tile12's load side.
// Usage: membench n nv bit3 bit4 ... bit11   (global bit of each local row bit; local bits 0-2 = global bits 0-2)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

typedef unsigned long long u64;

struct Map { u64 bitpos[9]; int n; };

__device__ __forceinline__ u64 local_to_global(const Map& m, int l) {
    u64 d = l & 7;
#pragma unroll
    for (int j = 0; j < 9; ++j) d |= (u64)((l >> (3 + j)) & 1) << m.bitpos[j];
    return d;
}

// loads with an L2 prefetch-size hint: the L2 fetches the whole aligned 256 B (128 B) chunk from DRAM on a miss
__device__ __forceinline__ double2 ld_l2_256(const double2* p) {
    double2 v;
    asm volatile("ld.global.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_l2_128(const double2* p) {
    double2 v;
    asm volatile("ld.global.L2::128B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// MODE 0 plain, 1 L2 prefetch of the next tile, 2 cluster barrier per tile, 3 one CTA takes adjacent tile pairs,
// 4 loads with L2::256B, 5 = 4 + adjacent tile pairs per CTA, 6 loads with L2::128B, 7 = 4 + streaming (evict-first) stores
template <int NV, int MODE>
__global__ void __launch_bounds__(512, 1) k_touch(double2* v0, double2* v1, double2* w0, double2* w1, Map m, u64 tile_mask_bits, long long num_tiles, const u64* tile_base) {
    const int tid = threadIdx.x;
    const u64 toff = local_to_global(m, tid);
    u64 roff[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) roff[r] = local_to_global(m, r << 9);
    constexpr bool PAIRS = MODE == 3 || MODE == 5;
    for (long long t = (PAIRS ? (long long)blockIdx.x * 2 : blockIdx.x); t < num_tiles; t += (PAIRS ? ((t & 1) ? 2 * (long long)gridDim.x - 1 : 1) : gridDim.x)) {
        const u64 tb = tile_base[t] | toff;
        if (MODE == 2) asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
        double2 a[NV][8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (MODE == 4 || MODE == 5 || MODE == 7) {
                a[0][r] = ld_l2_256(v0 + (tb | roff[r]));
                if (NV == 2) a[NV - 1][r] = ld_l2_256(v1 + (tb | roff[r]));
            } else if (MODE == 6) {
                a[0][r] = ld_l2_128(v0 + (tb | roff[r]));
                if (NV == 2) a[NV - 1][r] = ld_l2_128(v1 + (tb | roff[r]));
            } else {
                a[0][r] = v0[tb | roff[r]];
                if (NV == 2) a[NV - 1][r] = v1[tb | roff[r]];
            }
        }
        if (MODE == 1) {   // prefetch next tile to L2
            const long long nt = t + gridDim.x;
            if (nt < num_tiles) {
                const u64 d = tile_base[nt] | local_to_global(m, tid << 3);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(v0 + d));
                if (NV == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(v1 + d));
            }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a[0][r].x += 1.0;
            if (MODE == 7) {
                __stcs(w0 + (tb | roff[r]), a[0][r]);
                if (NV == 2) { a[NV - 1][r].y += 1.0; __stcs(w1 + (tb | roff[r]), a[NV - 1][r]); }
            } else {
                w0[tb | roff[r]] = a[0][r];
                if (NV == 2) { a[NV - 1][r].y += 1.0; w1[tb | roff[r]] = a[NV - 1][r]; }
            }
        }
    }
}

int main(int argc, char** argv) {
    if (argc < 12) { fprintf(stderr, "usage: membench n nv b3..b11 [prefetch]\n"); return 1; }
    const int n = atoi(argv[1]), nv = atoi(argv[2]);
    Map m; m.n = n;
    u64 used = 7;
    for (int j = 0; j < 9; ++j) { m.bitpos[j] = atoi(argv[3 + j]); used |= (u64)1 << m.bitpos[j]; }
    const int mode = argc > 12 ? atoi(argv[12]) : 0;   // 0 plain, 1 L2 prefetch, 2 cluster barrier per tile, 3 one CTA takes adjacent tile pairs
    const int cs = argc > 13 ? atoi(argv[13]) : 1;     // cluster size
    const int grid = argc > 14 ? atoi(argv[14]) : 148;
    const int oop = argc > 15 ? atoi(argv[15]) : 0;    // 1: out of place (separate destination buffers)
    const int l2gran = argc > 16 ? atoi(argv[16]) : 0;  // > 0: cudaLimitMaxL2FetchGranularity
    if (l2gran > 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)l2gran);
    size_t gran = 0; cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
    const u64 N = (u64)1 << n;
    const long long num_tiles = (long long)(N >> 12);
    // tile t -> base index: deposit t's bits into the unused global bits, ascending
    u64* h_base = (u64*)malloc(num_tiles * sizeof(u64));
    for (long long t = 0; t < num_tiles; ++t) {
        u64 d = 0; int tb = 0;
        for (int g = 0; g < n; ++g) if (!((used >> g) & 1)) { d |= (u64)((t >> tb) & 1) << g; ++tb; }
        h_base[t] = d;
    }
    u64* d_base; cudaMalloc(&d_base, num_tiles * sizeof(u64));
    cudaMemcpy(d_base, h_base, num_tiles * sizeof(u64), cudaMemcpyHostToDevice);
    double2 *v0, *v1 = nullptr;
    if (cudaMalloc(&v0, N * 16) != cudaSuccess) { fprintf(stderr, "alloc failed\n"); return 1; }
    if (nv == 2 && cudaMalloc(&v1, N * 16) != cudaSuccess) { fprintf(stderr, "alloc failed\n"); return 1; }
    double2 *w0 = nullptr, *w1 = nullptr;
    if (oop) { if (cudaMalloc(&w0, N * 16) != cudaSuccess) return 1; if (nv == 2 && cudaMalloc(&w1, N * 16) != cudaSuccess) return 1; }
    cudaMemset(v0, 0, N * 16);
    if (v1) cudaMemset(v1, 0, N * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        void (*fn)(double2*, double2*, double2*, double2*, Map, u64, long long, const u64*) = nullptr;
        typedef void (*fn_t)(double2*, double2*, double2*, double2*, Map, u64, long long, const u64*);
        static const fn_t f1[8] = {k_touch<1, 0>, k_touch<1, 1>, k_touch<1, 2>, k_touch<1, 3>, k_touch<1, 4>, k_touch<1, 5>, k_touch<1, 6>, k_touch<1, 7>};
        static const fn_t f2[8] = {k_touch<2, 0>, k_touch<2, 1>, k_touch<2, 2>, k_touch<2, 3>, k_touch<2, 4>, k_touch<2, 5>, k_touch<2, 6>, k_touch<2, 7>};
        fn = (nv == 1 ? f1 : f2)[mode & 7];
        cudaLaunchKernelEx(&cfg, fn, v0, v1, oop ? w0 : v0, oop ? w1 : v1, m, (u64)0, num_tiles, (const u64*)d_base);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("n=%d nv=%d mode=%d cs=%d grid=%d oop=%d l2gran=%d bits=", n, nv, mode, cs, grid, oop, (int)gran);
    for (int j = 0; j < 9; ++j) printf("%d,", (int)m.bitpos[j]);
    printf(" ms=%.3f GB/s=%.0f %s\n", best, nv * 32.0 * N / best / 1e6, err == cudaSuccess ? "" : cudaGetErrorString(err));
    return 0;
}
