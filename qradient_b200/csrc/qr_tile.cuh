// Fused tile pass: the hot kernel of run_expec_val / grad_run.
//
// One launch sweeps the whole state vector once (forward: psi; backward: psi and the co-state
// lambda together) and applies, per HBM pass, every single-qubit rotation whose index bit lies
// in the tile -- up to 12 gates per 32 B/amplitude of traffic instead of one
// (mc_clean.py:38-41 forward loop, :68-77 backward loop; qaoa.py:49-53, :59-69).
//
// Tile geometry.  A tile is 2^k amplitudes (k <= 12, 64 KiB per vector).  Its local index bits
// [0,c) are global index bits [0,c) (>= 128 B contiguous runs) and local bits [c,k) are global
// bits [h, h+k-c).  The first pass of a layer uses c = k (a contiguous tile) and can fold the
// CNOT ladder in as a gather: the ladder's index map j' = j ^ ((j>>1)&M1) ^ ((j>>2)&M2) is
// GF(2)-linear and each new bit depends only on MORE significant bits (state.py:229-241), so a
// destination tile reads exactly one contiguous source tile, permuted inside; the gather is
// applied on the fly to the global load addresses (coalescing is preserved because a warp's 32
// lanes map onto one 512 B block).  Ladder passes are out of place (ping-pong buffers).
//
// Thread program.  2^(k-4) threads per tile; each thread holds 16 amplitudes per vector in
// registers (4 index bits), applies the gates of those bits, and exchanges through XOR-swizzled
// shared memory to bring the next 4 bits into registers: 3 rounds cover 12 bits with 2 shared
// memory round trips.  Round 0 loads straight from global memory and the last round stores
// straight to global memory.  In the backward pass the per-parameter inner products
// Im<lambda|P_q|psi> (the closed form of mc_clean.py:73-75, since dR(-t)R(t) = -iP/2) are
// accumulated in fp64 registers across all tiles a CTA processes (persistent grid-stride
// loop) and reduced once per CTA; a second tiny kernel sums the per-CTA partials in a fixed
// order, so results are run-to-run deterministic.
//
// Roofline: HBM.  Algorithmic bytes per launch = NV * 32 B * 2^n.  No tensor cores: nothing
// here is a dense contraction.
#pragma once
#include "qr_platform.cuh"

#define QR_R 4                       // index bits held in registers per round
#define QR_RA (1 << QR_R)            // amplitudes per thread per vector
#define QR_MAXROUNDS 3
#define QR_SLOTS (QR_MAXROUNDS * QR_R + 1)   // gradient accumulators per thread (+1: diagonal generator)
#define QR_MAX_TILE_BITS 12

struct GateP {
    double c, s;    // cos(theta/2), sin(theta/2) (sign already folded in for un-rotation)
    int axis;       // 0,1,2 = X,Y,Z ; -1 = no gate on this bit in this round
    int pad;
};

struct TilePass {
    int k, c, h;                 // tile geometry (see header comment)
    int nrounds;
    int g[QR_MAXROUNDS];         // first local bit of the register group of each round
    int ladder;                  // 1: gather through the ladder map on load
    u64 M1, M2;
    int tiles_log2;              // log2(tiles per state) = n - k
    i64 num_tiles;               // batch * tiles per state
    i64 state_stride;            // amplitudes between consecutive states of a batch
    const double2* src0;         // psi in
    const double2* src1;         // lambda in (NV == 2)
    double2* dst0;
    double2* dst1;
    const GateP* gates;          // [batch][nrounds * QR_R]
    int gate_stride;
    const double* ham;           // diagonal Hamiltonian table (QAOA) or null
    int pre_phase, post_phase;   // multiply by exp(-i angle H) after load / before store
    double angle_pre, angle_post;
    double* partials;            // [units][QR_SLOTS]; unit = CTA, or tile when flush_per_tile
    int flush_per_tile;
    int prefetch;
};

__device__ __forceinline__ int qr_swz(int l) { return l ^ ((l >> QR_R) & 7); }

template <int NV, int BIT>
__device__ __forceinline__ void qr_gate_on_bit(double2 (&a)[NV][QR_RA], const GateP gp, double& acc) {
    const double c = gp.c, s = gp.s;
    if (gp.axis == 0) {
#pragma unroll
        for (int r = 0; r < QR_RA; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
            if (NV == 2) acc += im_conj_mul(a[NV - 1][r], a[0][r1]) + im_conj_mul(a[NV - 1][r1], a[0][r]);
#pragma unroll
            for (int v = 0; v < NV; ++v) {   // Rx: a' = c a - i s b ; b' = -i s a + c b   (state.py:90-92)
                const double2 x = a[v][r], y = a[v][r1];
                a[v][r] = make_double2(c * x.x + s * y.y, c * x.y - s * y.x);
                a[v][r1] = make_double2(c * y.x + s * x.y, c * y.y - s * x.x);
            }
        }
    } else if (gp.axis == 1) {
#pragma unroll
        for (int r = 0; r < QR_RA; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
            if (NV == 2) acc += re_conj_mul(a[NV - 1][r1], a[0][r]) - re_conj_mul(a[NV - 1][r], a[0][r1]);
#pragma unroll
            for (int v = 0; v < NV; ++v) {   // Ry: a' = c a - s b ; b' = s a + c b   (state.py:142-144)
                const double2 x = a[v][r], y = a[v][r1];
                a[v][r] = make_double2(c * x.x - s * y.x, c * x.y - s * y.y);
                a[v][r1] = make_double2(s * x.x + c * y.x, s * x.y + c * y.y);
            }
        }
    } else if (gp.axis == 2) {
#pragma unroll
        for (int r = 0; r < QR_RA; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
            if (NV == 2) acc += im_conj_mul(a[NV - 1][r], a[0][r]) - im_conj_mul(a[NV - 1][r1], a[0][r1]);
#pragma unroll
            for (int v = 0; v < NV; ++v) {   // Rz: a' = (c - i s) a ; b' = (c + i s) b   (state.py:168-170)
                const double2 x = a[v][r], y = a[v][r1];
                a[v][r] = make_double2(c * x.x + s * x.y, c * x.y - s * x.x);
                a[v][r1] = make_double2(c * y.x - s * y.y, c * y.y + s * y.x);
            }
        }
    }
}

template <int NV>
__device__ __forceinline__ void qr_round_compute(double2 (&a)[NV][QR_RA], const GateP* __restrict__ gt, double* acc) {
    qr_gate_on_bit<NV, 0>(a, gt[0], acc[0]);
    qr_gate_on_bit<NV, 1>(a, gt[1], acc[1]);
    qr_gate_on_bit<NV, 2>(a, gt[2], acc[2]);
    qr_gate_on_bit<NV, 3>(a, gt[3], acc[3]);
}

template <int NV, int NR>
__global__ void __launch_bounds__(256, (NV == 1 ? 2 : 1)) k_tile_pass(const TilePass p) {
    QR_DYN_SMEM(double2, smem);
    const int tid = threadIdx.x;
    const int T = 1 << p.k;
    double acc[QR_SLOTS];
#pragma unroll
    for (int i = 0; i < QR_SLOTS; ++i) acc[i] = 0.0;
    int tb[NR];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        const int g = p.g[r];
        tb[r] = (tid & ((1 << g) - 1)) | ((tid >> g) << (g + QR_R));
    }
    const int c = p.c, h = p.h;
    const int lomask = (1 << c) - 1;
    const int nlo = h - c;
    const u64 tmask = ((u64)1 << p.tiles_log2) - 1;

    for (i64 tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const i64 b = tile >> p.tiles_log2;
        const u64 t = (u64)tile & tmask;
        const u64 tbase = ((t & (((u64)1 << nlo) - 1)) << c) | ((t >> nlo) << (h + p.k - c));
        const GateP* __restrict__ gt = p.gates + b * p.gate_stride;
        const double2* __restrict__ s0 = p.src0 + b * p.state_stride;
        const double2* __restrict__ s1 = (NV == 2) ? p.src1 + b * p.state_stride : nullptr;
        double2* __restrict__ d0 = p.dst0 + b * p.state_stride;
        double2* __restrict__ d1 = (NV == 2) ? p.dst1 + b * p.state_stride : nullptr;

        double2 a[NV][QR_RA];
        // ---- round 0: global -> registers (ladder gather and QAOA phase folded in) ----
        {
            const int g = p.g[0];
#pragma unroll
            for (int r = 0; r < QR_RA; ++r) {
                const int l = tb[0] | (r << g);
                const u64 d = tbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
                const u64 s = p.ladder ? ladder_map(d, p.M1, p.M2) : d;
                a[0][r] = s0[s];
                if (NV == 2) a[NV - 1][r] = s1[s];
            }
            if (p.pre_phase) {
#pragma unroll
                for (int r = 0; r < QR_RA; ++r) {
                    const int l = tb[0] | (r << g);
                    const u64 d = tbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
                    double sn, cs;
                    sincos(p.angle_pre * p.ham[d], &sn, &cs);
                    const double2 ph = make_double2(cs, -sn);
#pragma unroll
                    for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
                }
            }
        }
#ifndef QR_HOST_EMUL
        if (p.prefetch) {   // pull the next tile of this CTA into L2 while this one is computed
            const i64 nt = tile + gridDim.x;
            if (nt < p.num_tiles) {
                const i64 nb = nt >> p.tiles_log2;
                const u64 t2 = (u64)nt & tmask;
                const u64 nbase = ((t2 & (((u64)1 << nlo) - 1)) << c) | ((t2 >> nlo) << (h + p.k - c));
                for (int line = tid; line < (T >> 3); line += blockDim.x) {
                    const int l = line << 3;
                    const u64 d = nbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
                    const u64 s = p.ladder ? ladder_map(d, p.M1, p.M2) : d;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src0 + nb * p.state_stride + s));
                    if (NV == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src1 + nb * p.state_stride + s));
                }
            }
        }
#endif
        qr_round_compute<NV>(a, gt, acc);
        // ---- rounds 1..NR-1: exchange through swizzled shared memory ----
#pragma unroll
        for (int rd = 1; rd < NR; ++rd) {
            const int gp = p.g[rd - 1], gn = p.g[rd];
#pragma unroll
            for (int r = 0; r < QR_RA; ++r) {
                const int l = qr_swz(tb[rd - 1] | (r << gp));
#pragma unroll
                for (int v = 0; v < NV; ++v) smem[v * T + l] = a[v][r];
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < QR_RA; ++r) {
                const int l = qr_swz(tb[rd] | (r << gn));
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = smem[v * T + l];
            }
            if (rd == NR - 1) __syncthreads();   // smem is free for the next tile's first exchange
            qr_round_compute<NV>(a, gt + rd * QR_R, acc + rd * QR_R);
        }
        // ---- registers -> global (QAOA: diagonal-generator inner product and un-phase) ----
        {
            const int g = p.g[NR - 1];
#pragma unroll
            for (int r = 0; r < QR_RA; ++r) {
                const int l = tb[NR - 1] | (r << g);
                const u64 d = tbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
                if (p.post_phase) {
                    const double hv = p.ham[d];
                    if (NV == 2) acc[QR_SLOTS - 1] += hv * im_conj_mul(a[NV - 1][r], a[0][r]);
                    double sn, cs;
                    sincos(p.angle_post * hv, &sn, &cs);
                    const double2 ph = make_double2(cs, -sn);
#pragma unroll
                    for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
                }
                d0[d] = a[0][r];
                if (NV == 2) d1[d] = a[NV - 1][r];
            }
        }
        if (NV == 2 && p.flush_per_tile) {
#pragma unroll
            for (int i = 0; i < QR_SLOTS; ++i) {
                const double s = block_reduce_sum(acc[i]);
                if (tid == 0) p.partials[(u64)tile * QR_SLOTS + i] = s;
                acc[i] = 0.0;
            }
        }
    }
    if (NV == 2 && !p.flush_per_tile) {
#pragma unroll
        for (int i = 0; i < QR_SLOTS; ++i) {
            const double s = block_reduce_sum(acc[i]);
            if (tid == 0) p.partials[(u64)blockIdx.x * QR_SLOTS + i] = s;
        }
    }
}
