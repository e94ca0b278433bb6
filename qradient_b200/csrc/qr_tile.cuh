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

#define QR_MAXROUNDS 4
#define QR_GATE_SLOTS 12                      // gate bits per pass (= QR_MAX_TILE_BITS)
#define QR_SLOTS (QR_GATE_SLOTS + 1)          // gradient accumulators per thread (+1: diagonal generator)
#define QR_MAX_TILE_BITS 12
#define QR_LUT_MAX 256                        // phase look-up table entries (integer-valued Hamiltonians)

struct GateP {
    double c, s;    // cos(theta/2), sin(theta/2) (sign already folded in for un-rotation)
    int axis;       // 0,1,2 = X,Y,Z ; -1 = no gate on this bit in this round
    int pad;
};

struct TilePass {
    int k, c, h;                 // tile geometry (see header comment)
    int m1, h2;                  // k_tile12 only: second segment of gate bits (qr_tile12.cuh, Geo12)
    int nrounds;
    int g[QR_MAXROUNDS];         // first local bit of the register group of each round
    int ladder;                  // 1: gather through the ladder map on load
    u64 M1, M2;
    u64 src_xor;                 // sharded states: carry of the rank bits into the local source index
    int tiles_log2;              // log2(tiles per state) = n - k
    i64 num_tiles;               // batch * tiles per state
    i64 state_stride;            // amplitudes between consecutive states of a batch
    const double2* src0;         // psi in
    const double2* src1;         // lambda in (NV == 2)
    double2* dst0;
    double2* dst1;
    const GateP* gates;          // [batch][nrounds * R]
    int gate_stride;
    const double* ham;           // diagonal Hamiltonian table (QAOA) or null
    const short* hidx;           // integer-valued H: H[j] = hmin + hidx[j] (2 B/amp instead of 8) or null
    const double2* lut;          // exp(-i angle (hmin + v)) for v in [0, lut_size): replaces sincos per amplitude
    int lut_size;
    double hmin;
    int pre_phase, post_phase;   // multiply by exp(-i angle H) after load / before store
    double angle_pre, angle_post;
    double* partials;            // [units][QR_SLOTS]; unit = CTA, or tile when flush_per_tile
    double* final_out;           // if set: the last CTA writes the QR_SLOTS sums over all CTAs here
    unsigned* done_counter;      // zero-initialised arrival counter for final_out
    int flush_per_tile;
    int prefetch;
    int hole;                    // k_tile12, sliced passes: index bits [9, 9+hole) are fixed by tile_or instead of enumerated
    u64 tile_or;
};

// Deferred second stage of the gradient reductions (QR_OPT_DEFER_REDUCE): every backward pass of a gradient leaves its
// per-CTA partials in its own slice of the scratch buffer, and ONE launch after the sweep adds them -- block g handles
// (layer, pass) g, warp w slot w: lane l adds the units l, l+32, ... and a fixed shuffle tree finishes.  Same order of
// additions as the fused last-CTA reduction of k_tile12, without its fence + atomic + tail on every pass.
__global__ void k_reduce_slots_strided(const double* __restrict__ partial, int units, u64 group_stride, double* __restrict__ out,
                                       int out_stride);
template <int R>
__device__ __forceinline__ int qr_swz(int l) { return l ^ ((l >> R) & 7); }

// Block-wide sums of QR_SLOTS accumulators with ONE barrier pair: warp shuffles, per-warp rows in
// shared memory, then thread i < QR_SLOTS adds the rows in warp order and writes out[i].
__device__ __forceinline__ void qr_block_reduce_slots(const double (&acc)[QR_SLOTS], double* out) {
    __shared__ double rows[32][QR_SLOTS + 1];
    const int tid = threadIdx.x;
    if (blockDim.x >= 32) {
        const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
        for (int i = 0; i < QR_SLOTS; ++i) {
            double v = acc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) rows[w][i] = v;
        }
        __syncthreads();
        if (tid < QR_SLOTS) {
            double v = 0.0;
            for (int r = 0; r < nw; ++r) v += rows[r][tid];
            out[tid] = v;
        }
        __syncthreads();
    } else {
#pragma unroll
        for (int i = 0; i < QR_SLOTS; ++i) rows[tid][i] = acc[i];
        __syncthreads();
        if (tid == 0)
            for (int i = 0; i < QR_SLOTS; ++i) {
                double v = 0.0;
                for (unsigned r = 0; r < blockDim.x; ++r) v += rows[r][i];
                out[i] = v;
            }
        __syncthreads();
    }
}

// one single-qubit rotation (and, for NV == 2, its gradient inner product) on register bit BIT
template <int NV, int NA, int BIT>
__device__ __forceinline__ void qr_gate_on_bit(double2 (&a)[NV][NA], const GateP gp, double& acc) {
    const double c = gp.c, s = gp.s;
    if (gp.axis == 0) {
#pragma unroll
        for (int r = 0; r < NA; ++r) {
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
        for (int r = 0; r < NA; ++r) {
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
        for (int r = 0; r < NA; ++r) {
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

template <int NV, int R>
__device__ __forceinline__ void qr_round_compute(double2 (&a)[NV][1 << R], const GateP* gt, double* acc) {
    qr_gate_on_bit<NV, (1 << R), 0>(a, gt[0], acc[0]);
    qr_gate_on_bit<NV, (1 << R), 1>(a, gt[1], acc[1]);
    qr_gate_on_bit<NV, (1 << R), 2>(a, gt[2], acc[2]);
    if (R > 3) qr_gate_on_bit<NV, (1 << R), (R > 3 ? 3 : 0)>(a, gt[R > 3 ? 3 : 0], acc[R > 3 ? 3 : 0]);
}

// Generic tile pass for tiles of fewer than 11 bits (registers of fewer than 11 qubits) and explicit QR_OPT_TILE_BITS:
// R = index bits per round held in registers (2^R amplitudes per thread per vector); threads per tile = 2^(k-R).
// The round loop is a run-time loop so the code stays inside the instruction cache; per-round accumulators are
// folded into a small per-thread array.  Registers of >= 11 qubits run k_tile12 (qr_tile12.cuh).
template <int NV, int R>
__global__ void __launch_bounds__(1 << (QR_MAX_TILE_BITS - R), (NV == 1 ? 2 : 1)) k_tile_pass(const TilePass p) {
    constexpr int RA = 1 << R;
    QR_DYN_SMEM(double2, smem);
    __shared__ GateP sgt[QR_GATE_SLOTS];    // this pass's gate table (a dependent global load per gate
                                            // would put an L2 round trip on every tile's critical path)
    __shared__ double2 lut_sm[QR_LUT_MAX];  // QAOA phase factors by integer Hamiltonian value
    const int tid = threadIdx.x;
    const int T = 1 << p.k;
    double2* exch = smem;
    double acc_all[QR_SLOTS];
#pragma unroll
    for (int i = 0; i < QR_SLOTS; ++i) acc_all[i] = 0.0;
    const int c = p.c, h = p.h;
    const int lomask = (1 << c) - 1;
    const int nlo = h - c;
    const u64 tmask = ((u64)1 << p.tiles_log2) - 1;
    const int nrounds = p.nrounds;
    const int g_first = p.g[0], g_last = p.g[nrounds - 1];
    const int tb_first = (tid & ((1 << g_first) - 1)) | ((tid >> g_first) << (g_first + R));
    const int tb_last = (tid & ((1 << g_last) - 1)) | ((tid >> g_last) << (g_last + R));
    i64 cur_b = -1;
    const bool use_lut = p.hidx != nullptr && (p.pre_phase || p.post_phase);
    if (use_lut) {
        for (int i = threadIdx.x; i < p.lut_size; i += blockDim.x) lut_sm[i] = p.lut[i];
        __syncthreads();
    }

    for (i64 tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const i64 b = tile >> p.tiles_log2;
        const u64 t = (u64)tile & tmask;
        const u64 tbase = ((t & (((u64)1 << nlo) - 1)) << c) | ((t >> nlo) << (h + p.k - c));
        if (b != cur_b) {   // block-uniform: (re)load the gate table of this batch element
            __syncthreads();
            for (int i = tid; i < QR_GATE_SLOTS; i += blockDim.x) sgt[i] = p.gates[b * p.gate_stride + i];
            __syncthreads();
            cur_b = b;
        }
        const GateP* gt = sgt;
        const double2* __restrict__ s0 = p.src0 + b * p.state_stride;
        const double2* __restrict__ s1 = (NV == 2) ? p.src1 + b * p.state_stride : nullptr;
        double2* __restrict__ d0 = p.dst0 + b * p.state_stride;
        double2* __restrict__ d1 = (NV == 2) ? p.dst1 + b * p.state_stride : nullptr;

        double2 a[NV][RA];
        // ---- global -> registers (ladder gather folded into the load addresses) ----
#pragma unroll
        for (int r = 0; r < RA; ++r) {
            const int l = tb_first | (r << g_first);
            const u64 d = tbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
            const u64 s = p.ladder ? (ladder_map(d, p.M1, p.M2) ^ p.src_xor) : d;
            a[0][r] = s0[s];
            if (NV == 2) a[NV - 1][r] = s1[s];
        }
        if (p.pre_phase) {
#pragma unroll
            for (int r = 0; r < RA; ++r) {
                const int l = tb_first | (r << g_first);
                const u64 d = tbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
                double2 ph;
                if (use_lut) ph = lut_sm[p.hidx[d]];
                else {
                    double sn, cs;
                    sincos(p.angle_pre * p.ham[d], &sn, &cs);
                    ph = make_double2(cs, -sn);
                }
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
            }
        }
#ifndef QR_HOST_EMUL
        if (p.prefetch) {   // pull the next tile of this CTA into L2 while this one is computed
            const i64 nt = tile + (i64)gridDim.x * p.prefetch;   // p.prefetch = distance in tiles
            if (nt < p.num_tiles) {
                const i64 nb = nt >> p.tiles_log2;
                const u64 t2 = (u64)nt & tmask;
                const u64 nbase = ((t2 & (((u64)1 << nlo) - 1)) << c) | ((t2 >> nlo) << (h + p.k - c));
                for (int line = tid; line < (T >> 3); line += blockDim.x) {
                    const int l = line << 3;
                    const u64 d = nbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
                    const u64 s = p.ladder ? (ladder_map(d, p.M1, p.M2) ^ p.src_xor) : d;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src0 + nb * p.state_stride + s));
                    if (NV == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src1 + nb * p.state_stride + s));
                }
            }
        }
#endif
        // ---- rounds: gates in registers, exchange through swizzled shared memory ----
#pragma unroll 1
        for (int rd = 0; rd < nrounds; ++rd) {
            double acc[4] = {0.0, 0.0, 0.0, 0.0};
            qr_round_compute<NV, R>(a, gt + rd * R, acc);
            if (NV == 2) {   // static indices keep the accumulators in registers
#pragma unroll
                for (int rr = 0; rr < QR_MAXROUNDS; ++rr)
                    if (rr == rd) {
#pragma unroll
                        for (int i = 0; i < R; ++i)
                            if (rr * R + i < QR_GATE_SLOTS) acc_all[rr * R + i] += acc[i];
                    }
            }
            if (rd + 1 < nrounds) {
                const int gp = p.g[rd], gn = p.g[rd + 1];
                const int tbp = (tid & ((1 << gp) - 1)) | ((tid >> gp) << (gp + R));
                const int tbn = (tid & ((1 << gn) - 1)) | ((tid >> gn) << (gn + R));
#pragma unroll
                for (int r = 0; r < RA; ++r) {
                    const int l = qr_swz<R>(tbp | (r << gp));
#pragma unroll
                    for (int v = 0; v < NV; ++v) exch[v * T + l] = a[v][r];
                }
                __syncthreads();
#pragma unroll
                for (int r = 0; r < RA; ++r) {
                    const int l = qr_swz<R>(tbn | (r << gn));
#pragma unroll
                    for (int v = 0; v < NV; ++v) a[v][r] = exch[v * T + l];
                }
                if (rd + 2 == nrounds) __syncthreads();   // smem is free for the next tile's first exchange
            }
        }
        // ---- registers -> global (QAOA: diagonal-generator inner product and un-phase) ----
#pragma unroll
        for (int r = 0; r < RA; ++r) {
            const int l = tb_last | (r << g_last);
            const u64 d = tbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
            if (p.post_phase) {
                double hv;
                double2 ph;
                if (use_lut) {
                    const int hi = p.hidx[d];
                    hv = p.hmin + (double)hi;
                    ph = lut_sm[hi];
                } else {
                    hv = p.ham[d];
                    double sn, cs;
                    sincos(p.angle_post * hv, &sn, &cs);
                    ph = make_double2(cs, -sn);
                }
                if (NV == 2) acc_all[QR_SLOTS - 1] += hv * im_conj_mul(a[NV - 1][r], a[0][r]);
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
            }
            d0[d] = a[0][r];
            if (NV == 2) d1[d] = a[NV - 1][r];
        }
        if (NV == 2 && p.flush_per_tile) {
            qr_block_reduce_slots(acc_all, p.partials + (u64)tile * QR_SLOTS);
#pragma unroll
            for (int i = 0; i < QR_SLOTS; ++i) acc_all[i] = 0.0;
        }
    }
    if (NV == 2 && !p.flush_per_tile) {
        qr_block_reduce_slots(acc_all, p.partials + (u64)blockIdx.x * QR_SLOTS);
        // second stage fused in: the last CTA to arrive adds the per-CTA partials in CTA order
        // (fixed order => run-to-run deterministic) and writes the QR_SLOTS sums of this pass.
        if (p.final_out) {
            __shared__ int is_last;
            __threadfence();
            if (tid == 0) {
                const unsigned prev = atomicAdd(p.done_counter, 1u);
                is_last = (prev + 1 == gridDim.x);
            }
            __syncthreads();
            if (is_last) {
                __threadfence();
                for (int i = tid; i < QR_SLOTS; i += blockDim.x) {
                    double v = 0.0;
                    for (unsigned b = 0; b < gridDim.x; ++b) v += ((volatile double*)p.partials)[(u64)b * QR_SLOTS + i];
                    p.final_out[i] = v;
                }
                if (tid == 0) *p.done_counter = 0u;   // re-arm for the next launch on this stream
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// Sharded states (top log2(G) qubits = rank bits): rotations on the GLOBAL qubits.
//
// One kernel does the exchange and the gates together over peer memory: rank r owns the slice
// [r*N_loc/G, (r+1)*N_loc/G) of the local index range and, for every index in it, loads the G
// amplitudes that differ only in the rank bits straight from the G peer shards (NVLink P2P
// loads through CUDA-IPC mappings), applies the log2(G) single-qubit gates in registers
// (backward: also the Im<lambda|P|psi> partials and both vectors), and stores them back to the
// peers.  No staging buffer, no layout change; 7/8 of a shard crosses NVLink in each direction
// per layer and vector, overlapped with the arithmetic by the usual load/store pipelining.
// ------------------------------------------------------------------------------------------
#define QR_MAX_RANKS 16
// Only X / Y rotations need the exchange.  An Rz on a global qubit is diagonal: the shards split into independent
// subgroups (fixed values of the Z bits), each subgroup exchanges among its 2^ga members only ((2^ga - 1) / 2^ga of a
// shard crosses NVLink instead of (G - 1) / G; nothing at all when every global gate of the layer is an Rz), the Z
// phases of a subgroup are ONE constant factor, and the Z gradients are the signed total of Im(conj(lambda) psi).
struct GlobalGates {
    int ga;                       // rank bits with X / Y rotations (the exchanged ones), 0 .. log2(ranks)
    u64 slice_off, slice_len;     // local index range handled by this rank: 1 / 2^ga of the shard
    double2* psi[QR_MAX_RANKS];   // the 2^ga shards of this rank's subgroup, ordered by the active bits of the LOGICAL shard id
    double2* lam[QR_MAX_RANKS];
    GateP gate[4];                // X / Y gate on active bit i
    int slot[4];                  // result slot of active gate i (= its rank bit b, qubit g-1-b)
    int nz;                       // Rz gates on the other rank bits
    int zslot[4];                 // their result slots
    double zsign[4];              // +1 if this subgroup's value of the bit is 0, else -1
    double2 zphase;               // product of their phases for this subgroup
    double* partials;             // [grid][4]
};

template <int NV, int GA>
__global__ void __launch_bounds__(256) k_global_gates(const GlobalGates p) {
    constexpr int NA = 1 << GA;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    double wtot = 0.0;
    const bool zg = p.nz > 0;
    const double2 zp = p.zphase;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < p.slice_len; j += (u64)gridDim.x * blockDim.x) {
        const u64 idx = p.slice_off + j;
        double2 a[NV][NA];
#pragma unroll
        for (int s = 0; s < NA; ++s) {
            a[0][s] = p.psi[s][idx];
            if (NV == 2) a[NV - 1][s] = p.lam[s][idx];
        }
        if (NV == 2 && zg) {
#pragma unroll
            for (int s = 0; s < NA; ++s) wtot += im_conj_mul(a[NV - 1][s], a[0][s]);
        }
        if (GA > 0) qr_gate_on_bit<NV, NA, 0>(a, p.gate[0], acc[0]);
        if (GA > 1) qr_gate_on_bit<NV, NA, (GA > 1 ? 1 : 0)>(a, p.gate[1], acc[1]);
        if (GA > 2) qr_gate_on_bit<NV, NA, (GA > 2 ? 2 : 0)>(a, p.gate[2], acc[2]);
        if (GA > 3) qr_gate_on_bit<NV, NA, (GA > 3 ? 3 : 0)>(a, p.gate[3], acc[3]);
#pragma unroll
        for (int s = 0; s < NA; ++s) {
            if (zg) {
                a[0][s] = cmul(a[0][s], zp);
                if (NV == 2) a[NV - 1][s] = cmul(a[NV - 1][s], zp);
            }
            p.psi[s][idx] = a[0][s];
            if (NV == 2) p.lam[s][idx] = a[NV - 1][s];
        }
    }
    if (NV == 2) {
        double out[4] = {0.0, 0.0, 0.0, 0.0};
        const double wsum = block_reduce_sum(wtot);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const double s = block_reduce_sum(acc[i]);
            if (i < GA) out[p.slot[i] & 3] = s;
        }
        for (int k = 0; k < p.nz; ++k) out[p.zslot[k] & 3] = p.zsign[k] * wsum;
        if (threadIdx.x == 0)
            for (int i = 0; i < 4; ++i) p.partials[(u64)blockIdx.x * 4 + i] = out[i];
    }
}

__global__ void k_reduce_slots_strided(const double* __restrict__ partial, int units, u64 group_stride, double* __restrict__ out,
                                       int out_stride) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* base = partial + (u64)blockIdx.x * group_stride;
    double v = 0.0;
    if (w < QR_SLOTS)
        for (int u = lane; u < units; u += 32) v += base[(u64)u * QR_SLOTS + w];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (w < QR_SLOTS && lane == 0) out[(u64)blockIdx.x * out_stride + w] = v;
}
