// Lean fused tile pass for full 12-bit tiles (the default hot kernel for n >= 12).
//
// Same role, tile geometry, gather and result layout as k_tile_pass (qr_tile.cuh): one launch
// sweeps the state vector once (forward: psi; backward: psi and the co-state lambda) and applies
// every rotation whose index bit lies in the tile (mc_clean.py:38-41 forward, :68-77 backward;
// qaoa.py:49-53, :59-69).  What differs is the arithmetic and the bookkeeping, both chosen from the
// ncu instruction mix of the generic kernel (profiles/README.md: 133 FP64 + ~190 other
// instructions per amplitude):
//
//  * Static geometry.  k = 12, 512 threads, 8 amplitudes per vector per thread; the register
//    groups are fixed, G0 = local bits 0-2, G1 = 3-5, G2 = 6-8, G3 = 9-11, visited in the order
//    G3 [G0] [G1] [G2] (G3 first and G2 last keep lane <-> local bits 0-4, i.e. 512 B per warp
//    access).  All shared-memory addresses are a per-thread base XOR a compile-time constant, all
//    global addresses a per-thread base XOR a per-register constant held in the constant bank,
//    and the four rounds are unrolled, so no register shuffling or shift/mask chains remain.
//  * tan-form rotations.  Rx/Ry with (c, s) are applied as f * [[1, -t], [t, 1]]-type updates with
//    f = the larger of |c|, |s| and t = the ratio: ONE fma per real component instead of mul+fma.
//    The scalar F = prod f of the pass is applied once per amplitude (folded into the Z phase
//    below when there is one).  Gradient partials taken on the not-yet-rescaled registers are
//    corrected by a per-gate constant (1 / remaining scale^2) when the CTA flushes them.
//  * Merged Z rotations.  All Rz of the pass are one diagonal: amplitude (thread t, register r)
//    is multiplied by zt(t) * zr[r]; zt is computed once per CTA, zr is an 8-entry table.
//  * Z gradients from one product.  Im<lambda|Z_q|psi> = sum_j (+-) w_j with
//    w_j = Im(conj(lambda_j) psi_j), which is invariant under the pass's diagonal and taken right
//    after the load: 2 flops per amplitude for ALL Z gates of the pass; for a Z gate on a thread
//    bit the sign is a per-thread constant, so the per-thread total of w is kept in one register
//    for the whole kernel and signed at the end.
//
// FP64 instructions per amplitude of a 12-gate backward pass: ~65 (was 133).
//
// Roofline: HBM.  Algorithmic bytes per launch = NV * 32 B * 2^n.
#pragma once
#include "qr_tile.cuh"

#define QR_T12_THREADS 512

struct Tile12X {
    int ngroups;            // active register groups: 1 = G3; 2 = G3,G2; 3 = G3,G1,G2; 4 = G3,G0,G1,G2
    u64 roff_first[8];      // global offset of register r at load time (group G3), gather map applied
    u64 roff_last[8];       // global offset of register r at store time (G2, or G3 when ngroups == 1)
    u64 droff_first[8];     // same as roff_first without the gather map (destination index: phase tables)
};

// converted gate of local bit b (shared memory, rebuilt per batch element)
struct Gate12 {
    double t;      // ratio (tan-form)
    double corr;   // gradient correction 1 / (scale still to come)^2
    int mode;      // -1 none; 0 X |c|>=|s|; 1 X |c|<|s|; 2 Y lo; 3 Y hi; 4 Z
    int pad;
};

template <int NV, int BIT>
__device__ __forceinline__ void qr12_gate(double2 (&a)[NV][8], const Gate12& g, double& acc) {
    const int m = g.mode;
    if (m < 0 || m > 3) return;
    const double t = g.t;
    if (m < 2) {   // X  (state.py:90-92)
        if (NV == 2) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r & (1 << BIT)) continue;
                const int r1 = r | (1 << BIT);
                s += im_conj_mul(a[NV - 1][r], a[0][r1]) + im_conj_mul(a[NV - 1][r1], a[0][r]);
            }
            acc += s;
        }
        if (m == 0) {   // f (a - i t b), f (b - i t a)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r & (1 << BIT)) continue;
                const int r1 = r | (1 << BIT);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double2 x = a[v][r], y = a[v][r1];
                    a[v][r] = make_double2(x.x + t * y.y, x.y - t * y.x);
                    a[v][r1] = make_double2(y.x + t * x.y, y.y - t * x.x);
                }
            }
        } else {        // f (t a - i b), f (t b - i a)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r & (1 << BIT)) continue;
                const int r1 = r | (1 << BIT);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double2 x = a[v][r], y = a[v][r1];
                    a[v][r] = make_double2(t * x.x + y.y, t * x.y - y.x);
                    a[v][r1] = make_double2(t * y.x + x.y, t * y.y - x.x);
                }
            }
        }
    } else {       // Y  (state.py:142-144)
        if (NV == 2) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r & (1 << BIT)) continue;
                const int r1 = r | (1 << BIT);
                s += re_conj_mul(a[NV - 1][r1], a[0][r]) - re_conj_mul(a[NV - 1][r], a[0][r1]);
            }
            acc += s;
        }
        if (m == 2) {   // f (a - t b), f (b + t a)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r & (1 << BIT)) continue;
                const int r1 = r | (1 << BIT);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double2 x = a[v][r], y = a[v][r1];
                    a[v][r] = make_double2(x.x - t * y.x, x.y - t * y.y);
                    a[v][r1] = make_double2(y.x + t * x.x, y.y + t * x.y);
                }
            }
        } else {        // f (t a - b), f (a + t b)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                if (r & (1 << BIT)) continue;
                const int r1 = r | (1 << BIT);
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double2 x = a[v][r], y = a[v][r1];
                    a[v][r] = make_double2(t * x.x - y.x, t * x.y - y.y);
                    a[v][r1] = make_double2(x.x + t * y.x, x.y + t * y.y);
                }
            }
        }
    }
}

// gates of the register group whose first local bit is G (slots G, G+1, G+2)
template <int NV, int G>
__device__ __forceinline__ void qr12_round(double2 (&a)[NV][8], const Gate12* sg, double (&acc)[QR_SLOTS]) {
    qr12_gate<NV, 0>(a, sg[G + 0], acc[G + 0]);
    qr12_gate<NV, 1>(a, sg[G + 1], acc[G + 1]);
    qr12_gate<NV, 2>(a, sg[G + 2], acc[G + 2]);
}

// thread's local-index base when the register group starts at local bit g (3 zero bits inserted)
__device__ __forceinline__ int qr12_tb(int tid, int g) { return (tid & ((1 << g) - 1)) | ((tid >> g) << (g + 3)); }

// swizzled shared-memory index of register r for register group G: base ^ (r * MUL)
template <int G>
__device__ __forceinline__ int qr12_sbase(int tid) {
    const int tb = qr12_tb(tid, G);
    if (G == 0) return tb ^ ((tb >> 3) & 7);   // low bits: r ^ (bits 3-5)
    if (G == 3) return tb;                      // low bits ^ r, bits 3-5 = r
    return tb ^ ((tb >> 3) & 7);                // G >= 6: swizzle does not involve r
}
template <int G>
__device__ __forceinline__ constexpr int qr12_smul() { return G == 0 ? 1 : (G == 3 ? 9 : (1 << G)); }

// registers of group GP -> shared memory -> registers of group GN (one block barrier)
template <int NV, int GP, int GN>
__device__ __forceinline__ void qr12_exchange(double2 (&a)[NV][8], double2* smem, int tid) {
    constexpr int T = 1 << QR_MAX_TILE_BITS;
    const int bp = qr12_sbase<GP>(tid), bn = qr12_sbase<GN>(tid);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int l = bp ^ (r * qr12_smul<GP>());
#pragma unroll
        for (int v = 0; v < NV; ++v) smem[v * T + l] = a[v][r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int l = bn ^ (r * qr12_smul<GN>());
#pragma unroll
        for (int v = 0; v < NV; ++v) a[v][r] = smem[v * T + l];
    }
}

// diagonal phase exp(-i angle H[d]) from the integer look-up table or, for general H, sincos
__device__ __noinline__ double2 qr12_phase_slow(const double* __restrict__ ham, u64 d, double angle, double* hv) {
    const double v = ham[d];
    double sn, cs;
    sincos(angle * v, &sn, &cs);
    *hv = v;
    return make_double2(cs, -sn);
}

// PHASE: QAOA diagonal phase before the gates (forward) / generator inner product + un-phase after them
// (backward); compiled out of the McClean instantiations.
template <int NV, bool PHASE>
__global__ void __launch_bounds__(QR_T12_THREADS, (NV == 1 ? 2 : 1)) k_tile12(const TilePass p, const Tile12X x) {
    constexpr int T = 1 << QR_MAX_TILE_BITS;
    QR_DYN_SMEM(double2, smem);
    __shared__ Gate12 sg[QR_GATE_SLOTS];
    __shared__ double2 szr[8];                  // Z phases of the G3 register bits (times nothing else)
    __shared__ double2 szb[QR_GATE_SLOTS][2];   // per gate bit: Z phase for bit value 0 / 1 (identity if not Z)
    __shared__ double s_scale[2];               // F = product of the tan-form factors; has_z flag
    __shared__ double2 lut_sm[QR_LUT_MAX];
    const int tid = threadIdx.x;
    const int c = p.c, h = p.h;
    const int lomask = (1 << c) - 1;
    const int nlo = h - c;
    const u64 tmask = ((u64)1 << p.tiles_log2) - 1;
    const int ng = x.ngroups;

    double acc_all[QR_SLOTS];
#pragma unroll
    for (int i = 0; i < QR_SLOTS; ++i) acc_all[i] = 0.0;
    double wtot = 0.0;   // running sum of Im(conj(lambda) psi) over this thread's amplitudes

    // per-thread global offsets (local bits 0-8 at load time; the last group's thread bits at store time)
    const u64 toff_d = (u64)(tid & lomask) | ((u64)(tid >> c) << h);                     // destination index bits
    const u64 toff_s = p.ladder ? ladder_map(toff_d, p.M1, p.M2) : toff_d;               // gathered source bits
    const int tbl = ng > 1 ? qr12_tb(tid, 6) : tid;
    const u64 toff_l = (u64)(tbl & lomask) | ((u64)(tbl >> c) << h);

    const bool use_lut = PHASE && p.hidx != nullptr && (p.pre_phase || p.post_phase);
    if (use_lut) {
        for (int i = tid; i < p.lut_size; i += blockDim.x) lut_sm[i] = p.lut[i];
    }
    i64 cur_b = -1;
    double2 zt = make_double2(1.0, 0.0);   // thread factor of the merged diagonal (includes F)
    double fscale = 1.0;
    bool has_z = false;

    // flush helper state: sign pattern of the thread-bit Z gates is applied when partials leave the thread
    auto finalize = [&]() {
#pragma unroll
        for (int b = 0; b < QR_GATE_SLOTS; ++b) {
            const int m = sg[b].mode;
            double v = acc_all[b];
            if (b < 9 && m == 4) v = ((tid >> b) & 1) ? -wtot : wtot;
            acc_all[b] = v * sg[b].corr;
        }
    };

    for (i64 tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const i64 b = tile >> p.tiles_log2;
        const u64 t = (u64)tile & tmask;
        const u64 tbase = ((t & (((u64)1 << nlo) - 1)) << c) | ((t >> nlo) << (h + QR_MAX_TILE_BITS - c));
        if (b != cur_b) {   // block-uniform: convert the gate table of this batch element
            __syncthreads();
            if (tid < QR_GATE_SLOTS) {
                const GateP g = p.gates[b * p.gate_stride + tid];
                Gate12 o;
                o.t = 0.0; o.corr = 1.0; o.mode = -1; o.pad = 0;
                double2 z0 = make_double2(1.0, 0.0), z1 = z0;
                double f = 1.0;
                if (g.axis == 0 || g.axis == 1) {
                    const bool lo = fabs(g.c) >= fabs(g.s);
                    f = lo ? g.c : g.s;
                    o.t = lo ? g.s / g.c : g.c / g.s;
                    o.mode = g.axis * 2 + (lo ? 0 : 1);
                } else if (g.axis == 2) {   // Rz: (c - i s) on bit value 0, (c + i s) on bit value 1 (state.py:168-170)
                    o.mode = 4;
                    z0 = make_double2(g.c, -g.s);
                    z1 = make_double2(g.c, g.s);
                }
                o.corr = f;   // temporarily: this gate's factor
                sg[tid] = o;
                szb[tid][0] = z0;
                szb[tid][1] = z1;
            }
            __syncthreads();
            if (tid == 0) {
                // application order G3, G0, G1, G2: a gate's partial sees the factors of the gates after it
                const int ord[12] = {9, 10, 11, 0, 1, 2, 3, 4, 5, 6, 7, 8};
                double rem = 1.0;   // product of the factors not yet applied
                bool z = false;
                for (int i = 11; i >= 0; --i) {
                    const int bb = ord[i];
                    rem *= sg[bb].corr;   // a partial is taken before its own gate: its factor is still to come too
                    const bool isz = sg[bb].mode == 4;
                    sg[bb].corr = isz ? 1.0 : 1.0 / (rem * rem);   // Z partials come from the raw (unscaled) load
                    z = z || isz;
                }
                s_scale[0] = rem;
                s_scale[1] = z ? 1.0 : 0.0;
            }
            if (tid < 8) {
                double2 z = make_double2(1.0, 0.0);
#pragma unroll
                for (int j = 0; j < 3; ++j) z = cmul(z, szb[9 + j][(tid >> j) & 1]);
                szr[tid] = z;
            }
            __syncthreads();
            fscale = s_scale[0];
            has_z = s_scale[1] != 0.0;
            zt = make_double2(fscale, 0.0);
#pragma unroll
            for (int j = 0; j < 9; ++j) zt = cmul(zt, szb[j][(tid >> j) & 1]);
            cur_b = b;
        }
        const double2* __restrict__ s0 = p.src0 + b * p.state_stride;
        const double2* __restrict__ s1 = (NV == 2) ? p.src1 + b * p.state_stride : nullptr;
        double2* __restrict__ d0 = p.dst0 + b * p.state_stride;
        double2* __restrict__ d1 = (NV == 2) ? p.dst1 + b * p.state_stride : nullptr;

        // ---- global -> registers (group G3; ladder gather folded into the load addresses) ----
        const u64 sbt = (p.ladder ? (ladder_map(tbase, p.M1, p.M2) ^ p.src_xor) : tbase) ^ toff_s;
        double2 a[NV][8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const u64 s = sbt ^ x.roff_first[r];
            a[0][r] = s0[s];
            if (NV == 2) a[NV - 1][r] = s1[s];
        }
#ifndef QR_HOST_EMUL
        if (p.prefetch) {   // pull the next tile of this CTA into L2 while this one is computed
            const i64 nt = tile + (i64)gridDim.x * p.prefetch;
            if (nt < p.num_tiles) {
                const i64 nb = nt >> p.tiles_log2;
                const u64 t2 = (u64)nt & tmask;
                const u64 nbase = ((t2 & (((u64)1 << nlo) - 1)) << c) | ((t2 >> nlo) << (h + QR_MAX_TILE_BITS - c));
                const int l = tid << 3;   // one 128 B line per thread
                const u64 d = nbase | (u64)(l & lomask) | ((u64)(l >> c) << h);
                const u64 s = p.ladder ? (ladder_map(d, p.M1, p.M2) ^ p.src_xor) : d;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src0 + nb * p.state_stride + s));
                if (NV == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src1 + nb * p.state_stride + s));
            }
        }
#endif
        // ---- Z gradients: w = Im(conj(lambda) psi), signed sums over the register bits, total for the thread bits ----
        if (NV == 2 && has_z) {
            double w[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) w[r] = im_conj_mul(a[NV - 1][r], a[0][r]);
            const double e0 = w[0] + w[1], o0 = w[0] - w[1], e1 = w[2] + w[3], o1 = w[2] - w[3];
            const double e2 = w[4] + w[5], o2 = w[4] - w[5], e3 = w[6] + w[7], o3 = w[6] - w[7];
            const double ee0 = e0 + e1, eo0 = e0 - e1, ee1 = e2 + e3, eo1 = e2 - e3;
            wtot += ee0 + ee1;
            if (sg[9].mode == 4) acc_all[9] += (o0 + o1) + (o2 + o3);
            if (sg[10].mode == 4) acc_all[10] += eo0 + eo1;
            if (sg[11].mode == 4) acc_all[11] += ee0 - ee1;
        }
        // ---- QAOA forward: exp(-i gamma H) before the mixer ----
        if (PHASE && p.pre_phase) {
            const u64 dbt = tbase | toff_d;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const u64 d = dbt | x.droff_first[r];
                double2 ph;
                double hv;
                if (use_lut) ph = lut_sm[p.hidx[d]];
                else ph = qr12_phase_slow(p.ham, d, p.angle_pre, &hv);
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
            }
        }
        // ---- merged diagonal (all Rz of the pass) and the pass scale F ----
        if (has_z) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 ph = cmul(zt, szr[r]);
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
            }
        } else if (fscale != 1.0) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = make_double2(a[v][r].x * fscale, a[v][r].y * fscale);
            }
        }
        // ---- rounds ----
        qr12_round<NV, 9>(a, sg, acc_all);
        if (ng == 4) {
            qr12_exchange<NV, 9, 0>(a, smem, tid);
            qr12_round<NV, 0>(a, sg, acc_all);
            qr12_exchange<NV, 0, 3>(a, smem, tid);
        } else if (ng == 3) {
            qr12_exchange<NV, 9, 3>(a, smem, tid);
        }
        if (ng >= 3) {
            qr12_round<NV, 3>(a, sg, acc_all);
            qr12_exchange<NV, 3, 6>(a, smem, tid);
        } else if (ng == 2) {
            qr12_exchange<NV, 9, 6>(a, smem, tid);
        }
        if (ng >= 2) {
            qr12_round<NV, 6>(a, sg, acc_all);
            __syncthreads();   // every thread has read its last exchange: smem is free for the next tile
        }
        // ---- registers -> global (QAOA backward: diagonal-generator inner product and un-phase) ----
        const u64 dlt = tbase | toff_l;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const u64 d = dlt | x.roff_last[r];
            if (PHASE && p.post_phase) {
                double hv;
                double2 ph;
                if (use_lut) {
                    const int hi = p.hidx[d];
                    hv = p.hmin + (double)hi;
                    ph = lut_sm[hi];
                } else ph = qr12_phase_slow(p.ham, d, p.angle_post, &hv);
                if (NV == 2) acc_all[QR_SLOTS - 1] += hv * im_conj_mul(a[NV - 1][r], a[0][r]);
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
            }
            d0[d] = a[0][r];
            if (NV == 2) d1[d] = a[NV - 1][r];
        }
        if (NV == 2 && p.flush_per_tile) {
            finalize();
            qr_block_reduce_slots(acc_all, p.partials + (u64)tile * QR_SLOTS);
#pragma unroll
            for (int i = 0; i < QR_SLOTS; ++i) acc_all[i] = 0.0;
            wtot = 0.0;
        }
    }
    if (NV == 2 && !p.flush_per_tile) {
        if (cur_b >= 0) finalize();
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
                    for (unsigned bb = 0; bb < gridDim.x; ++bb) v += ((volatile double*)p.partials)[(u64)bb * QR_SLOTS + i];
                    p.final_out[i] = v;
                }
                if (tid == 0) *p.done_counter = 0u;   // re-arm for the next launch on this stream
            }
        }
    }
}
