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
//  * Rotations as three in-place shears (lifting steps).  [[c, -s], [s, c]] with c >= 0 is
//    a -= tau b;  b += s a;  a -= tau b   with tau = s / (1 + c) = tan(phi/2), |tau| <= 1
//    (c < 0: the same for (-c, -s) and a sign that is folded into the pass's diagonal).  Three fma
//    per real component pair instead of two mul + two fma, exactly unitary, and -- the point --
//    every step overwrites one operand with a function of the other, so no temporaries: the
//    two-output form (new a and new b both from old a and old b) cost 32 live temporaries per gate,
//    one MOV per FMA after branch joins and local-memory spills on every tile's critical path.
//  * Merged Z rotations.  All Rz of the pass are one diagonal: amplitude (thread t, register r)
//    is multiplied by zt(t) * zr[r]; zt is computed once per CTA, zr is an 8-entry table.
//  * Z gradients from one product.  Im<lambda|Z_q|psi> = sum_j (+-) w_j with
//    w_j = Im(conj(lambda_j) psi_j), which is invariant under the pass's diagonal and taken right
//    after the load: 2 flops per amplitude for ALL Z gates of the pass; for a Z gate on a thread
//    bit the sign is a per-thread constant, so the per-thread total of w is kept in one register
//    for the whole kernel and signed at the end.
//
// FP64 instructions per amplitude of a 12-gate backward pass: ~80 (was 133), no register moves.
//
// Roofline: HBM.  Algorithmic bytes per launch = NV * 32 B * 2^n.
#pragma once
#include "qr_tile.cuh"

#define QR_T12_THREADS 512

// Two-segment tile geometry (k = 12): local bits [0,c) are global bits [0,c); local bits [c,c+m1) are
// global bits [h,h+m1); local bits [c+m1,12) are global bits [h2,h2+m2).  The planner uses the
// second segment to spread the page-selecting index bits (>= 2 MiB) over the strided passes, so
// no pass touches more than ~128 distinct pages per tile and vector (a pass whose 512 rows lie
// in 512 different pages runs 1.7x slower in the two-vector backward sweep: TLB reach).
struct Geo12 { int c, h, m1, h2, k; };   // k = tile bits (12, or 11 for the half-size tiles)
__host__ __device__ __forceinline__ u64 geo12_local(const Geo12 g, u64 l) {
    return (l & (((u64)1 << g.c) - 1)) | (((l >> g.c) & (((u64)1 << g.m1) - 1)) << g.h) | ((l >> (g.c + g.m1)) << g.h2);
}
__host__ __device__ __forceinline__ u64 geo12_tile(const Geo12 g, u64 t) {
    const int nlo = g.h - g.c, nmid = g.h2 - g.h - g.m1, m2 = g.k - g.c - g.m1;
    return ((t & (((u64)1 << nlo) - 1)) << g.c) | (((t >> nlo) & (((u64)1 << nmid) - 1)) << (g.h + g.m1)) |
           ((t >> (nlo + nmid)) << (g.h2 + m2));
}

struct Tile12X {
    int ngroups;            // chain of register groups (first local bit of each; K = tile bits, L = K-3):
                            // 1 = L; 2 = L,6; 3 = L,3,6; 4 = L,0,3,6; 5 (K = 11, 64 B rows) = L,2,5
    int last_group;         // first local bit of the register group held at store time
    int cluster;            // thread-block cluster size of the launch (1 or 2)
    int cache_hints;        // bit0: streaming (evict-first) stores, bit1: streaming loads
    int pair_order;         // strided passes: a CTA takes its tiles in ADJACENT pairs (t, t ^ 1: the two 128 B halves of the same
                            // 256 B chunks) and fetches the partner / the next pair into L2 together, so DRAM sees 256 B requests
    u64 roff_first[8];      // global offset of register r at load time (group G3), gather map applied
    u64 roff_last[8];       // global offset of register r at store time (G2, or G3 when ngroups == 1)
    u64 droff_first[8];     // same as roff_first without the gather map (destination index: phase tables)
};

// converted gate of local bit b (shared memory, rebuilt per batch element)
struct Gate12 {
    double tau;    // tan(phi/2) of the reduced rotation
    double sig;    // sin(phi) of the reduced rotation
    int mode;      // -1 none; 0 X; 1 Y; 4 Z
    int neg;       // reduced rotation = -(rotation): sign folded into the diagonal of the pass
};

template <int NV, int BIT>
__device__ __forceinline__ void qr12_gate(double2 (&a)[NV][8], const Gate12& g, double& acc) {
    const int m = g.mode;
    if (m != 0 && m != 1) return;
    const double tau = g.tau, sig = g.sig;
    if (m == 0) {   // X (state.py:90-92): a' = c a - i s b, b' = -i s a + c b
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
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
#pragma unroll
            for (int v = 0; v < NV; ++v) {   // a -= i tau b ; b -= i sig a ; a -= i tau b
                a[v][r].x += tau * a[v][r1].y;
                a[v][r].y -= tau * a[v][r1].x;
                a[v][r1].x += sig * a[v][r].y;
                a[v][r1].y -= sig * a[v][r].x;
                a[v][r].x += tau * a[v][r1].y;
                a[v][r].y -= tau * a[v][r1].x;
            }
        }
    } else {        // Y (state.py:142-144): a' = c a - s b, b' = s a + c b
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
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
#pragma unroll
            for (int v = 0; v < NV; ++v) {   // a -= tau b ; b += sig a ; a -= tau b
                a[v][r].x -= tau * a[v][r1].x;
                a[v][r].y -= tau * a[v][r1].y;
                a[v][r1].x += sig * a[v][r].x;
                a[v][r1].y += sig * a[v][r].y;
                a[v][r].x -= tau * a[v][r1].x;
                a[v][r].y -= tau * a[v][r1].y;
            }
        }
    }
}

// gates of the register group whose first local bit is G (slots G, G+1, G+2)
template <int NV, int G, int NB = 3>
__device__ __forceinline__ void qr12_round(double2 (&a)[NV][8], const Gate12* sg, double (&acc)[QR_SLOTS]) {
    qr12_gate<NV, 0>(a, sg[G + 0], acc[G + 0]);
    if (NB > 1) qr12_gate<NV, 1>(a, sg[G + (NB > 1 ? 1 : 0)], acc[G + (NB > 1 ? 1 : 0)]);
    if (NB > 2) qr12_gate<NV, 2>(a, sg[G + (NB > 2 ? 2 : 0)], acc[G + (NB > 2 ? 2 : 0)]);
}

// thread's local-index base when the register group starts at local bit g (3 zero bits inserted)
__device__ __forceinline__ int qr12_tb(int tid, int g) { return (tid & ((1 << g) - 1)) | ((tid >> g) << (g + 3)); }

// swizzled shared-memory index sw(l) = l ^ ((l >> 3) & 7) of register r for register group G
// (l = tb | r << G): a per-thread base XOR a compile-time constant per register.
template <int G>
__device__ __forceinline__ int qr12_sbase(int tid) {
    const int tb = qr12_tb(tid, G);
    return tb ^ ((tb >> 3) & 7);
}
template <int G>
__device__ __forceinline__ constexpr int qr12_cr(int r) { return (r << G) ^ (((r << G) >> 3) & 7); }

// registers of group GP -> shared memory -> registers of group GN (one block barrier)
template <int NV, int GP, int GN, int K = QR_MAX_TILE_BITS>
__device__ __forceinline__ void qr12_exchange(double2 (&a)[NV][8], double2* smem, int tid) {
    constexpr int T = 1 << K;
    const int bp = qr12_sbase<GP>(tid), bn = qr12_sbase<GN>(tid);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int l = bp ^ qr12_cr<GP>(r);
#pragma unroll
        for (int v = 0; v < NV; ++v) smem[v * T + l] = a[v][r];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int l = bn ^ qr12_cr<GN>(r);
#pragma unroll
        for (int v = 0; v < NV; ++v) a[v][r] = smem[v * T + l];
    }
}

// same exchange through ONE tile-sized buffer: the vectors take turns (staged kernel: the other
// 128 KiB of shared memory hold the next tile).  The buffer holds vector NV-1 when it returns.
template <int NV, int GP, int GN>
__device__ __forceinline__ void qr12_exchange_1buf(double2 (&a)[NV][8], double2* xbuf, int tid) {
    const int bp = qr12_sbase<GP>(tid), bn = qr12_sbase<GN>(tid);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        // safe without a barrier: this thread overwrites only the slots it read itself last time
        // (v == 0: its GP slots of the previous exchange's last vector; v > 0: needs the barrier below)
        if (v > 0) __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) xbuf[bp ^ qr12_cr<GP>(r)] = a[v][r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) a[v][r] = xbuf[bn ^ qr12_cr<GN>(r)];
    }
}

// ---- split rounds (QR_T12_SPLIT_XCHG): the shared-memory writes of psi are issued between the FP64 work of psi and
// lambda, and the first FP64 instructions of the next round need psi only, so the shared-memory pipe and the FP64 pipe
// overlap inside a warp instead of taking turns (ncu at n = 30: mio_throttle 2.0, math_pipe_throttle 1.5 stalls per issue
// with the all-FP64-then-all-exchange order).  Round = shear psi; write psi; shear lambda; inner products; write lambda.
// The inner products are taken AFTER the un-rotations of the round: <lambda|P_q|psi> is invariant under a rotation about
// P_q and under rotations of other qubits as long as they are applied to BOTH vectors.
// MEASURED SLOWER (profiles/r1_ab_split_rounds.log: backward sweep +1 % at n = 30, +2 % at n = 26 and n = 20), so it is
// compiled out by default and kept for A/B builds (-DQR_T12_SPLIT_XCHG=1).
#ifndef QR_T12_SPLIT_XCHG
#define QR_T12_SPLIT_XCHG 0
#endif
template <int BIT>
__device__ __forceinline__ void qr12_shear(double2 (&av)[8], const Gate12& g) {
    const int m = g.mode;
    if (m != 0 && m != 1) return;
    const double tau = g.tau, sig = g.sig;
    if (m == 0) {   // a -= i tau b ; b -= i sig a ; a -= i tau b
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
            av[r].x += tau * av[r1].y;
            av[r].y -= tau * av[r1].x;
            av[r1].x += sig * av[r].y;
            av[r1].y -= sig * av[r].x;
            av[r].x += tau * av[r1].y;
            av[r].y -= tau * av[r1].x;
        }
    } else {        // a -= tau b ; b += sig a ; a -= tau b
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
            av[r].x -= tau * av[r1].x;
            av[r].y -= tau * av[r1].y;
            av[r1].x += sig * av[r].x;
            av[r1].y += sig * av[r].y;
            av[r].x -= tau * av[r1].x;
            av[r].y -= tau * av[r1].y;
        }
    }
}
template <int BIT>
__device__ __forceinline__ void qr12_ip(const double2 (&l)[8], const double2 (&p)[8], const Gate12& g, double& acc) {
    const int m = g.mode;
    if (m != 0 && m != 1) return;
    double s = 0.0;
    if (m == 0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
            s += im_conj_mul(l[r], p[r1]) + im_conj_mul(l[r1], p[r]);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            if (r & (1 << BIT)) continue;
            const int r1 = r | (1 << BIT);
            s += re_conj_mul(l[r1], p[r]) - re_conj_mul(l[r], p[r1]);
        }
    }
    acc += s;
}
template <int NV, int G, int NB, int K>
__device__ __forceinline__ void qr12_round_split(double2 (&a)[NV][8], const Gate12* sg, double (&acc)[QR_SLOTS], double2* smem, int tid,
                                                 bool write) {
    constexpr int T = 1 << K;
    const int bp = qr12_sbase<G>(tid);
    qr12_shear<0>(a[0], sg[G + 0]);
    if (NB > 1) qr12_shear<1>(a[0], sg[G + (NB > 1 ? 1 : 0)]);
    if (NB > 2) qr12_shear<2>(a[0], sg[G + (NB > 2 ? 2 : 0)]);
    if (write) {
#pragma unroll
        for (int r = 0; r < 8; ++r) smem[bp ^ qr12_cr<G>(r)] = a[0][r];
    }
    if (NV == 2) {
        qr12_shear<0>(a[NV - 1], sg[G + 0]);
        if (NB > 1) qr12_shear<1>(a[NV - 1], sg[G + (NB > 1 ? 1 : 0)]);
        if (NB > 2) qr12_shear<2>(a[NV - 1], sg[G + (NB > 2 ? 2 : 0)]);
        qr12_ip<0>(a[NV - 1], a[0], sg[G + 0], acc[G + 0]);
        if (NB > 1) qr12_ip<1>(a[NV - 1], a[0], sg[G + (NB > 1 ? 1 : 0)], acc[G + (NB > 1 ? 1 : 0)]);
        if (NB > 2) qr12_ip<2>(a[NV - 1], a[0], sg[G + (NB > 2 ? 2 : 0)], acc[G + (NB > 2 ? 2 : 0)]);
        if (write) {
#pragma unroll
            for (int r = 0; r < 8; ++r) smem[T + (bp ^ qr12_cr<G>(r))] = a[NV - 1][r];
        }
    }
}
// second half of an exchange: the registers of group GN (psi first: the next round starts with psi)
template <int NV, int GN, int K>
__device__ __forceinline__ void qr12_xread(double2 (&a)[NV][8], const double2* smem, int tid) {
    constexpr int T = 1 << K;
    const int bn = qr12_sbase<GN>(tid);
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int r = 0; r < 8; ++r) a[v][r] = smem[v * T + (bn ^ qr12_cr<GN>(r))];
}

// streaming (evict-first) accesses: every amplitude is read once and written once per pass
#ifndef QR_HOST_EMUL
__device__ __forceinline__ double2 qr_ldcs(const double2* p) { return __ldcs(p); }
__device__ __forceinline__ double qr_ldcg(const double* p) { return __ldcg(p); }   // L2 only: written by other CTAs
__device__ __forceinline__ void qr_stcs(double2* p, double2 v) { __stcs(p, v); }
#else
__device__ __forceinline__ double2 qr_ldcs(const double2* p) { return *p; }
__device__ __forceinline__ double qr_ldcg(const double* p) { return *p; }
__device__ __forceinline__ void qr_stcs(double2* p, double2 v) { *p = v; }
#endif

// ---- per-thread asynchronous copies (LDGSTS): 16 B global -> shared, no register staging ----
#ifndef QR_HOST_EMUL
__device__ __forceinline__ void qr_cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void qr_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void qr_cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
#else
__device__ __forceinline__ void qr_cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
__device__ __forceinline__ void qr_cp_async_commit() {}
__device__ __forceinline__ void qr_cp_async_wait_all() {}
#endif

// ---- programmatic dependent launch (QR_OPT_PDL): the CTAs of a pass may become resident while the previous pass
// drains (its CTAs that ran out of tiles free their slots); nothing is read from or written to global memory before
// qr_pdl_wait(), which returns once the previous kernel on the stream has completed and its stores are visible.
// Both instructions are no-ops when the kernel was launched without the attribute.
#ifndef QR_HOST_EMUL
__device__ __forceinline__ void qr_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void qr_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void qr_pdl_wait() {}
__device__ __forceinline__ void qr_pdl_launch_dependents() {}
#endif

// ---- thread-block cluster helpers (pair kernel): rank, distributed-shared-memory loads, split barrier ----
#ifndef QR_HOST_EMUL
__device__ __forceinline__ unsigned qr_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned qr_map_remote(const void* smem_ptr, unsigned rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(smem_ptr), o;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(rank));
    return o;
}
__device__ __forceinline__ double2 qr_ld_remote(unsigned addr) {
    double2 v;
    asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void qr_cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void qr_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// 16-byte store into the partner CTA's shared memory that credits `bytes` to the partner's mbarrier on completion:
// no fence on the producer side, the consumer's mbarrier wait orders the data (async proxy)
__device__ __forceinline__ void qr_st_async_remote(unsigned remote_addr, double2 v, unsigned remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr), "d"(v.x), "d"(v.y),
                 "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void qr_mbar_arrive_remote(unsigned remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
#else   // the emulation runs blocks one after another: the pair kernel is never launched there
__device__ __forceinline__ void qr_st_async_remote(unsigned, double2, unsigned) {}
__device__ __forceinline__ void qr_mbar_arrive_remote(unsigned) {}
__device__ __forceinline__ unsigned qr_cluster_rank() { return 0; }
__device__ __forceinline__ unsigned qr_map_remote(const void*, unsigned) { return 0; }
__device__ __forceinline__ double2 qr_ld_remote(unsigned) { return make_double2(0.0, 0.0); }
__device__ __forceinline__ void qr_cluster_arrive() {}
__device__ __forceinline__ void qr_cluster_wait() {}
#endif

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
//
// STAGED: the CTA's NEXT tile is copied global -> shared memory by per-thread asynchronous copies
// (LDGSTS) while the current tile is computed.  Every thread copies exactly the 8 amplitudes per
// vector it will hold itself at load time into slots nobody else touches, so the stage needs no
// barrier at all: wait for the own copy group, read the slots into registers, re-issue the copies
// of the tile after that.  Shared memory: NV tile-sized stages + ONE exchange buffer (the vectors
// take turns) = 192 KiB for the backward pass; HBM reads overlap the whole gate/exchange phase
// instead of only reaching L2 (prefetch) or being waited for (direct loads).
// STAGED == 2 (backward only): only psi is staged, lambda is loaded directly (L2 prefetch) and the
// exchange keeps its two buffers and single barrier: [exchange psi][exchange lambda][stage psi].
// K = 11: half-size tiles (2048 amplitudes, 256 threads, 64 KiB of shared memory for the backward pass): two
// backward CTAs per SM, whose load / FP64 / exchange phases overlap; used where it does not cost a pass.
//
// PAIR (K = 11 only): a cluster of two such CTAs covers one 12-bit tile -- rank rho owns the half with local bit
// 11 = rho -- so a pass keeps its 12 gate bits AND two backward CTAs fit on an SM (2 x 96 KiB of shared memory,
// 2 x 256 x 128 registers).  After the gates of bits 8-10 the CTAs trade, through distributed shared memory, the
// halves of their registers that differ in bit 10: rank rho keeps bit 10 = rho and receives the partner's
// amplitudes with the other value of bit 11, which land in the register slots just vacated.  Bit 11 is then a
// register bit (gate applied), bit 10 is the rank bit for the rest of the pass (its gate is done), and only the
// store addresses notice the swap.  One cluster barrier per tile (+ one split arrive/wait pair).
template <int NV, bool PHASE, int STAGED, int K = QR_MAX_TILE_BITS, bool PAIR = false>
__global__ void __launch_bounds__(1 << (K - 3), (K == 11 ? (STAGED ? 1 : (NV == 1 ? 4 : 2)) : ((NV == 1 && !STAGED) ? 2 : 1)))
    k_tile12(const TilePass p, const Tile12X x) {
    static_assert(!PAIR || (K == 11 && STAGED == 0), "pair kernel: half-size tiles, direct loads");
    constexpr int NSV = STAGED == 1 ? NV : (STAGED == 2 ? 1 : 0);   // staged vectors
    constexpr int T = 1 << K;
    constexpr int LG = K - 3;   // first local bit of the register group held at load time
    QR_DYN_SMEM(double2, smem);
    double2* const stage = smem + (STAGED == 1 ? T : (STAGED == 2 ? NV * T : 0));   // STAGED 1: [exchange][stage psi][stage lambda]
    __shared__ Gate12 sg[QR_GATE_SLOTS];
    __shared__ double2 szr[8];                  // Z phases of the G3 register bits (times nothing else)
    __shared__ double2 szb[QR_GATE_SLOTS][2];   // per gate bit: Z phase for bit value 0 / 1 (identity if not Z)
    __shared__ int s_flags[2];                  // [0]: the pass needs its diagonal (an Rz, or an odd number of sign flips)
    __shared__ double2 lut_sm[QR_LUT_MAX];
    const int tid = threadIdx.x;
    const Geo12 geo = {p.c, p.h, p.m1, p.h2, PAIR ? 12 : K};   // PAIR: addresses follow the 12-bit tile geometry
    const unsigned rho = PAIR ? qr_cluster_rank() : 0u;            // which half of the 12-bit tile (local bit 11 at load time)
    double2* const pairbuf = smem + NV * T;                        // PAIR: [NV][4][256] amplitudes handed over by the partner CTA
    __shared__ u64 bar_full, bar_empty;                            // PAIR: hand-over landed / hand-over consumed by the partner
    unsigned remote_buf = 0, remote_full = 0, remote_empty = 0;
    if (PAIR) {
#ifndef QR_HOST_EMUL
        if (threadIdx.x == 0) {
            qr_mbar_init(&bar_full, 1);                            // my expect_tx arrival + the partner's bytes
            qr_mbar_init(&bar_empty, (1 << LG) / 32);              // one arrival per partner warp
        }
        __syncthreads();
        qr_cluster_arrive();                                       // both CTAs run and their barriers are initialised
        qr_cluster_wait();
        remote_buf = qr_map_remote(pairbuf, rho ^ 1u);
        remote_full = qr_map_remote(&bar_full, rho ^ 1u);
        remote_empty = qr_map_remote(&bar_empty, rho ^ 1u);
#endif
    }
    const u64 tmask = ((u64)1 << p.tiles_log2) - 1;
    const int ng = x.ngroups;
    // destination base index of tile t.  Ladder passes may enumerate the tiles in SOURCE order: the gather map
    // scatters consecutive destination tiles over the whole source vector (a new 2 MiB page per tile and CTA, on
    // which the L2 prefetch is far less effective); enumerating source tiles keeps the reads (and the prefetch)
    // sequential and scatters the fire-and-forget writes instead.  The map is banded towards the less significant
    // bits, so the tile bits of the inverse map depend on the tile bits of the source only.
    auto dest_base = [&](u64 t) -> u64 {
        const u64 g0 = geo12_tile(geo, t);
        return p.src_order ? (ladder_map(g0, p.iM1, p.iM2) & ~(u64)((PAIR ? 2 * T : T) - 1)) : g0;
    };

    double acc_all[QR_SLOTS];
#pragma unroll
    for (int i = 0; i < QR_SLOTS; ++i) acc_all[i] = 0.0;
    double wtot = 0.0;   // running sum of Im(conj(lambda) psi) over this thread's amplitudes

    // per-thread global offsets (local bits 0-8 at load time; the last group's thread bits at store time)
    const u64 toff_d = geo12_local(geo, (u64)tid | (PAIR ? (u64)rho << 11 : 0));          // destination index bits
    const u64 toff_s = p.ladder ? ladder_map(toff_d, p.M1, p.M2) : toff_d;               // gathered source bits
    const int tbl = ng > 1 ? qr12_tb(tid, K == 12 ? 6 : x.last_group) : tid;   // K = 12: the last group is always 6
    // PAIR: at store time local bit 10 of the half tile is tile bit 11 and the rank is tile bit 10
    const u64 toff_l = PAIR ? geo12_local(geo, (u64)(tbl & 0x3FF) | ((u64)rho << 10) | ((u64)((tbl >> 10) & 1) << 11))
                            : geo12_local(geo, (u64)tbl);

    // Uniform trip count over the grid: with x.cluster > 1 every CTA of a cluster must reach the
    // per-tile cluster barrier the same number of times (a CTA without a tile just arrives).
    // PAIR: the two CTAs of a cluster work on the same tile; the tile loop runs over clusters.
    const i64 nworkers = PAIR ? (i64)(gridDim.x >> 1) : (i64)gridDim.x;
    const i64 worker = PAIR ? (i64)(blockIdx.x >> 1) : (i64)blockIdx.x;
    const bool po = x.pair_order != 0;   // (the host sets it only for an even number of tiles per state)
    const i64 iters = po ? 2 * (((p.num_tiles >> 1) + nworkers - 1) / nworkers) : (p.num_tiles + nworkers - 1) / nworkers;
    auto tile_at = [&](i64 it) -> i64 { return po ? 2 * (worker + (it >> 1) * nworkers) + (it & 1) : worker + it * nworkers; };
    i64 cur_b = -1;
    double2 zt = make_double2(1.0, 0.0);   // thread factor of the merged diagonal (includes F)
    bool has_z = false;       // apply the diagonal zt * zr after the load
    bool has_zgate = false;   // the pass has an Rz: take the Z-gradient product w
    // convert the gate table of batch element b (block-uniform): reduced shears, merged Z diagonal, per-thread factors
    auto convert_gates = [&](i64 b) {
        __syncthreads();
        if (tid < QR_GATE_SLOTS) {
            GateP g = p.gates[b * p.gate_stride + tid];
            if (!PAIR && K < QR_GATE_SLOTS && tid >= K) g.axis = -1;
            Gate12 o;
            o.tau = 0.0; o.sig = 0.0; o.mode = -1; o.neg = 0;
            double2 z0 = make_double2(1.0, 0.0), z1 = z0;
            if (g.axis == 0 || g.axis == 1) {
                // reduce to |phi| <= pi/2 (c >= 0): R(c, s) = -R(-c, -s)
                const double cc = g.c < 0.0 ? -g.c : g.c, ss = g.c < 0.0 ? -g.s : g.s;
                o.neg = g.c < 0.0 ? 1 : 0;
                o.tau = ss / (1.0 + cc);
                o.sig = ss;
                o.mode = g.axis;
            } else if (g.axis == 2) {   // Rz: (c - i s) on bit value 0, (c + i s) on bit value 1 (state.py:168-170)
                o.mode = 4;
                z0 = make_double2(g.c, -g.s);
                z1 = make_double2(g.c, g.s);
            }
            sg[tid] = o;
            szb[tid][0] = z0;
            szb[tid][1] = z1;
        }
        __syncthreads();
        if (tid == 0) {
            int z = 0, neg = 0;
            for (int i = 0; i < QR_GATE_SLOTS; ++i) { z |= (sg[i].mode == 4); neg ^= sg[i].neg; }
            s_flags[0] = z | neg;
            s_flags[1] = neg;
        }
        if (tid < 8) {
            double2 z = make_double2(1.0, 0.0);
#pragma unroll
            for (int j = 0; j < 3; ++j) z = cmul(z, szb[LG + j][(tid >> j) & 1]);
            szr[tid] = z;
        }
        __syncthreads();
        has_z = s_flags[0] != 0;
        has_zgate = false;
#pragma unroll
        for (int j = 0; j < QR_GATE_SLOTS; ++j) has_zgate = has_zgate || sg[j].mode == 4;
        zt = make_double2(s_flags[1] ? -1.0 : 1.0, 0.0);
#pragma unroll
        for (int j = 0; j < LG; ++j) zt = cmul(zt, szb[j][(tid >> j) & 1]);
        if (PAIR) zt = cmul(zt, szb[11][rho]);
        cur_b = b;
    };

    // Programmatic dependent launch: everything up to qr_pdl_wait() may run while the previous pass is still draining.
    // That part touches no state vector and no reduction scratch -- only kernel parameters and the gate / phase tables,
    // which are written once per API call BEFORE its first pass; the host launches that first pass fully serialized
    // (launch_pass: tables_fresh), so every later pass may read them early.  The table conversion (an L2 round trip, a
    // division, three barriers: ~1 us) thus leaves the critical path of a short pass.  After the wait, let the next
    // pass's CTAs queue up behind this one (at most two grids are ever co-resident: the trigger comes after the wait).
    const bool use_lut = PHASE && p.hidx != nullptr && (p.pre_phase || p.post_phase);
    if (use_lut) {
        for (int i = tid; i < p.lut_size; i += blockDim.x) lut_sm[i] = p.lut[i];
    }
    if (tile_at(0) < p.num_tiles) convert_gates(tile_at(0) >> p.tiles_log2);
    qr_pdl_wait();
    qr_pdl_launch_dependents();

    // flush helper state: sign pattern of the thread-bit Z gates is applied when partials leave the thread
    auto finalize = [&]() {
#pragma unroll
        for (int b = 0; b < QR_GATE_SLOTS; ++b) {
            const int m = sg[b].mode;
            double v = acc_all[b];
            if (b < LG && m == 4) v = ((tid >> b) & 1) ? -wtot : wtot;
            if (PAIR && b == 11 && m == 4) v = rho ? -wtot : wtot;
            acc_all[b] = v;
        }
    };

    // pull tile `nt` into L2: one 128 B line per thread and vector
    auto prefetch_tile = [&](i64 nt, bool both) {
#ifndef QR_HOST_EMUL
        const i64 nb = nt >> p.tiles_log2;
        const u64 nbase = dest_base((u64)nt & tmask);
        const int l = tid << 3;
        const u64 d = nbase | geo12_local(geo, (u64)l | (PAIR ? (u64)rho << 11 : 0));
        const u64 s = p.ladder ? (ladder_map(d, p.M1, p.M2) ^ p.src_xor) : d;
        if (both || STAGED != 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src0 + nb * p.state_stride + s));
        if (NV == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src1 + nb * p.state_stride + s));
#endif
    };
    // issue the asynchronous copies of this thread's amplitudes of tile `tl` into its stage slots
    auto issue_stage = [&](i64 tl) {
        const i64 nb = tl >> p.tiles_log2;
        const u64 nbase = dest_base((u64)tl & tmask);
        const u64 sb = (p.ladder ? (ladder_map(nbase, p.M1, p.M2) ^ p.src_xor) : nbase) ^ toff_s;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const u64 sidx = sb ^ x.roff_first[r];
            qr_cp_async16(stage + tid + (r << LG), p.src0 + nb * p.state_stride + sidx);
            if (NSV == 2) qr_cp_async16(stage + T + tid + (r << LG), p.src1 + nb * p.state_stride + sidx);
        }
        qr_cp_async_commit();
        if (x.pair_order && !(tl & 1) && tl + 1 < p.num_tiles) prefetch_tile(tl + 1, true);   // the other halves of the same 256 B chunks
    };

    if (STAGED && tile_at(0) < p.num_tiles) {
#ifndef QR_HOST_EMUL
        if (x.cluster > 1) asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
#endif
        issue_stage(tile_at(0));
    }

    unsigned pair_count = 0;     // PAIR: hand-overs done so far (mbarrier phase parities)
    for (i64 it = 0; it < iters; ++it) {
        const i64 tile = tile_at(it);
#ifndef QR_HOST_EMUL
        // CTAs of a cluster own ADJACENT tiles (rows 128 B apart in the strided passes).  Aligning their
        // loads in time lets the DRAM controller serve both halves of a 256 B chunk from one row
        // activation: measured 4.8 -> 5.8 TB/s on the bare two-vector access pattern (scripts/membench.cu).
        if (!PAIR && x.cluster > 1 && (STAGED == 0 || it + 1 < iters)) asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
#endif
        if (tile >= p.num_tiles) continue;
        const i64 b = tile >> p.tiles_log2;
        const u64 t = (u64)tile & tmask;
        const u64 tbase = dest_base(t);
        if (b != cur_b) convert_gates(b);   // block-uniform: a persistent CTA of a batched pass moves on to the next circuit
        // batch element offset: state_stride is a multiple of 2^n, so it can be OR-ed into the index bits
        const u64 boff = (u64)b * (u64)p.state_stride;

        // ---- global -> registers (group G3; ladder gather folded into the load addresses) ----
        const u64 sbt = ((p.ladder ? (ladder_map(tbase, p.M1, p.M2) ^ p.src_xor) : tbase) ^ toff_s) | boff;
        double2 a[NV][8];
        if (STAGED) {
            qr_cp_async_wait_all();
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                a[0][r] = stage[tid + (r << LG)];
                if (NSV == 2) a[NV - 1][r] = stage[T + tid + (r << LG)];
                else if (NV == 2) a[NV - 1][r] = p.src1[sbt ^ x.roff_first[r]];
            }
            if (it + 1 < iters && tile_at(it + 1) < p.num_tiles) issue_stage(tile_at(it + 1));   // lands while this tile is computed
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const u64 s = sbt ^ x.roff_first[r];
                a[0][r] = (x.cache_hints & 2) ? qr_ldcs(p.src0 + s) : p.src0[s];
                if (NV == 2) a[NV - 1][r] = (x.cache_hints & 2) ? qr_ldcs(p.src1 + s) : p.src1[s];
            }
        }
        if (p.prefetch) {   // pull the next tile of this CTA into L2 while this one is computed
            if (po) {       // pairs: the partner of the very first tile, then both tiles of the next pair at once
                if (it == 0 && tile + 1 < p.num_tiles) prefetch_tile(tile + 1, false);
                if ((it & 1) && it + 1 < iters) {
                    const i64 nt = tile_at(it + 1);
                    if (nt + 1 < p.num_tiles) { prefetch_tile(nt, false); prefetch_tile(nt + 1, false); }
                }
            } else {
                const i64 nt = tile + nworkers * p.prefetch;
                if (nt < p.num_tiles) prefetch_tile(nt, false);
            }
        }
        // ---- Z gradients: w = Im(conj(lambda) psi), signed sums over the register bits, total for the thread bits ----
        if (NV == 2 && has_zgate) {
            double w[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) w[r] = im_conj_mul(a[NV - 1][r], a[0][r]);
            const double e0 = w[0] + w[1], o0 = w[0] - w[1], e1 = w[2] + w[3], o1 = w[2] - w[3];
            const double e2 = w[4] + w[5], o2 = w[4] - w[5], e3 = w[6] + w[7], o3 = w[6] - w[7];
            const double ee0 = e0 + e1, eo0 = e0 - e1, ee1 = e2 + e3, eo1 = e2 - e3;
            wtot += ee0 + ee1;
            if (sg[LG].mode == 4) acc_all[LG] += (o0 + o1) + (o2 + o3);
            if (sg[LG + 1].mode == 4) acc_all[LG + 1] += eo0 + eo1;
            if (sg[LG + 2].mode == 4) acc_all[LG + 2] += ee0 - ee1;
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
        }
        // ---- rounds ----
        if (QR_T12_SPLIT_XCHG && !PAIR && STAGED != 1) {
#define QR12_RS(G, NB, W) qr12_round_split<NV, G, NB, K>(a, sg, acc_all, smem, tid, W)
#define QR12_XR(GN) qr12_xread<NV, GN, K>(a, smem, tid)
            // one block barrier per exchange: a thread writes only the slots it read itself in the previous exchange
            QR12_RS(LG, 3, ng > 1);
            if (ng > 1) __syncthreads();
            if (ng == 4) {
                QR12_XR(0);
                QR12_RS(0, 3, true);
                __syncthreads();
                QR12_XR(3);
            } else if (ng == 3) {
                QR12_XR(3);
            }
            if (K == 12 ? ng >= 3 : (ng == 3 || ng == 4)) {
                QR12_RS(3, 3, true);
                __syncthreads();
                QR12_XR(6);
            } else if (ng == 2) {
                QR12_XR(6);
            }
            if (K == 12 ? ng >= 2 : (ng >= 2 && ng <= 4)) {
                QR12_RS(6, (K == 12 ? 3 : 2), false);   // K = 11: bit 8 belongs to the load group
                __syncthreads();   // every thread has read its last exchange: smem is free for the next tile
            }
            if (K == 11 && ng == 5) {   // 64 B rows: gate bits 2-10 = groups 8 | 2 | 5
                QR12_XR(2);
                QR12_RS(2, 3, true);
                __syncthreads();
                QR12_XR(5);
                QR12_RS(5, 3, false);
                __syncthreads();
            }
#undef QR12_RS
#undef QR12_XR
        } else {
#define QR12_X(GP, GN) do { if (STAGED == 1) qr12_exchange_1buf<NV, GP, GN>(a, smem, tid); else qr12_exchange<NV, GP, GN, K>(a, smem, tid); } while (0)
        qr12_round<NV, LG>(a, sg, acc_all);
        if (PAIR) {
            // Hand the register half with bit 10 != rho to the partner CTA (asynchronous stores into ITS buffer, credited
            // to ITS `full` mbarrier) and take its half with bit 10 == rho, bit 11 = 1 - rho, from my own buffer.  No
            // cluster-wide barrier or fence per tile: a release fence would wait for the previous tile's global stores.
            if (pair_count > 0) qr_mbar_wait(&bar_empty, (pair_count - 1) & 1u);   // the partner has read my previous hand-over
            if (tid == 0) qr_mbar_expect_tx(&bar_full, (unsigned)(NV * 4 * (1 << LG) * sizeof(double2)));
            if (rho == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        qr_st_async_remote(remote_buf + (unsigned)(((v * 4 + k) * (1 << LG) + tid) * sizeof(double2)), a[v][k | 4], remote_full);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < NV; ++v)
                        qr_st_async_remote(remote_buf + (unsigned)(((v * 4 + k) * (1 << LG) + tid) * sizeof(double2)), a[v][k], remote_full);
            }
            qr_mbar_wait(&bar_full, pair_count & 1u);   // the partner's half has landed in my buffer
            if (rho == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < NV; ++v) a[v][k | 4] = pairbuf[(v * 4 + k) * (1 << LG) + tid];
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int v = 0; v < NV; ++v) a[v][k] = pairbuf[(v * 4 + k) * (1 << LG) + tid];
            }
            qr12_gate<NV, 2>(a, sg[11], acc_all[11]);   // register bit 2 is tile bit 11 now
            __syncwarp();
            if ((tid & 31) == 0) qr_mbar_arrive_remote(remote_empty);   // my buffer may be overwritten (one arrival per warp)
            ++pair_count;
        }
        if (ng == 4) {
            QR12_X(LG, 0);
            qr12_round<NV, 0>(a, sg, acc_all);
            QR12_X(0, 3);
        } else if (ng == 3) {
            QR12_X(LG, 3);
        }
        if (K == 12 ? ng >= 3 : (ng == 3 || ng == 4)) {
            qr12_round<NV, 3>(a, sg, acc_all);
            QR12_X(3, 6);
        } else if (ng == 2) {
            QR12_X(LG, 6);
        }
        if (K == 12 ? ng >= 2 : (ng >= 2 && ng <= 4)) {
            qr12_round<NV, 6, (K == 12 ? 3 : 2)>(a, sg, acc_all);   // K = 11: bit 8 belongs to the load group
            __syncthreads();   // every thread has read its last exchange: smem is free for the next tile
        }
        if (K == 11 && ng == 5) {   // 64 B rows: gate bits 2-10 = groups 8 | 2 | 5
            QR12_X(LG, 2);
            qr12_round<NV, 2>(a, sg, acc_all);
            QR12_X(2, 5);
            qr12_round<NV, 5>(a, sg, acc_all);
            __syncthreads();
        }
#undef QR12_X
        }
        // ---- registers -> global (QAOA backward: diagonal-generator inner product and un-phase) ----
        const u64 dlt = tbase | toff_l | boff;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const u64 d = dlt | x.roff_last[r];
            if (PHASE && p.post_phase) {
                double hv;
                double2 ph;
                if (use_lut) {
                    const int hi = p.hidx[d ^ boff];
                    hv = p.hmin + (double)hi;
                    ph = lut_sm[hi];
                } else ph = qr12_phase_slow(p.ham, d ^ boff, p.angle_post, &hv);
                if (NV == 2) acc_all[QR_SLOTS - 1] += hv * im_conj_mul(a[NV - 1][r], a[0][r]);
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
            }
            if (x.cache_hints & 1) {
                qr_stcs(p.dst0 + d, a[0][r]);
                if (NV == 2) qr_stcs(p.dst1 + d, a[NV - 1][r]);
            } else {
                p.dst0[d] = a[0][r];
                if (NV == 2) p.dst1[d] = a[NV - 1][r];
            }
        }
        if (NV == 2 && p.flush_per_tile) {
            finalize();
            qr_block_reduce_slots(acc_all, p.partials + (u64)tile * QR_SLOTS);
#pragma unroll
            for (int i = 0; i < QR_SLOTS; ++i) acc_all[i] = 0.0;
            wtot = 0.0;
        }
    }
    if (PAIR) {   // do not exit while the partner may still store into my buffer or arrive on my barriers
        qr_cluster_arrive();
        qr_cluster_wait();
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
                // one warp per slot: lane l adds the partials of CTAs l, l+32, ... (independent L2 loads), then a
                // fixed-order shuffle tree -> deterministic for a given grid, ~5 load latencies instead of gridDim
                // dependent ones (the serial loop cost ~6 us per launch, 20 % of a 20-qubit backward pass)
                const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
                for (int base = 0; base < QR_SLOTS; base += nw) {   // block-uniform trip count
                    const int i = base + w;
                    double v = 0.0;
                    if (i < QR_SLOTS)
                        for (unsigned bb = lane; bb < gridDim.x; bb += 32) v += qr_ldcg(p.partials + (u64)bb * QR_SLOTS + i);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (i < QR_SLOTS && lane == 0) p.final_out[i] = v;
                }
                if (tid == 0) *p.done_counter = 0u;   // re-arm for the next launch on this stream
            }
        }
    }
}
