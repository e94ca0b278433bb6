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
// hole > 0 (sliced passes of sharded registers): the index bits [9, 9+hole) are not enumerated by the tile index -- the
// launch fixes them (TilePass::tile_or), so that a pass can be issued slice by slice and the exchange pass of a slice can
// start while the local pass still works on the next one.  They must lie between the rows and the first gate run.
#define QR_HOLE_POS 9
struct Geo12 { int c, h, m1, h2, k, hole; };   // k = tile bits (12, or 11 for the half-size tiles)
__host__ __device__ __forceinline__ u64 geo12_local(const Geo12 g, u64 l) {
    return (l & (((u64)1 << g.c) - 1)) | (((l >> g.c) & (((u64)1 << g.m1) - 1)) << g.h) | ((l >> (g.c + g.m1)) << g.h2);
}
__host__ __device__ __forceinline__ u64 geo12_tile(const Geo12 g, u64 t) {
    const int nlo = g.h - g.c - g.hole, nmid = g.h2 - g.h - g.m1, m2 = g.k - g.c - g.m1;
    const u64 lo = t & (((u64)1 << nlo) - 1);
    const u64 lo_placed = g.hole ? (((lo & (((u64)1 << (QR_HOLE_POS - g.c)) - 1)) << g.c) | ((lo >> (QR_HOLE_POS - g.c)) << (QR_HOLE_POS + g.hole)))
                                 : (lo << g.c);
    return lo_placed | (((t >> nlo) & (((u64)1 << nmid) - 1)) << (g.h + g.m1)) | ((t >> (nlo + nmid)) << (g.h2 + m2));
}

struct Tile12X {
    int ngroups;            // chain of register groups (first local bit of each; K = tile bits, L = K-3):
                            // 1 = L; 2 = L,6; 3 = L,3,6; 4 = L,0,3,6; 5 (K = 11, 64 B rows) = L,2,5; 6 (K = 12, GX) = L,0,3
    int last_group;         // first local bit of the register group held at store time
    u64 roff_first[8];      // global offset of register r at load time (group G3), gather map applied
    u64 roff_last[8];       // global offset of register r at store time (G2, or G3 when ngroups == 1)
    u64 droff_first[8];     // same as roff_first without the gather map (destination index: phase tables)
    // GX passes (axis-aware plans): general tile geometry -- local bit b is global index bit lpos[b] (rows: lpos[b] = b),
    // tile-index bit j is global index bit tpos[j] (the index bits outside the tile, ascending)
    unsigned char lpos[QR_MAX_TILE_BITS];
    unsigned char tpos[24];
};

// XMAP passes (sharded registers, qr_shard.cuh): the source of a tile is given by a general GF(2)-linear map instead of
// the banded ladder masks, and by a table of source POINTERS -- the shard a value is read from (this rank's or a peer's,
// over NVLink) is selected by the register index (exchange passes: the register bits at load time are the rank bits of
// the source layout) and / or by up to two bits of the tile index (ladder passes in the swapped layout).
//   source local index = tmap(tile index) ^ XOR_{thread bit b set} lcol[b] ^ roff_first[register] ^ src_const
//   source pointer     = src[vector][selector][register]; selector bit k = tile-index bit sel_bit[k] or thread bit thr_bit[k]
struct TileXMap {
    const double2* src[2][4][8];   // [vector][selector][register]
    u64 tcol[24];                  // image of tile-index bit j (source local index bits)
    u64 lcol[9];                   // image of tile-local bit b < 9 (the thread bits at load time)
    u64 src_const;                 // contribution of this rank's id
    int sel_bit[2];                // tile-index bit of selector bit k (or -1)
    int thr_bit[2];                // thread bit (tile-local bit < 9) of selector bit k (or -1)
    int local_only;                // every pointer is this rank's own buffer: the L2 prefetch of the next tile is useful
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

// Barrier of one half of the CTA (even / odd warps): an exchange between the register groups 9-11 and 6-8 only moves data
// between threads that agree in tile-local bits 0-5, i.e. in thread-id bit 5 = the lowest warp-index bit, so the even and
// the odd warps of a 12-bit tile never exchange anything in a pass with two rounds.  With their own barriers the halves
// drift apart and one loads / stores while the other computes (TilePass::split_bar).
// split = 2: the exchange between the register groups 0-2 and 3-5 stays inside aligned groups of 8 threads (a thread that
// holds local bits 3-11 = tid with bits 0-2 in registers hands over to the threads with the same tid >> 3): a warp barrier
// is all it needs.
#ifndef QR_HOST_EMUL
__device__ __forceinline__ void qr12_bar(int split, int tid) {
    if (split == 0) __syncthreads();
    else if (split == 2) __syncwarp();
    else __barrier_sync_count(1 + ((tid >> 5) & 1), 256);   // 512-thread CTAs only (K = 12)
}
#else
__device__ __forceinline__ void qr12_bar(int, int) { __syncthreads(); }
#endif

// registers of group GP -> shared memory -> registers of group GN (one block barrier)
template <int NV, int GP, int GN, int K = QR_MAX_TILE_BITS>
__device__ __forceinline__ void qr12_exchange(double2 (&a)[NV][8], double2* smem, int tid, int split = 0) {
    constexpr int T = 1 << K;
    const int bp = qr12_sbase<GP>(tid), bn = qr12_sbase<GN>(tid);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int l = bp ^ qr12_cr<GP>(r);
#pragma unroll
        for (int v = 0; v < NV; ++v) smem[v * T + l] = a[v][r];
    }
    qr12_bar(split, tid);
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
__device__ __forceinline__ void qr12_exchange_1buf(double2 (&a)[NV][8], double2* xbuf, int tid, int split = 0) {
    const int bp = qr12_sbase<GP>(tid), bn = qr12_sbase<GN>(tid);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        // safe without a barrier: this thread overwrites only the slots it read itself last time
        // (v == 0: its GP slots of the previous exchange's last vector; v > 0: needs the barrier below)
        if (v > 0) qr12_bar(split, tid);
#pragma unroll
        for (int r = 0; r < 8; ++r) xbuf[bp ^ qr12_cr<GP>(r)] = a[v][r];
        qr12_bar(split, tid);
#pragma unroll
        for (int r = 0; r < 8; ++r) a[v][r] = xbuf[bn ^ qr12_cr<GN>(r)];
    }
}

// L2-only loads of data written by other CTAs of the same grid (the fused final reduction)
#ifndef QR_HOST_EMUL
__device__ __forceinline__ double qr_ldcg(const double* p) { return __ldcg(p); }   // L2 only: written by other CTAs
#else
__device__ __forceinline__ double qr_ldcg(const double* p) { return *p; }
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
// STAGED == 1: the CTA's NEXT tile is copied global -> shared memory by per-thread asynchronous copies
// (LDGSTS) while the current tile is computed.  Every thread copies exactly the 8 amplitudes per
// vector it will hold itself at load time into slots nobody else touches, so the stage needs no
// barrier at all: wait for the own copy group, read the slots into registers, re-issue the copies
// of the tile after that.  Shared memory: NV tile-sized stages + ONE exchange buffer (the vectors
// take turns) = 192 KiB for the backward pass; HBM reads overlap the whole gate/exchange phase
// instead of only reaching L2 (prefetch) or being waited for (direct loads).
// K = 11: half-size tiles (2048 amplitudes, 256 threads, 64 KiB of shared memory for the backward pass): two
// backward CTAs per SM, whose load / FP64 / exchange phases overlap; used where it does not cost a pass.
//
// GX (axis-aware plans of single McClean circuits, qr_lib.cu: plan_axis_layer): the tile's index bits are an arbitrary
// subset (rows + any high bits, Tile12X::lpos / tpos; tile base from three byte tables in shared memory), and an Rz
// whose index bit lies OUTSIDE the tile is applied without a tile bit: its phase is a factor of the tile (selected by
// the tile's own index bit) and its gradient is +- the tile's total of w.  Such gates sit in the gate slots of the row
// bits (local bits < c, which carry no gates in a strided pass) with GateP::pad = 1 + index bit.  Diagonal gates thus
// never cost tile capacity: the strided passes of a layer only need tile bits for its X / Y rotations, which buys wider
// rows (DRAM efficiency) and fewer exchange rounds.
#define QR_GX_ZSLOTS 6    // more slots spill (the two-vector kernel sits at 128 registers)
template <int NV, bool PHASE, int STAGED, int K, bool XMAP, bool GX = false, bool SPLIT = false>
__device__ __forceinline__ void qr12_body(const TilePass& p, const Tile12X& x, const TileXMap* xmp) {
    static_assert(!XMAP || (STAGED == 0 && K == 12), "XMAP passes: direct loads, 12-bit tiles");
    static_assert(!GX || (!XMAP && !PHASE), "GX passes: McClean layers of one register");
    static_assert(!SPLIT || (GX && K == 12), "split barriers: two-round passes of 12-bit tiles");
    constexpr int T = 1 << K;
    constexpr int LG = K - 3;   // first local bit of the register group held at load time
    QR_DYN_SMEM(double2, smem);
    double2* const stage = smem + (STAGED == 1 ? T : 0);   // STAGED 1: [exchange][stage psi][stage lambda]
    __shared__ Gate12 sg[QR_GATE_SLOTS];
    __shared__ double2 szr[8];                  // Z phases of the G3 register bits (times nothing else)
    __shared__ double2 szb[QR_GATE_SLOTS][2];   // per gate bit: Z phase for bit value 0 / 1 (identity if not Z)
    __shared__ int s_flags[2];                  // [0]: the pass needs its diagonal (an Rz, or an odd number of sign flips)
    __shared__ double2 lut_sm[PHASE ? QR_LUT_MAX : 1];
    __shared__ u64 gx_tab[GX ? 3 : 1][GX ? 256 : 1];   // GX: tile-index byte -> global index bits
    __shared__ int s_zq[GX ? QR_GX_ZSLOTS : 1];        // GX: index bit of the out-of-tile Rz in slot j, or -1
    const int tid = threadIdx.x;
    const Geo12 geo = {p.c, p.h, p.m1, p.h2, K, p.hole};
    if (GX) {
        for (int i = tid; i < 3 * 256; i += blockDim.x) {
            const int j = i >> 8, v = i & 255;
            u64 m = 0;
            for (int b = 0; b < 8; ++b)
                if ((v >> b) & 1) m |= (u64)1 << x.tpos[8 * j + b];
            gx_tab[j][v] = m;
        }
        __syncthreads();
    }
    // global index bits of a tile-local index / of a tile index
    auto local_bits = [&](u64 l) -> u64 {
        if (!GX) return geo12_local(geo, l);
        u64 m = 0;
#pragma unroll
        for (int b = 0; b < K; ++b)
            if ((l >> b) & 1) m |= (u64)1 << x.lpos[b];
        return m;
    };
    auto tile_bits = [&](u64 t) -> u64 {
        if (GX) return gx_tab[0][t & 255] | gx_tab[1][(t >> 8) & 255] | gx_tab[2][(t >> 16) & 255];
        return geo12_tile(geo, t) | p.tile_or;
    };
    const u64 tmask = ((u64)1 << p.tiles_log2) - 1;
    const int ng = x.ngroups;

    double acc_all[QR_SLOTS];
#pragma unroll
    for (int i = 0; i < QR_SLOTS; ++i) acc_all[i] = 0.0;
    double wtot = 0.0;   // running sum of Im(conj(lambda) psi) over this thread's amplitudes

    // per-thread global offsets (local bits 0-8 at load time; the last group's thread bits at store time)
    const u64 toff_d = local_bits((u64)tid);                                  // destination index bits
    u64 toff_s = p.ladder ? ladder_map(toff_d, p.M1, p.M2) : toff_d;                // gathered source bits
    __shared__ u64 xm_tab[XMAP ? 3 : 1][XMAP ? 256 : 1];                            // XMAP: tile index byte -> source index bits
    if (XMAP) {
        toff_s = xmp->src_const;
#pragma unroll
        for (int b = 0; b < LG; ++b)
            if ((tid >> b) & 1) toff_s ^= xmp->lcol[b];
        for (int i = tid; i < 3 * 256; i += blockDim.x) {
            const int j = i >> 8, v = i & 255;
            u64 m = 0;
            for (int b = 0; b < 8; ++b)
                if ((v >> b) & 1) m ^= xmp->tcol[8 * j + b];
            xm_tab[j][v] = m;
        }
        __syncthreads();
    }
    // XMAP: source index bits and pointer selector of tile t
    auto xm_base = [&](u64 t) -> u64 { return xm_tab[0][t & 255] ^ xm_tab[1][(t >> 8) & 255] ^ xm_tab[2][(t >> 16) & 255]; };
    int xm_thr_sel = 0;   // selector bits taken from this thread's id (fixed for the whole kernel)
    if (XMAP) {
        if (xmp->thr_bit[0] >= 0) xm_thr_sel |= (tid >> xmp->thr_bit[0]) & 1;
        if (xmp->thr_bit[1] >= 0) xm_thr_sel |= ((tid >> xmp->thr_bit[1]) & 1) << 1;
    }
    auto xm_sel = [&](u64 t) -> int {
        int s = xm_thr_sel;
        if (xmp->sel_bit[0] >= 0) s |= (int)((t >> xmp->sel_bit[0]) & 1);
        if (xmp->sel_bit[1] >= 0) s |= (int)((t >> xmp->sel_bit[1]) & 1) << 1;
        return s;
    };
    const int tbl = ng > 1 ? qr12_tb(tid, K == 12 ? (GX && ng == 6 ? 3 : 6) : x.last_group) : tid;   // K = 12: the last group is 6 (3 for the chain L | 0 | 3)
    const u64 toff_l = local_bits((u64)tbl);

    const i64 nworkers = (i64)gridDim.x, worker = (i64)blockIdx.x;
    const int iters = (int)((p.num_tiles + nworkers - 1) / nworkers);   // tiles per CTA (32-bit loop state: the kernel sits at its register budget)
    // GX passes are never batched: tile indices fit 32 bits (n - K <= 24)
    auto tile_at = [&](int it) -> i64 { return GX ? (i64)((int)blockIdx.x + it * (int)gridDim.x) : worker + (i64)it * nworkers; };
    int cur_b = -1;
    double2 zt = make_double2(1.0, 0.0);   // thread factor of the merged diagonal (includes F)
    bool has_z = false;       // apply the diagonal zt * zr after the load
    bool has_zgate = false;   // the pass has an Rz: take the Z-gradient product w
    // convert the gate table of batch element b (block-uniform): reduced shears, merged Z diagonal, per-thread factors
    auto convert_gates = [&](int b) {
        __syncthreads();
        if (tid < QR_GATE_SLOTS) {
            GateP g = p.gates[(i64)b * p.gate_stride + tid];
            if (K < QR_GATE_SLOTS && tid >= K) g.axis = -1;
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
                o.mode = (GX && g.pad > 0) ? 5 : 4;   // 5: the index bit lies outside the tile (GX)
                z0 = make_double2(g.c, -g.s);
                z1 = make_double2(g.c, g.s);
            }
            if (GX && tid < QR_GX_ZSLOTS) s_zq[tid] = o.mode == 5 ? g.pad - 1 : -1;
            sg[tid] = o;
            szb[tid][0] = z0;
            szb[tid][1] = z1;
        }
        __syncthreads();
        if (tid == 0) {
            int z = 0, neg = 0;
            for (int i = 0; i < QR_GATE_SLOTS; ++i) { z |= (sg[i].mode >= 4); neg ^= sg[i].neg; }
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
        for (int j = 0; j < QR_GATE_SLOTS; ++j) has_zgate = has_zgate || sg[j].mode >= 4;
        zt = make_double2(s_flags[1] ? -1.0 : 1.0, 0.0);
#pragma unroll
        for (int j = 0; j < LG; ++j)
            if (!GX || sg[j].mode != 5) zt = cmul(zt, szb[j][(tid >> j) & 1]);
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
    if (tile_at(0) < p.num_tiles) convert_gates((int)(tile_at(0) >> p.tiles_log2));
    qr_pdl_wait();
    qr_pdl_launch_dependents();

    // flush helper state: sign pattern of the thread-bit Z gates is applied when partials leave the thread
    auto finalize = [&]() {
#pragma unroll
        for (int b = 0; b < QR_GATE_SLOTS; ++b) {
            const int m = sg[b].mode;
            double v = acc_all[b];
            if (b < LG && m == 4) v = ((tid >> b) & 1) ? -wtot : wtot;
            acc_all[b] = v;
        }
    };

    // pull tile `nt` into L2: one 128 B line per thread and vector
    auto prefetch_tile = [&](i64 nt) {
#ifndef QR_HOST_EMUL
        const i64 nb = GX ? 0 : (nt >> p.tiles_log2);   // GX passes are never batched
        if (XMAP) {   // line tid of the tile: tile-local bits 3..11 = tid bits 0..8 (bits 9-11 are the load-time register bits)
            const u64 t2 = (u64)nt & tmask;
            u64 s = xm_base(t2) ^ xmp->src_const ^ x.roff_first[tid >> 6];
#pragma unroll
            for (int b = 3; b < LG; ++b)
                if ((tid >> (b - 3)) & 1) s ^= xmp->lcol[b];
            const int sl = xm_sel(t2);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(xmp->src[0][sl][tid >> 6] + s));
            if (NV == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(xmp->src[1][sl][tid >> 6] + s));
            return;
        }
        const u64 nbase = tile_bits((u64)nt & tmask);
        const int l = tid << 3;
        const u64 d = nbase | local_bits((u64)l);
        const u64 s = p.ladder ? (ladder_map(d, p.M1, p.M2) ^ p.src_xor) : d;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src0 + nb * p.state_stride + s));
        if (NV == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.src1 + nb * p.state_stride + s));
#endif
    };
    // issue the asynchronous copies of this thread's amplitudes of tile `tl` into its stage slots
    auto issue_stage = [&](i64 tl) {
        const i64 nb = GX ? 0 : (tl >> p.tiles_log2);
        const u64 nbase = tile_bits((u64)tl & tmask);
        const u64 sb = (p.ladder ? (ladder_map(nbase, p.M1, p.M2) ^ p.src_xor) : nbase) ^ toff_s;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const u64 sidx = sb ^ x.roff_first[r];
            qr_cp_async16(stage + tid + (r << LG), p.src0 + nb * p.state_stride + sidx);
            if (NV == 2) qr_cp_async16(stage + T + tid + (r << LG), p.src1 + nb * p.state_stride + sidx);
        }
        qr_cp_async_commit();
    };

    if (STAGED && tile_at(0) < p.num_tiles) issue_stage(tile_at(0));

    for (int it = 0; it < iters; ++it) {
        const i64 tile = tile_at(it);
        if (tile >= p.num_tiles) continue;
        const int b = GX ? 0 : (int)(tile >> p.tiles_log2);
        const u64 t = GX ? (u64)tile : ((u64)tile & tmask);
        const u64 tbase = tile_bits(t);
        if (b != cur_b) convert_gates(b);   // block-uniform: a persistent CTA of a batched pass moves on to the next circuit
        // batch element offset: state_stride is a multiple of 2^n, so it can be OR-ed into the index bits
        const u64 boff = GX ? (u64)0 : (u64)b * (u64)p.state_stride;

        // ---- global -> registers (group G3; ladder gather folded into the load addresses) ----
        const u64 sbt = XMAP ? (xm_base(t) ^ toff_s) : (((p.ladder ? (ladder_map(tbase, p.M1, p.M2) ^ p.src_xor) : tbase) ^ toff_s) | boff);
        double2 a[NV][8];
        if (STAGED) {
            qr_cp_async_wait_all();
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                a[0][r] = stage[tid + (r << LG)];
                if (NV == 2) a[NV - 1][r] = stage[T + tid + (r << LG)];
            }
            if (it + 1 < iters && tile_at(it + 1) < p.num_tiles) issue_stage(tile_at(it + 1));   // lands while this tile is computed
        } else if (XMAP) {
            const int sl = xm_sel(t);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const u64 s = sbt ^ x.roff_first[r];
                a[0][r] = xmp->src[0][sl][r][s];
                if (NV == 2) a[NV - 1][r] = xmp->src[1][sl][r][s];
            }
        } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const u64 s = sbt ^ x.roff_first[r];
                a[0][r] = p.src0[s];
                if (NV == 2) a[NV - 1][r] = p.src1[s];
            }
        }
        if (p.prefetch) {   // pull the next tile of this CTA into L2 while this one is computed
            const i64 nt = tile + nworkers * p.prefetch;
            if (nt < p.num_tiles) prefetch_tile(nt);
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
            if (GX) {   // Rz on an index bit outside the tile: the sign is the tile's
                const double wt = ee0 + ee1;
#pragma unroll
                for (int j = 0; j < QR_GX_ZSLOTS; ++j) {
                    const int q = s_zq[j];
                    if (q >= 0) acc_all[j] += ((tbase >> q) & 1) ? -wt : wt;
                }
            }
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
            double2 ztt = zt;
            if (GX) {   // phases of the out-of-tile Rz gates: one factor per tile
#pragma unroll
                for (int j = 0; j < QR_GX_ZSLOTS; ++j) {
                    const int q = s_zq[j];
                    if (q >= 0) ztt = cmul(ztt, szb[j][(tbase >> q) & 1]);
                }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const double2 ph = cmul(ztt, szr[r]);
#pragma unroll
                for (int v = 0; v < NV; ++v) a[v][r] = cmul(a[v][r], ph);
            }
        }
        // ---- rounds ----
#define QR12_X(GP, GN) do { if (STAGED == 1) qr12_exchange_1buf<NV, GP, GN>(a, smem, tid); else qr12_exchange<NV, GP, GN, K>(a, smem, tid); } while (0)
        qr12_round<NV, LG>(a, sg, acc_all);
        if (GX && K == 12 && ng == 6) {
            // chain L | 0 | 3 (axis-aware contiguous pass whose local bits 6-8 carry no X / Y gate): two exchanges instead of
            // three, the second one warp local; stores from the group-3 layout (a warp writes four 128 B segments)
            QR12_X(LG, 0);
            qr12_round<NV, 0>(a, sg, acc_all);
            if (STAGED == 1) qr12_exchange_1buf<NV, 0, 3>(a, smem, tid, 2); else qr12_exchange<NV, 0, 3, K>(a, smem, tid, 2);
            qr12_round<NV, 3>(a, sg, acc_all);
            __syncthreads();   // every thread has read its last exchange: smem is free for the next tile
        } else
        if (ng == 4) {
            QR12_X(LG, 0);
            qr12_round<NV, 0>(a, sg, acc_all);
            if (STAGED == 1) qr12_exchange_1buf<NV, 0, 3>(a, smem, tid, 2); else qr12_exchange<NV, 0, 3, K>(a, smem, tid, 2);
        } else if (ng == 3) {
            QR12_X(LG, 3);
        }
        if (K == 12 ? (ng >= 3 && ng <= 4) : (ng == 3 || ng == 4)) {
            qr12_round<NV, 3>(a, sg, acc_all);
            QR12_X(3, 6);
        } else if (ng == 2) {
            if (SPLIT) {   // the halves of the CTA run apart
                if (STAGED == 1) qr12_exchange_1buf<NV, LG, 6>(a, smem, tid, 1); else qr12_exchange<NV, LG, 6, K>(a, smem, tid, 1);
            } else QR12_X(LG, 6);
        }
        if (ng >= 2 && ng <= 4) {
            qr12_round<NV, 6, (K == 12 ? 3 : 2)>(a, sg, acc_all);   // K = 11: bit 8 belongs to the load group
            if (SPLIT) qr12_bar(1, tid);
            else __syncthreads();   // every thread has read its last exchange: smem is free for the next tile
        }
        if (K == 11 && ng == 5) {   // 64 B rows: gate bits 2-10 = groups 8 | 2 | 5
            QR12_X(LG, 2);
            qr12_round<NV, 2>(a, sg, acc_all);
            QR12_X(2, 5);
            qr12_round<NV, 5>(a, sg, acc_all);
            __syncthreads();
        }
#undef QR12_X
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
            p.dst0[d] = a[0][r];
            if (NV == 2) p.dst1[d] = a[NV - 1][r];
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

template <int NV, bool PHASE, int STAGED, int K = QR_MAX_TILE_BITS>
__global__ void __launch_bounds__(1 << (K - 3), (K == 11 ? (STAGED ? 1 : (NV == 1 ? 4 : 2)) : ((NV == 1 && !STAGED) ? 2 : 1)))
    k_tile12(const TilePass p, const Tile12X x) {
    qr12_body<NV, PHASE, STAGED, K, false>(p, x, nullptr);
}

// axis-aware plans (GX): general tile geometry, out-of-tile Rz gates
template <int NV, int STAGED, int K = QR_MAX_TILE_BITS>
__global__ void __launch_bounds__(1 << (K - 3), (K == 11 ? (STAGED ? 1 : (NV == 1 ? 4 : 2)) : ((NV == 1 && !STAGED) ? 2 : 1)))
    k_tile12_g(const TilePass p, const Tile12X x) {
    qr12_body<NV, false, STAGED, K, false, true>(p, x, nullptr);
}
// two-round passes (one exchange, between the register groups 9-11 and 6-8): the even and the odd warps synchronise separately
template <int NV>
__global__ void __launch_bounds__(512, (NV == 1 ? 2 : 1)) k_tile12_gs(const TilePass p, const Tile12X x) {
    qr12_body<NV, false, 0, 12, false, true, true>(p, x, nullptr);
}

// sharded registers: general source map and source pointer table (exchange passes read the peers' shards over NVLink)
template <int NV, bool PHASE = false>
__global__ void __launch_bounds__(512, (NV == 1 ? 2 : 1)) k_tile12_x(const TilePass p, const Tile12X x, const TileXMap xm) {
    qr12_body<NV, PHASE, 0, 12, true>(p, x, &xm);
}
