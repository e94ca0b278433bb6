// Gate-at-a-time kernels, reductions and the sampling kernels (complex128 state vector).
//
// These implement the `State` / `Observable` method surface one call at a time
// (physical_components/state.py, observable.py) and are the un-fused cross-check for the
// fused tile passes in qr_tile.cuh.  All kernels are grid-stride, one 16-byte amplitude per
// load/store instruction (LDG.128/STG.128), coalesced along the fastest-varying index bit.
#pragma once
#include "qr_platform.cuh"

#define QR_BLOCK 256

struct ObsTerm {
    int kind;     // QR_TERM_X/Y/Z/ZZ
    int bit_i;    // index-bit position of qubit i  (n-1-i)
    int bit_j;    // index-bit position of qubit j  (ZZ only)
    int pad;
    double w;
};

// ------------------------------------------------------------------------------------------
// block-wide sum; result valid in thread 0.  blockDim.x is a power of two.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_reduce_sum(double v) {
    __shared__ double red[32];
    const int tid = threadIdx.x;
    if (blockDim.x >= 32) {
        const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[w] = v;
        __syncthreads();
        v = (lane < nw) ? red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
    } else {
        red[tid] = v;
        __syncthreads();
        if (tid == 0) { v = 0.0; for (unsigned t = 0; t < blockDim.x; ++t) v += red[t]; }
        __syncthreads();
    }
    return v;
}

// ------------------------------------------------------------------------------------------
// State.reset (state.py:61-71) and the McClean input layer prod_q Ry_q(pi/4)|0..0>
// ------------------------------------------------------------------------------------------
__global__ void k_init_basis(double2* __restrict__ v, u64 N, int which, double amp) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x) {
        double re = which == 0 ? (j == 0 ? 1.0 : 0.0) : amp;
        v[j] = make_double2(re, 0.0);
    }
}

// amplitude_j = table[popcount(j)]: product state of identical single-qubit states (mc_clean.py:35-36)
__global__ void k_init_product(double2* __restrict__ v, u64 N, const double* __restrict__ table, u64 state_mask) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x)
        v[j] = make_double2(table[__popcll(j & state_mask)], 0.0);
}

// ------------------------------------------------------------------------------------------
// general single-qubit matrix on index bit `bit`:  (a,b) -> (m00 a + m01 b, m10 a + m11 b)
// used for xrot/yrot/zrot and dxrot/dyrot/dzrot (state.py:90-97,142-149,168-175)
// ------------------------------------------------------------------------------------------
struct Mat2 { double2 m00, m01, m10, m11; };

__global__ void k_apply_1q(double2* __restrict__ v, u64 npairs, int bit, Mat2 m) {
    const u64 lowmask = ((u64)1 << bit) - 1;
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (u64)gridDim.x * blockDim.x) {
        const u64 i0 = ((p & ~lowmask) << 1) | (p & lowmask);
        const u64 i1 = i0 | ((u64)1 << bit);
        const double2 a = v[i0], b = v[i1];
        v[i0] = cadd(cmul(m.m00, a), cmul(m.m01, b));
        v[i1] = cadd(cmul(m.m10, a), cmul(m.m11, b));
    }
}

// CNOT (state.py:336-356): swap target pair where the control bit is set
__global__ void k_cnot(double2* __restrict__ v, u64 nquads, int cbit, int tbit) {
    const int lo = cbit < tbit ? cbit : tbit, hi = cbit < tbit ? tbit : cbit;
    const u64 m_lo = ((u64)1 << lo) - 1, m_hi = ((u64)1 << (hi - 1)) - 1;
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < nquads; p += (u64)gridDim.x * blockDim.x) {
        // insert zero bits at positions lo and hi
        u64 x = ((p & ~m_lo) << 1) | (p & m_lo);
        x = ((x & ~((m_hi << 1) | 1)) << 1) | (x & ((m_hi << 1) | 1));
        const u64 i0 = x | ((u64)1 << cbit);
        const u64 i1 = i0 | ((u64)1 << tbit);
        const double2 a = v[i0], b = v[i1];
        v[i0] = b;
        v[i1] = a;
    }
}

// CNOT ladder as one gather permutation (state.py:229-251): dst[i] = src[map(i)]
__global__ void k_ladder_gather(const double2* __restrict__ src, double2* __restrict__ dst, u64 N, u64 m1, u64 m2) {
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (u64)gridDim.x * blockDim.x)
        dst[i] = src[ladder_map(i, m1, m2)];
}

// x_summed (state.py:107-123): dst_j = sum_q (-i/2) src_{j ^ bit_q}
__global__ void k_x_summed(const double2* __restrict__ src, double2* __restrict__ dst, u64 N, int n) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x) {
        double sx = 0.0, sy = 0.0;
        for (int b = 0; b < n; ++b) {
            const double2 p = src[j ^ ((u64)1 << b)];
            sx += p.x; sy += p.y;
        }
        dst[j] = make_double2(0.5 * sy, -0.5 * sx);   // (-i/2)(sx + i sy)
    }
}

// ------------------------------------------------------------------------------------------
// diagonal Hamiltonian (state.py:261-321)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double diag_value(u64 j, const ObsTerm* __restrict__ terms, int nterms) {
    double d = 0.0;
    for (int k = 0; k < nterms; ++k) {
        const ObsTerm t = terms[k];
        if (t.kind == 2) d += ((j >> t.bit_i) & 1) ? -t.w : t.w;
        else if (t.kind == 3) d += (((j >> t.bit_i) ^ (j >> t.bit_j)) & 1) ? -t.w : t.w;
    }
    return d;
}

// idx_off: index bits above the local range (sharded registers: the logical value of the rank-held bits)
__global__ void k_ham_build(double* __restrict__ ham, u64 N, const ObsTerm* __restrict__ terms, int nterms, u64 idx_off) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x)
        ham[j] = diag_value(idx_off | j, terms, nterms);
}

// vec *= exp(-i angle H)   (state.py:299-301)
__global__ void k_exp_ham(double2* __restrict__ v, const double* __restrict__ ham, u64 N, double angle) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x) {
        double s, c;
        sincos(angle * ham[j], &s, &c);
        v[j] = cmul(v[j], make_double2(c, -s));
    }
}

// mode 0: vec *= -i H (state.py:319-321);  mode 1: vec *= H (qaoa.py:56)
__global__ void k_mul_ham(double2* __restrict__ v, const double* __restrict__ ham, u64 N, int mode) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x) {
        const double h = ham[j];
        const double2 a = v[j];
        v[j] = mode == 0 ? make_double2(h * a.y, -h * a.x) : make_double2(h * a.x, h * a.y);
    }
}

// ------------------------------------------------------------------------------------------
// observable application (mc_clean.py:65) fused with Re<src|O src> (mc_clean.py:66)
//   dst = O src ; partial[block] = sum_j Re(conj(src_j) dst_j)
// dst may be null (expectation only).
// ------------------------------------------------------------------------------------------
struct PeerTable { const double2* p[16]; };   // shard base pointers by logical shard id (sharded states)

// `jl` is the local index; j = idx_off | jl the full amplitude index (idx_off = shard id << nl).
// X / Y terms on a rank bit (bit >= nl) read the partner amplitude from the peer shard.
__global__ void k_apply_obs(const double2* __restrict__ src, double2* __restrict__ dst, u64 N,
                            const ObsTerm* __restrict__ terms, int nterms, double* __restrict__ partial,
                            u64 idx_off, int nl, PeerTable peers) {
    // gridDim.y = number of independent states laid out back to back (batched circuits)
    src += (u64)blockIdx.y * N;
    if (dst) dst += (u64)blockIdx.y * N;
    partial += (u64)blockIdx.y * gridDim.x;
    double acc = 0.0;
    for (u64 jl = (u64)blockIdx.x * blockDim.x + threadIdx.x; jl < N; jl += (u64)gridDim.x * blockDim.x) {
        const u64 j = idx_off | jl;
        const double2 a = src[jl];
        double d = 0.0, ox = 0.0, oy = 0.0;
        for (int k = 0; k < nterms; ++k) {
            const ObsTerm t = terms[k];
            if (t.kind == 2) d += ((j >> t.bit_i) & 1) ? -t.w : t.w;
            else if (t.kind == 3) d += (((j >> t.bit_i) ^ (j >> t.bit_j)) & 1) ? -t.w : t.w;
            else {
                const double2 p = t.bit_i < nl ? src[jl ^ ((u64)1 << t.bit_i)]
                                               : peers.p[(j >> nl) ^ ((u64)1 << (t.bit_i - nl))][jl];
                if (t.kind == 0) { ox += t.w * p.x; oy += t.w * p.y; }
                else {  // Y = [[0,-i],[i,0]]: bit 0 row gets -i p, bit 1 row gets +i p
                    const double sgn = ((j >> t.bit_i) & 1) ? 1.0 : -1.0;
                    ox += -sgn * t.w * p.y;   // (i sgn w)(p.x + i p.y) = sgn w (-p.y + i p.x)
                    oy += sgn * t.w * p.x;
                }
            }
        }
        const double2 o = make_double2(d * a.x + ox, d * a.y + oy);
        if (dst) dst[jl] = o;
        acc += re_conj_mul(a, o);
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// per-term expectations <P_k> for up to 8 terms per launch (base.py:22-33 probabilities)
#define QR_TERMS_PER_LAUNCH 8
__global__ void k_term_expecs(const double2* __restrict__ v, u64 N, const ObsTerm* __restrict__ terms, int nterms,
                              double* __restrict__ partial /*[grid][8]*/) {
    double acc[QR_TERMS_PER_LAUNCH];
#pragma unroll
    for (int k = 0; k < QR_TERMS_PER_LAUNCH; ++k) acc[k] = 0.0;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x) {
        const double2 a = v[j];
        const double p2 = a.x * a.x + a.y * a.y;
#pragma unroll
        for (int k = 0; k < QR_TERMS_PER_LAUNCH; ++k) {
            if (k < nterms) {
                const ObsTerm t = terms[k];
                if (t.kind == 2) acc[k] += ((j >> t.bit_i) & 1) ? -p2 : p2;
                else if (t.kind == 3) acc[k] += (((j >> t.bit_i) ^ (j >> t.bit_j)) & 1) ? -p2 : p2;
                else {
                    const double2 p = v[j ^ ((u64)1 << t.bit_i)];
                    if (t.kind == 0) acc[k] += re_conj_mul(a, p);
                    else acc[k] += (((j >> t.bit_i) & 1) ? -1.0 : 1.0) * im_conj_mul(a, p);  // Re(conj(a)(-+i p))
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < QR_TERMS_PER_LAUNCH; ++k) {
        const double s = block_reduce_sum(acc[k]);
        if (threadIdx.x == 0) partial[(u64)blockIdx.x * QR_TERMS_PER_LAUNCH + k] = s;
    }
}

// ||v||^2 partials
__global__ void k_norm2(const double2* __restrict__ v, u64 N, double* __restrict__ partial) {
    double acc = 0.0;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x) {
        const double2 a = v[j];
        acc += a.x * a.x + a.y * a.y;
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// QAOA lambda = H psi and E = <psi|H|psi> (qaoa.py:56-57)
__global__ void k_ham_costate(const double2* __restrict__ psi, double2* __restrict__ lam, const double* __restrict__ ham,
                              u64 N, double* __restrict__ partial) {
    double acc = 0.0;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x) {
        const double2 a = psi[j];
        const double h = ham[j];
        if (lam) lam[j] = make_double2(h * a.x, h * a.y);
        acc += h * (a.x * a.x + a.y * a.y);
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// Im<lam| P |psi> for one Pauli on one index bit (un-fused gradient path, mc_clean.py:75)
__global__ void k_pauli_inner(const double2* __restrict__ psi, const double2* __restrict__ lam, u64 npairs, int bit,
                              int axis, double* __restrict__ partial) {
    const u64 lowmask = ((u64)1 << bit) - 1;
    double acc = 0.0;
    for (u64 p = (u64)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (u64)gridDim.x * blockDim.x) {
        const u64 i0 = ((p & ~lowmask) << 1) | (p & lowmask);
        const u64 i1 = i0 | ((u64)1 << bit);
        const double2 p0 = psi[i0], p1 = psi[i1], l0 = lam[i0], l1 = lam[i1];
        if (axis == 0) acc += im_conj_mul(l0, p1) + im_conj_mul(l1, p0);
        else if (axis == 1) acc += re_conj_mul(l1, p0) - re_conj_mul(l0, p1);
        else acc += im_conj_mul(l0, p0) - im_conj_mul(l1, p1);
    }
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// sum_j H_j Im(conj(lam_j) psi_j)  (un-fused QAOA gamma gradient, qaoa.py:67-68)
__global__ void k_ham_inner(const double2* __restrict__ psi, const double2* __restrict__ lam,
                            const double* __restrict__ ham, u64 N, double* __restrict__ partial) {
    double acc = 0.0;
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x)
        acc += ham[j] * im_conj_mul(lam[j], psi[j]);
    acc = block_reduce_sum(acc);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// deterministic second stage: out[v] = sum_u partial[u*nvals + v], one block, fixed order
__global__ void k_reduce_partials(const double* __restrict__ partial, int nunits, int nvals, double* __restrict__ out) {
    for (int v = 0; v < nvals; ++v) {
        double acc = 0.0;
        for (int u = threadIdx.x; u < nunits; u += blockDim.x) acc += partial[(u64)u * nvals + v];
        acc = block_reduce_sum(acc);
        if (threadIdx.x == 0) out[v] = acc;
    }
}

// grouped variant for batches: out[g*out_stride + v] = sum over the units of group g
__global__ void k_reduce_partials_grouped(const double* __restrict__ partial, int units_per_group, int nvals,
                                          double* __restrict__ out, int out_stride) {
    const u64 base = (u64)blockIdx.x * units_per_group * nvals;
    for (int v = 0; v < nvals; ++v) {
        double acc = 0.0;
        for (int u = threadIdx.x; u < units_per_group; u += blockDim.x) acc += partial[base + (u64)u * nvals + v];
        acc = block_reduce_sum(acc);
        if (threadIdx.x == 0) out[(u64)blockIdx.x * out_stride + v] = acc;
    }
}

// ------------------------------------------------------------------------------------------
// finite-shot bitstring sampling (qaoa.py:196-198): inverse CDF of |psi|^2.
//   stage 1: chunk sums of |psi|^2 (chunk = QR_SCAN_CHUNK amplitudes, one block each)
//   stage 2: exclusive scan of the chunk sums by one block (sequential per thread + block scan)
//   stage 3: one block per shot: binary-search the chunk, then scan inside the chunk
// "first k with cdf[k] >= u" is evaluated on prefix sums accumulated in index order inside a
// chunk and chunk-wise across chunks, i.e. the same left-to-right order as numpy.cumsum up to
// the association of the partial sums.
// ------------------------------------------------------------------------------------------
#define QR_SCAN_CHUNK 4096

__global__ void k_prob_chunk_sums(const double2* __restrict__ v, u64 N, double* __restrict__ sums) {
    for (u64 chunk = blockIdx.x; chunk * QR_SCAN_CHUNK < N; chunk += gridDim.x) {
        const u64 base = chunk * QR_SCAN_CHUNK;
        double acc = 0.0;
        for (u64 j = base + threadIdx.x; j < base + QR_SCAN_CHUNK && j < N; j += blockDim.x) {
            const double2 a = v[j];
            acc += a.x * a.x + a.y * a.y;
        }
        acc = block_reduce_sum(acc);
        if (threadIdx.x == 0) sums[chunk] = acc;
    }
}

// in-place inclusive scan of `sums[0..n)` by a single block (n up to a few million)
__global__ void k_scan_inclusive_single(double* __restrict__ sums, u64 n) {
    __shared__ double tot[1024];
    const u64 per = (n + blockDim.x - 1) / blockDim.x;
    const u64 lo = (u64)threadIdx.x * per, hi = (lo + per < n) ? lo + per : n;
    double acc = 0.0;
    for (u64 i = lo; i < hi; ++i) acc += sums[i];
    tot[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = 0.0;
        for (unsigned t = 0; t < blockDim.x; ++t) { const double x = tot[t]; tot[t] = run; run += x; }
    }
    __syncthreads();
    double run = tot[threadIdx.x];
    for (u64 i = lo; i < hi; ++i) { run += sums[i]; sums[i] = run; }
}

__global__ void k_sample_search(const double2* __restrict__ v, u64 N, const double* __restrict__ chunk_cdf,
                                u64 nchunks, const double* __restrict__ uniforms, i64* __restrict__ out) {
    __shared__ double tot[QR_BLOCK];
    __shared__ i64 hits[QR_BLOCK];
    const int shot = blockIdx.x;
    const double u = uniforms[shot];
    // chunk = first c with chunk_cdf[c] >= u
    u64 lo = 0, hi = nchunks;   // search in [lo, hi)
    while (lo < hi) {
        const u64 mid = (lo + hi) >> 1;
        if (chunk_cdf[mid] >= u) hi = mid; else lo = mid + 1;
    }
    if (lo >= nchunks) {   // u above the last cdf value: scipy's argmax over all-False returns 0
        if (threadIdx.x == 0) out[shot] = 0;
        return;
    }
    const u64 chunk = lo;
    const double before = chunk ? chunk_cdf[chunk - 1] : 0.0;
    const u64 base = chunk * QR_SCAN_CHUNK;
    const int per = QR_SCAN_CHUNK / QR_BLOCK;
    double p[QR_SCAN_CHUNK / QR_BLOCK];
    double acc = 0.0;
    for (int i = 0; i < per; ++i) {
        const u64 j = base + (u64)threadIdx.x * per + i;
        double q = 0.0;
        if (j < N) { const double2 a = v[j]; q = a.x * a.x + a.y * a.y; }
        p[i] = q;
        acc += q;
    }
    tot[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double run = before;
        for (int t = 0; t < QR_BLOCK; ++t) { const double x = tot[t]; tot[t] = run; run += x; }
    }
    __syncthreads();
    // the first thread whose running range crosses u owns the answer
    double run = tot[threadIdx.x];
    i64 mine = -1;
    for (int i = 0; i < per; ++i) {
        run += p[i];
        if (mine < 0 && run >= u) mine = (i64)(base + (u64)threadIdx.x * per + i);
    }
    // threads are ordered by index: the first thread with a hit owns the answer
    hits[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        i64 r = -1;
        for (int t = 0; t < QR_BLOCK && r < 0; ++t) r = hits[t];
        if (r < 0) r = (i64)((base + QR_SCAN_CHUNK < N ? base + QR_SCAN_CHUNK : N) - 1);  // rounding at the chunk edge
        out[shot] = r;
    }
}

// psi'[k] = psi[perm[k]]: measurement-basis ordering for the inverse-CDF sampler (mc_clean.py:255-261 samples
// in the order of the sorted eigenvalues of the observable, not in index order)
__global__ void k_permute_gather(const double2* __restrict__ src, const i64* __restrict__ perm, double2* __restrict__ dst, u64 N) {
    for (u64 k = (u64)blockIdx.x * blockDim.x + threadIdx.x; k < N; k += (u64)gridDim.x * blockDim.x) dst[k] = src[perm[k]];
}

// psi' = M psi for a dense N x N matrix (row major): the change into the eigenbasis of an observable with x / y terms
// (mc_clean.py:221-224 diagonalises the dense observable; small registers only).  One warp per row, fp64 complex dot
// product, fixed shuffle tree.  HBM bound on the matrix: 16 B per entry.
__global__ void k_dense_matvec(const double2* __restrict__ M, const double2* __restrict__ x, double2* __restrict__ y, u64 N) {
    const int lane = threadIdx.x & 31;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 row = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < N; row += warps) {
        double2 acc = make_double2(0.0, 0.0);
        for (u64 c = lane; c < N; c += 32) acc = cadd(acc, cmul(M[row * N + c], x[c]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (lane == 0) y[row] = acc;
    }
}

__global__ void k_gather_f64(const double* __restrict__ table, const i64* __restrict__ idx, int n, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = table[idx[i]];
}


// ------------------------------------------------------------------------------------------
// Batched circuits: build the per-(circuit, layer, pass) gate tables on the device from the raw
// axes / angles arrays (the host loop over 8192 x 28 x 2 x 12 entries dominated the e2e time).
//   tab[b][lay][p][s], lay = dir * L + i;  qmap[dir][p][s] = qubit handled by slot s, or -1.
// ------------------------------------------------------------------------------------------
struct GatePOut { double c, s; int axis; int pad; };   // same layout as GateP (qr_tile.cuh)

__global__ void k_build_gates(const int* __restrict__ axes, const double* __restrict__ angles,
                              const int* __restrict__ qmap, GatePOut* __restrict__ tab, i64 batch, int L, int n,
                              int P, int GS, int ndir, int qmap_per_layer) {
    // qmap[dir][layer or 0][pass][slot] = qubit + 64 * pad (pad > 0: Rz applied through the tile's index bit pad - 1,
    // axis-aware plans) or -1
    const i64 per_batch = (i64)ndir * L * P * GS;
    const i64 total = batch * per_batch;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        const i64 b = idx / per_batch;
        i64 r = idx - b * per_batch;
        const int s = (int)(r % GS); r /= GS;
        const int p = (int)(r % P); r /= P;
        const int lay = (int)r;
        const int dir = lay / L, i = lay - dir * L;
        const int qm = qmap[(((size_t)dir * (qmap_per_layer ? L : 1) + (qmap_per_layer ? i : 0)) * P + p) * GS + s];
        const int q = qm < 0 ? -1 : (qm & 63);
        GatePOut g;
        g.c = 1.0; g.s = 0.0; g.axis = -1; g.pad = qm < 0 ? 0 : (qm >> 6);
        if (q >= 0) {
            const double th = 0.5 * angles[((size_t)b * L + i) * n + q];
            double sn, cs;
            sincos(th, &sn, &cs);
            g.c = cs;
            g.s = dir == 0 ? sn : -sn;
            g.axis = axes[((size_t)b * L + i) * n + q];
        }
        tab[idx] = g;
    }
}


// ------------------------------------------------------------------------------------------
// Device-side optimiser step (optimization.py:131-194: Adam, GradientDescent with a constant step,
// RateDecayOnPlateau).  One block: scatter the slot sums of the backward sweep into grad[L][n], record the
// cost, update the parameters in place.  Enqueued right after the gradient's kernels, so a whole optimisation
// run needs no host round trip.
// ------------------------------------------------------------------------------------------
struct OptDev {
    int rule;              // 0 Adam, 1 GradientDescent, 2 RateDecayOnPlateau
    int iter;
    int plateau_length, plateau_counter;
    double step_size, beta1, beta2, eps, decay_rate, cost;
    int step;              // index of the next history entry: kept on the device so that every step is the same launch sequence (CUDA graph replay)
    int pad_;
};

// the update of one optimiser step on np_ parameters whose gradient is in grad[] (shared tail of both loops)
__device__ __forceinline__ void qr_opt_update(int np_, double e, double* __restrict__ params, double* __restrict__ m, double* __restrict__ v,
                                              const double* __restrict__ grad, OptDev* st, double* __restrict__ cost_hist,
                                              double* __restrict__ param_hist) {
    __shared__ double sh[3];   // step size of this step, 1 - beta1^iter, 1 - beta2^iter
    __shared__ int s_it;
    if (threadIdx.x == 0) s_it = st->step;
    __syncthreads();           // grad[] was written by this block
    const int it = s_it;
    if (threadIdx.x == 0) {
        cost_hist[it] = e;
        st->step = it + 1;
        st->iter += 1;
        if (st->rule == 2) {               // optimization.py:183-192
            if (e > st->cost) {
                st->plateau_counter += 1;
                if (st->plateau_counter >= st->plateau_length) { st->step_size *= st->decay_rate; st->plateau_counter = 0; }
            } else {
                st->cost = e;
                st->plateau_counter = 0;
            }
        }
        sh[0] = st->step_size;
        sh[1] = 1.0 - pow(st->beta1, (double)st->iter);
        sh[2] = 1.0 - pow(st->beta2, (double)st->iter);
    }
    __syncthreads();
    const double step = sh[0];
    for (int k = threadIdx.x; k < np_; k += blockDim.x) {
        const double g = grad[k];
        double x = params[k];
        if (st->rule == 0) {               // optimization.py:148-154
            const double mk = st->beta1 * m[k] + (1.0 - st->beta1) * g;
            const double vk = st->beta2 * v[k] + (1.0 - st->beta2) * (g * g);
            m[k] = mk;
            v[k] = vk;
            x -= step * (mk / sh[1]) / (sqrt(vk / sh[2]) + st->eps);
        } else {
            x -= step * g;
        }
        params[k] = x;
        if (param_hist) param_hist[(size_t)it * np_ + k] = x;
    }
}

// McClean: gradient [L][n] from the slot sums of the backward passes, then the update (optimization.py:41-91)
__global__ void k_opt_step(const double* __restrict__ result, const int* __restrict__ slot_qubit, int L, int n, int P, int GS, int SL,
                           double* __restrict__ params, double* __restrict__ m, double* __restrict__ v, double* __restrict__ grad,
                           OptDev* st, double* __restrict__ cost_hist, double* __restrict__ param_hist) {
    // slot_qubit[layer][pass][slot]: the plans (hence the slot maps) differ between layers with the axis-aware plans
    const int total_slots = L * P * GS, np_ = L * n;
    for (int idx = threadIdx.x; idx < total_slots; idx += blockDim.x) {
        const int i = idx / (GS * P);
        const int q = slot_qubit[idx];
        if (q >= 0) grad[i * n + q] = result[1 + (size_t)(idx / GS) * SL + idx % GS];
    }
    qr_opt_update(np_, result[0], params, m, v, grad, st, cost_hist, param_hist);
}

// QAOA: gradient rows (d/d beta_i, d/d gamma_i) from the slot sums (qaoa.py:62-68), then the update (optimization.py:113-129)
__global__ void k_qaoa_opt_step(const double* __restrict__ result, const int* __restrict__ slot_qubit, int L, int P, int GS, int SL,
                                double* __restrict__ params, double* __restrict__ m, double* __restrict__ v, double* __restrict__ grad,
                                OptDev* st, double* __restrict__ cost_hist, double* __restrict__ param_hist) {
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        double gb = 0.0;
        for (int p = 0; p < P; ++p)
            for (int s = 0; s < GS; ++s)
                if (slot_qubit[p * GS + s] >= 0) gb += result[1 + (size_t)(i * P + p) * SL + s];
        grad[2 * i] = gb;
        grad[2 * i + 1] = 2.0 * result[1 + (size_t)(i * P + (P - 1)) * SL + (SL - 1)];
    }
    qr_opt_update(2 * L, result[0], params, m, v, grad, st, cost_hist, param_hist);
}

// QAOA gate tables and phase look-up tables from device-resident parameter rows (beta_i, gamma_i):
// tab[dir][layer][pass][slot] = X rotation by +-beta_i on every slot that holds a qubit; lut[dir][layer][v] = exp(-+ i gamma_i (hmin + v))
__global__ void k_qaoa_tables(const double* __restrict__ params, const int* __restrict__ slot_qubit, int L, int P, int GS, int dirs,
                              double hmin, int range, GatePOut* __restrict__ tab, double2* __restrict__ lut) {
    const i64 ntab = (i64)dirs * L * P * GS, nlut = (i64)2 * L * range;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < ntab + nlut; idx += (i64)gridDim.x * blockDim.x) {
        if (idx < ntab) {
            const int s = (int)(idx % GS), p = (int)((idx / GS) % P), i = (int)((idx / ((i64)GS * P)) % L), dir = (int)(idx / ((i64)GS * P * L));
            GatePOut g;
            g.c = 1.0; g.s = 0.0; g.axis = -1; g.pad = 0;
            if (slot_qubit[p * GS + s] >= 0) {
                double sn, cs;
                sincos(0.5 * params[2 * i], &sn, &cs);
                g.c = cs; g.s = dir == 0 ? sn : -sn; g.axis = 0;
            }
            tab[idx] = g;
        } else {
            const i64 j = idx - ntab;
            const int v = (int)(j % range), i = (int)((j / range) % L), dir = (int)(j / ((i64)range * L));
            const double ang = (dir == 0 ? params[2 * i + 1] : -params[2 * i + 1]) * (hmin + (double)v);
            double sn, cs;
            sincos(ang, &sn, &cs);
            lut[j] = make_double2(cs, -sn);
        }
    }
}

// integer-valued diagonal Hamiltonian -> 16-bit index table: hidx[j] = H[j] - hmin
__global__ void k_ham_index(const double* __restrict__ ham, short* __restrict__ hidx, u64 N, double hmin) {
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += (u64)gridDim.x * blockDim.x)
        hidx[j] = (short)__double2int_rn(ham[j] - hmin);
}
