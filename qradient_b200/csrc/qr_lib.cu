// libqradient_b200: C ABI (include/qradient_b200.h) + host-side pass planner.
//
// Single translation unit: kernels live in qr_kernels.cuh (gate-at-a-time, reductions,
// sampling) and qr_tile.cuh (fused tile pass).  Everything here is host orchestration: buffer
// ping-pong, the per-layer pass schedule, gate tables, and the mapping of per-CTA gradient
// partials back to (layer, qubit).
#include "../../include/qradient_b200.h"
#include "qr_kernels.cuh"
#include "qr_tile.cuh"
#include "qr_tile12.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess)                                                                        \
            return fail(e_ == cudaErrorMemoryAllocation ? QR_ENOMEM : QR_ECUDA, "%s: %s (%s:%d)", #x, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                  \
    } while (0)
#define QR_TRY(x)            \
    do {                     \
        int r_ = (x);        \
        if (r_) return r_;   \
    } while (0)
#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())

// ------------------------------------------------------------------------------------------
// objects
// ------------------------------------------------------------------------------------------
struct qr_obs {
    int n;
    std::vector<ObsTerm> terms;   // projector order
    bool diagonal;                // only z / zz terms
};

struct PassPlan {
    int k, c, h, nrounds;
    int g[QR_MAXROUNDS];
    int gbit[QR_GATE_SLOTS];   // global index bit handled by slot (round * R + register bit) or -1
    bool lean;                 // lean static kernel k_tile12 (k == 12 or 11): slot = local bit
    int ngroups;               // lean: active register groups (1..4)
    int m1, h2;                // lean: two-segment geometry (Geo12); single segment: m1 = k - c, h2 = h + m1
    bool gx;                   // axis-aware pass (plan_axis_layer): general geometry lbit[], out-of-tile Rz gates in the row slots
    int lbit[QR_MAX_TILE_BITS];   // gx: global index bit of local bit b
    unsigned zmask;            // gx: slots (local bits < c) that hold an Rz on an index bit outside the tile
};

struct LayerPlan {
    int n, k, R, npasses;
    PassPlan pass[16];
};

#define QR_NBUF 4
#define QR_PDL_AUTO_MAX_QUBITS 22   // measured (profiles/r1_pdl_sweep.log, r1_ab_pdl.log): +24 % at n = 16, +12 % at 20, +3.5 % at 22, +1 % at 24; -4 % for QAOA-26 and n = 30

struct qr_ctx {
    int n = 0;
    u64 N = 0;
    int device = 0;
    int sm_count = 1;
    cudaStream_t stream = nullptr;
    double2* buf[QR_NBUF] = {nullptr, nullptr, nullptr, nullptr};
    void* buf_base[QR_NBUF] = {nullptr, nullptr, nullptr, nullptr};   // raw allocations
    u64 buf_amps = 0;            // capacity of each buffer in amplitudes (>= N; batch paths grow it)
    int psi = 0;                 // buffer holding the state vector
    double* d_ham = nullptr;     // diagonal Hamiltonian table [N]
    i64* d_perm = nullptr;       // index permutation for qr_state_permute [N]
    double2* d_dense = nullptr;  // dense N x N basis change for qr_state_apply_dense (small registers)
    bool ham_loaded = false;
    short* d_hidx = nullptr;     // integer-valued H: H = hmin + hidx (phase look-up table path)
    bool ham_integer = false;
    double ham_min = 0.0;
    int ham_range = 0;
    double* d_scratch = nullptr; // reduction partials
    size_t scratch_cap = 0;      // in doubles
    unsigned* d_counter = nullptr;   // arrival counter of the fused final reduction (kept at zero between launches)
    double* d_result = nullptr;  // small results
    size_t result_cap = 0;
    void* d_small = nullptr;     // gate tables / observable terms / tables
    size_t small_cap = 0;
    unsigned char* h_pin = nullptr;
    size_t pin_cap = 0;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // options
    long long opt_fusion = 1, opt_tile_bits = 0 /* auto */, opt_prefetch = 1;
    long long opt_ctas_fwd = 2, opt_ctas_bwd = 1, opt_final_ladder = 1, opt_ham_lut = 1;
    long long opt_tile_bits_x = 0, opt_min_row_bits = 3, opt_batch_chunk_mb = 0;
    long long opt_staged = 4;      // bit0 backward, bit1 forward: next tile staged in shared memory by asynchronous copies; bit2: auto (backward)
    long long opt_staged_min_bit = 21;   // auto mode: strided backward passes whose lowest gate bit is >= this are staged
    long long opt_defer_reduce = 1; // single circuits: one reduction launch per gradient instead of a last-CTA reduction in every backward pass
    long long opt_pdl = 1;         // k_tile12 passes launched with programmatic stream serialization: 0 off, 1 auto (n <= 22), 2 always
    long long opt_axis_plan = 15;  // single McClean circuits: bit 0 per-layer plans from the axes (plan_axis_layer); bit 1 pass 0 trades Rz-only bits for high X / Y bits where that saves a round; bit 2 passes without an exchange run the two-round program; bit 3 split barriers in two-round backward passes
    long long opt_shard_zskip = 1; // sharded states: Rz on a global qubit is applied as a per-subgroup phase, without the exchange
    bool axis_plan_forced = false; // QR_OPT_AXIS_PLAN was set explicitly: the plans apply at every size; by default only from 20 qubits on (n = 16: +5 % with them, the per-launch table set-up; n = 20: -3 %)
    long long opt_loop_graph = 1;  // device optimiser loops: steps 2..N replay a CUDA graph captured from the second step
    bool capturing = false;        // a step of a device optimiser loop is being captured into a graph: no tracing
    bool tables_fresh = true;      // gate / phase tables were (re)written since the last tile pass: see launch_pass
    qr_perf perf;
    // ---- sharded states: this context holds one shard of an n_total-qubit register ----
    int n_total = 0, g = 0, rank = 0;
    double2* peer[QR_MAX_RANKS][QR_NBUF];
    bool peer_mapped[QR_MAX_RANKS][QR_NBUF];
    struct ShardRun* run = nullptr;
    struct SwapRun* srun = nullptr;                       // swap engine (qr_shard.cuh)
    unsigned long long* d_flags = nullptr;                // device-side ordering flags of this rank (READY / DONE / ERR per rank)
    unsigned long long* peer_flags[QR_MAX_RANKS];         // every rank's flag array (own entry = d_flags)
    bool flags_mapped[QR_MAX_RANKS];
    long long flag_gen = 0;                               // last generation used by a swap-engine run
    double* d_ham_ly[2] = {nullptr, nullptr};             // sharded QAOA: H over this shard's local index in the natural / swapped layout
    short* d_hidx_ly[2] = {nullptr, nullptr};
    std::vector<ObsTerm> ham_ly_terms;                    // terms the tables were built from
    long long opt_shard_mode = 0;                         // 0 auto, 1 "peer" engine (round 1), 2 "swap" engine
    long long opt_shard_lockstep = 1;                     // 1: the caller runs the steps in lockstep over the ranks; 0: device-side flags
    long long opt_shard_slices = 1;                       // swap engine: the last local pass and the exchange pass of a layer are issued in this many slices (1 = off, the default: measured -4 % at 8 GPUs, +11 % at 2, profiles/README.md)
    long long opt_shard_xsms = 60;                        // asynchronous mode: SMs given to the exchange pass while the local pass of the next slice runs on the others
    cudaStream_t stream2 = nullptr;                       // second stream of the sliced exchange passes
    cudaEvent_t ev_pair[2] = {nullptr, nullptr};
    std::vector<double2*> snapshots;   // device copies of the state vector (qr_state_save / qr_state_load)
};

static inline int use_device(qr_ctx* c) {
    CUDA_TRY(cudaSetDevice(c->device));
    return 0;
}

static int ensure_dev(void** p, size_t* cap, size_t need_bytes) {
    if (*cap >= need_bytes) return 0;
    if (*p) CUDA_TRY(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t sz = std::max(need_bytes, (size_t)4096);
    CUDA_TRY(cudaMalloc(p, sz));
    *cap = sz;
    return 0;
}

static int ensure_scratch(qr_ctx* c, size_t doubles) {
    size_t cap = c->scratch_cap * sizeof(double);
    QR_TRY(ensure_dev((void**)&c->d_scratch, &cap, doubles * sizeof(double)));
    c->scratch_cap = cap / sizeof(double);
    return 0;
}
static int ensure_result(qr_ctx* c, size_t doubles) {
    size_t cap = c->result_cap * sizeof(double);
    QR_TRY(ensure_dev((void**)&c->d_result, &cap, doubles * sizeof(double)));
    c->result_cap = cap / sizeof(double);
    return 0;
}
static int ensure_small(qr_ctx* c, size_t bytes) { return ensure_dev(&c->d_small, &c->small_cap, bytes); }

static int ensure_pin(qr_ctx* c, size_t bytes) {
    if (c->pin_cap >= bytes) return 0;
    if (c->h_pin) CUDA_TRY(cudaFreeHost(c->h_pin));
    c->h_pin = nullptr;
    c->pin_cap = 0;
    size_t sz = std::max(bytes, (size_t)1 << 16);
    CUDA_TRY(cudaMallocHost((void**)&c->h_pin, sz));
    c->pin_cap = sz;
    return 0;
}

static int ensure_buf(qr_ctx* c, int i) {
    if (c->buf[i]) return 0;
    cudaError_t e = cudaMalloc(&c->buf_base[i], c->buf_amps * sizeof(double2));
    if (e == cudaSuccess) c->buf[i] = (double2*)c->buf_base[i];
    if (e != cudaSuccess) {
        c->buf[i] = nullptr;
        c->buf_base[i] = nullptr;
        return fail(QR_ENOMEM, "cannot allocate state buffer %d of %.2f GiB for %d qubits: %s", i,
                    (double)(c->buf_amps * sizeof(double2)) / (double)(1ull << 30), c->n, cudaGetErrorString(e));
    }
    return 0;
}

// a buffer index different from a, b, d (pass -1 for unused)
static int other_buf(qr_ctx* c, int a, int b = -1, int d = -1) {
    for (int i = 0; i < QR_NBUF; ++i)
        if (i != a && i != b && i != d) return i;
    return -1;
}

static inline int grid_for(const qr_ctx* c, u64 work) {
    u64 blocks = (work + QR_BLOCK - 1) / QR_BLOCK;
    u64 cap = (u64)c->sm_count * 8;
    if (blocks < 1) blocks = 1;
    return (int)std::min(blocks, cap);
}

// sum scratch partials [nunits][nvals] -> host
static int reduce_to_host(qr_ctx* c, int nunits, int nvals, double* out) {
    QR_TRY(ensure_result(c, nvals));
    QR_TRY(ensure_pin(c, nvals * sizeof(double)));
    QR_LAUNCH(k_reduce_partials, 1, QR_BLOCK, 0, c->stream, c->d_scratch, nunits, nvals, c->d_result);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemcpyAsync(c->h_pin, c->d_result, nvals * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out, c->h_pin, nvals * sizeof(double));
    return 0;
}

// upload a small host array into d_small at byte offset `off` (stream ordered, via pinned staging)
static int upload_small(qr_ctx* c, size_t off, const void* src, size_t bytes, size_t pin_off) {
    memcpy(c->h_pin + pin_off, src, bytes);
    CUDA_TRY(cudaMemcpyAsync((char*)c->d_small + off, c->h_pin + pin_off, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// misc API
// ------------------------------------------------------------------------------------------
extern "C" const char* qr_last_error(void) { return g_err.c_str(); }
extern "C" int qr_version(void) { return 100; }

extern "C" int qr_device_count(int* out) {
    if (!out) return fail(QR_EINVAL, "null output");
    CUDA_TRY(cudaGetDeviceCount(out));
    return 0;
}

extern "C" int qr_ctx_create(int n_qubits, int device, qr_ctx** out) {
    if (!out) return fail(QR_EINVAL, "null output");
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > 34) return fail(QR_EINVAL, "qubit_number must be in [1, 34], got %d", n_qubits);
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(QR_ECUDA, "no CUDA device visible: qradient_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(QR_EINVAL, "device %d out of range (%d visible)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    qr_ctx* c = new qr_ctx();
    c->n = n_qubits;
    c->N = (u64)1 << n_qubits;
    c->buf_amps = c->N;
    c->device = device;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    memset(&c->perf, 0, sizeof(c->perf));
    memset(c->peer, 0, sizeof(c->peer));
    memset(c->peer_mapped, 0, sizeof(c->peer_mapped));
    memset(c->peer_flags, 0, sizeof(c->peer_flags));
    memset(c->flags_mapped, 0, sizeof(c->flags_mapped));
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(QR_ECUDA, "stream creation failed"); break; }
        for (int i = 0; i < 4; ++i) cudaEventCreate(&c->ev[i]);
        if ((rc = ensure_buf(c, 0))) break;
        if ((rc = ensure_scratch(c, (size_t)c->sm_count * 8 * 16))) break;
        if ((rc = ensure_result(c, 64))) break;
        if (cudaMalloc((void**)&c->d_counter, 64) != cudaSuccess) { rc = fail(QR_ENOMEM, "counter allocation failed"); break; }
        cudaMemsetAsync(c->d_counter, 0, 64, c->stream);
        if ((rc = ensure_small(c, 1 << 16))) break;
        if ((rc = ensure_pin(c, 1 << 16))) break;
    } while (0);
    if (rc) { qr_ctx_destroy(c); return rc; }
    *out = c;
    return qr_state_init(c, 0);
}

static void shard_release(qr_ctx* c);

extern "C" int qr_ctx_destroy(qr_ctx* c) {
    if (!c) return 0;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    shard_release(c);
    for (int i = 0; i < QR_NBUF; ++i) if (c->buf_base[i]) cudaFree(c->buf_base[i]);
    if (c->d_ham) cudaFree(c->d_ham);
    if (c->d_perm) cudaFree(c->d_perm);
    if (c->d_dense) cudaFree(c->d_dense);
    if (c->d_hidx) cudaFree(c->d_hidx);
    for (double2* sp : c->snapshots) if (sp) cudaFree(sp);
    if (c->d_scratch) cudaFree(c->d_scratch);
    if (c->d_result) cudaFree(c->d_result);
    if (c->d_counter) cudaFree(c->d_counter);
    if (c->d_flags) cudaFree(c->d_flags);
    for (int i = 0; i < 2; ++i) { if (c->d_ham_ly[i]) cudaFree(c->d_ham_ly[i]); if (c->d_hidx_ly[i]) cudaFree(c->d_hidx_ly[i]); }
    if (c->d_small) cudaFree(c->d_small);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; ++i) if (c->ev_pair[i]) cudaEventDestroy(c->ev_pair[i]);
    if (c->stream2) { cudaStreamSynchronize(c->stream2); cudaStreamDestroy(c->stream2); }
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

extern "C" int qr_set_option(qr_ctx* c, int key, long long v) {
    if (!c) return fail(QR_EINVAL, "null context");
    switch (key) {
        case QR_OPT_FUSION: c->opt_fusion = v ? 1 : 0; break;
        case QR_OPT_TILE_BITS:
            if (v != 0 && (v < 4 || v > QR_MAX_TILE_BITS)) return fail(QR_EINVAL, "tile bits must be 0 (auto) or in [4, %d]", QR_MAX_TILE_BITS);
            c->opt_tile_bits = v; break;
        case QR_OPT_PREFETCH: if (v < 0 || v > 31) return fail(QR_EINVAL, "prefetch must be in [0, 31]"); c->opt_prefetch = v; break;
        case QR_OPT_STAGED: if (v < 0 || v > 7) return fail(QR_EINVAL, "bad staged mode"); c->opt_staged = v; break;
        case QR_OPT_STAGED_MIN_BIT: if (v < 0 || v > 64) return fail(QR_EINVAL, "bad staged min bit"); c->opt_staged_min_bit = v; break;
        case QR_OPT_CTAS_PER_SM_FWD: if (v < 1 || v > 8) return fail(QR_EINVAL, "bad CTAs/SM"); c->opt_ctas_fwd = v; break;
        case QR_OPT_CTAS_PER_SM_BWD: if (v < 1 || v > 8) return fail(QR_EINVAL, "bad CTAs/SM"); c->opt_ctas_bwd = v; break;
        case QR_OPT_FINAL_LADDER: c->opt_final_ladder = v ? 1 : 0; break;
        case QR_OPT_HAM_LUT: c->opt_ham_lut = v ? 1 : 0; break;
        case QR_OPT_TILE_BITS_STRIDED: if (v != 0 && (v < 4 || v > QR_MAX_TILE_BITS)) return fail(QR_EINVAL, "bad strided tile bits"); c->opt_tile_bits_x = v; break;
        case QR_OPT_MIN_ROW_BITS: if (v < 1 || v > 11) return fail(QR_EINVAL, "bad min row bits"); c->opt_min_row_bits = v; break;
        case QR_OPT_BATCH_CHUNK_MB: if (v < 0 || v > 65536) return fail(QR_EINVAL, "bad batch chunk"); c->opt_batch_chunk_mb = v; break;
        case QR_OPT_DEFER_REDUCE: c->opt_defer_reduce = v ? 1 : 0; break;
        case QR_OPT_SHARD_ZSKIP: c->opt_shard_zskip = v ? 1 : 0; break;
        case QR_OPT_LOOP_GRAPH: c->opt_loop_graph = v ? 1 : 0; break;
        case QR_OPT_AXIS_PLAN: if (v < 0 || v > 15) return fail(QR_EINVAL, "bad axis-plan mode"); c->opt_axis_plan = v; c->axis_plan_forced = true; break;
        case QR_OPT_PDL: if (v < 0 || v > 2) return fail(QR_EINVAL, "bad PDL mode"); c->opt_pdl = v; break;
        case QR_OPT_SHARD_MODE: if (v < 0 || v > 2) return fail(QR_EINVAL, "bad shard mode"); c->opt_shard_mode = v; break;
        case QR_OPT_SHARD_LOCKSTEP: c->opt_shard_lockstep = v ? 1 : 0; break;
        case QR_OPT_SHARD_SLICES: if (v != 1 && v != 2 && v != 4 && v != 8) return fail(QR_EINVAL, "slices must be 1, 2, 4 or 8"); c->opt_shard_slices = v; break;
        case QR_OPT_SHARD_XSMS: if (v < 1 || v > 1024) return fail(QR_EINVAL, "bad SM count"); c->opt_shard_xsms = v; break;
        default: return fail(QR_EINVAL, "unknown option %d", key);
    }
    return 0;
}

extern "C" int qr_get_option(qr_ctx* c, int key, long long* v) {
    if (!c || !v) return fail(QR_EINVAL, "null argument");
    switch (key) {
        case QR_OPT_FUSION: *v = c->opt_fusion; break;
        case QR_OPT_TILE_BITS: *v = c->opt_tile_bits; break;
        case QR_OPT_PREFETCH: *v = c->opt_prefetch; break;
        case QR_OPT_CTAS_PER_SM_FWD: *v = c->opt_ctas_fwd; break;
        case QR_OPT_CTAS_PER_SM_BWD: *v = c->opt_ctas_bwd; break;
        case QR_OPT_FINAL_LADDER: *v = c->opt_final_ladder; break;
        case QR_OPT_HAM_LUT: *v = c->opt_ham_lut; break;
        case QR_OPT_TILE_BITS_STRIDED: *v = c->opt_tile_bits_x; break;
        case QR_OPT_MIN_ROW_BITS: *v = c->opt_min_row_bits; break;
        case QR_OPT_BATCH_CHUNK_MB: *v = c->opt_batch_chunk_mb; break;
        case QR_OPT_STAGED: *v = c->opt_staged; break;
        case QR_OPT_STAGED_MIN_BIT: *v = c->opt_staged_min_bit; break;
        case QR_OPT_PDL: *v = c->opt_pdl; break;
        case QR_OPT_SHARD_ZSKIP: *v = c->opt_shard_zskip; break;
        case QR_OPT_AXIS_PLAN: *v = c->opt_axis_plan; break;
        case QR_OPT_LOOP_GRAPH: *v = c->opt_loop_graph; break;
        case QR_OPT_DEFER_REDUCE: *v = c->opt_defer_reduce; break;
        case QR_OPT_SHARD_MODE: *v = c->opt_shard_mode; break;
        case QR_OPT_SHARD_LOCKSTEP: *v = c->opt_shard_lockstep; break;
        case QR_OPT_SHARD_SLICES: *v = c->opt_shard_slices; break;
        case QR_OPT_SHARD_XSMS: *v = c->opt_shard_xsms; break;
        default: return fail(QR_EINVAL, "unknown option %d", key);
    }
    return 0;
}

extern "C" int qr_perf_last(qr_ctx* c, qr_perf* out) {
    if (!c || !out) return fail(QR_EINVAL, "null argument");
    *out = c->perf;
    return 0;
}

extern "C" int qr_sync(qr_ctx* c) {
    if (!c) return fail(QR_EINVAL, "null context");
    QR_TRY(use_device(c));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// ------------------------------------------------------------------------------------------
// state vector
// ------------------------------------------------------------------------------------------
extern "C" int qr_state_init(qr_ctx* c, int which) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (which != 0 && which != 1) return fail(QR_EINVAL, "Invalid initialization format %d.", which);
    QR_TRY(use_device(c));
    const double amp = std::pow(2.0, -0.5 * c->n);   // state.py:69
    QR_LAUNCH(k_init_basis, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N, which, amp);
    KERNEL_CHECK();
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_state_upload(qr_ctx* c, const double* re_im, size_t n_amps) {
    if (!c || !re_im) return fail(QR_EINVAL, "null argument");
    if (n_amps != c->N) return fail(QR_EINVAL, "state vector must have 2^%d = %llu amplitudes, got %llu", c->n, c->N, (u64)n_amps);
    QR_TRY(use_device(c));
    CUDA_TRY(cudaMemcpyAsync(c->buf[c->psi], re_im, n_amps * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qr_state_download(qr_ctx* c, double* re_im, size_t n_amps) {
    if (!c || !re_im) return fail(QR_EINVAL, "null argument");
    if (n_amps != c->N) return fail(QR_EINVAL, "state vector has 2^%d = %llu amplitudes, got room for %llu", c->n, c->N, (u64)n_amps);
    QR_TRY(use_device(c));
    CUDA_TRY(cudaMemcpyAsync(re_im, c->buf[c->psi], n_amps * sizeof(double2), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qr_state_device_ptr(qr_ctx* c, void** out) {
    if (!c || !out) return fail(QR_EINVAL, "null argument");
    *out = c->buf[c->psi];
    return 0;
}

// device-side snapshots of the state vector (the reference keeps numpy copies: state_history)
extern "C" int qr_state_save(qr_ctx* c, int slot) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (slot < 0 || slot > 4096) return fail(QR_EINVAL, "bad snapshot slot %d", slot);
    QR_TRY(use_device(c));
    if ((size_t)slot >= c->snapshots.size()) c->snapshots.resize(slot + 1, nullptr);
    if (!c->snapshots[slot]) {
        cudaError_t e = cudaMalloc((void**)&c->snapshots[slot], c->N * sizeof(double2));
        if (e != cudaSuccess) { c->snapshots[slot] = nullptr; return fail(QR_ENOMEM, "cannot allocate snapshot %d: %s", slot, cudaGetErrorString(e)); }
    }
    CUDA_TRY(cudaMemcpyAsync(c->snapshots[slot], c->buf[c->psi], c->N * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_state_load(qr_ctx* c, int slot) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (slot < 0 || (size_t)slot >= c->snapshots.size() || !c->snapshots[slot]) return fail(QR_EINVAL, "snapshot %d does not exist", slot);
    QR_TRY(use_device(c));
    CUDA_TRY(cudaMemcpyAsync(c->buf[c->psi], c->snapshots[slot], c->N * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_state_free_snapshots(qr_ctx* c) {
    if (!c) return fail(QR_EINVAL, "null context");
    QR_TRY(use_device(c));
    for (double2*& sp : c->snapshots) { if (sp) cudaFree(sp); sp = nullptr; }
    c->snapshots.clear();
    return 0;
}

static int check_axis_qubit(qr_ctx* c, int axis, int qubit) {
    if (axis < 0 || axis > 2) return fail(QR_EINVAL, "Invalid axis %d", axis);
    if (qubit < 0 || qubit >= c->n) return fail(QR_EINVAL, "Invalid qubit index %d for %d qubits.", qubit, c->n);
    return 0;
}

// rotation matrices exp(-i angle P / 2) (state.py:90-92,142-144,168-170)
static Mat2 rot_matrix(int axis, double angle) {
    const double cs = std::cos(0.5 * angle), sn = std::sin(0.5 * angle);
    Mat2 m;
    if (axis == 0) {
        m.m00 = make_double2(cs, 0); m.m01 = make_double2(0, -sn);
        m.m10 = make_double2(0, -sn); m.m11 = make_double2(cs, 0);
    } else if (axis == 1) {
        m.m00 = make_double2(cs, 0); m.m01 = make_double2(-sn, 0);
        m.m10 = make_double2(sn, 0); m.m11 = make_double2(cs, 0);
    } else {
        m.m00 = make_double2(cs, -sn); m.m01 = make_double2(0, 0);
        m.m10 = make_double2(0, 0); m.m11 = make_double2(cs, sn);
    }
    return m;
}

// derivative matrices (state.py:94-97,146-149,172-175): 1/2 cos G - 1/2 sin 1 ; dz: -+ i/2 e^{-+ i a/2}
static Mat2 drot_matrix(int axis, double angle) {
    const double cs = std::cos(0.5 * angle), sn = std::sin(0.5 * angle);
    Mat2 m;
    if (axis == 0) {
        m.m00 = make_double2(-0.5 * sn, 0); m.m01 = make_double2(0, -0.5 * cs);
        m.m10 = make_double2(0, -0.5 * cs); m.m11 = make_double2(-0.5 * sn, 0);
    } else if (axis == 1) {
        m.m00 = make_double2(-0.5 * sn, 0); m.m01 = make_double2(-0.5 * cs, 0);
        m.m10 = make_double2(0.5 * cs, 0); m.m11 = make_double2(-0.5 * sn, 0);
    } else {
        m.m00 = make_double2(-0.5 * sn, -0.5 * cs); m.m01 = make_double2(0, 0);
        m.m10 = make_double2(0, 0); m.m11 = make_double2(-0.5 * sn, 0.5 * cs);
    }
    return m;
}

static int launch_1q(qr_ctx* c, double2* v, int bit, const Mat2& m) {
    const u64 npairs = c->N >> 1;
    QR_LAUNCH(k_apply_1q, grid_for(c, npairs), QR_BLOCK, 0, c->stream, v, npairs, bit, m);
    KERNEL_CHECK();
    return 0;
}

extern "C" int qr_apply_rot(qr_ctx* c, int axis, double angle, int qubit) {
    if (!c) return fail(QR_EINVAL, "null context");
    QR_TRY(check_axis_qubit(c, axis, qubit));
    QR_TRY(use_device(c));
    QR_TRY(launch_1q(c, c->buf[c->psi], c->n - 1 - qubit, rot_matrix(axis, angle)));
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_apply_drot(qr_ctx* c, int axis, double angle, int qubit) {
    if (!c) return fail(QR_EINVAL, "null context");
    QR_TRY(check_axis_qubit(c, axis, qubit));
    QR_TRY(use_device(c));
    QR_TRY(launch_1q(c, c->buf[c->psi], c->n - 1 - qubit, drot_matrix(axis, angle)));
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

static int launch_cnot(qr_ctx* c, double2* v, int control, int target) {
    const u64 nquads = c->N >> 2;
    QR_LAUNCH(k_cnot, grid_for(c, nquads), QR_BLOCK, 0, c->stream, v, nquads, c->n - 1 - control, c->n - 1 - target);
    KERNEL_CHECK();
    return 0;
}

extern "C" int qr_apply_cnot(qr_ctx* c, int control, int target) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (control == target || control < 0 || target < 0 || control >= c->n || target >= c->n)
        return fail(QR_EINVAL, "Invalid CNOT indecies %d and %d, for %d qubits.", control, target, c->n);
    QR_TRY(use_device(c));
    QR_TRY(launch_cnot(c, c->buf[c->psi], control, target));
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

// scatter masks of ladder `stacking`; the gather for stacking s uses the
// masks of 1-s because ladder(1) = ladder(0)^-1.
static void ladder_masks(int n, int stacking, u64* m1, u64* m2) {
    u64 a = 0, b = 0;
    for (int t = 1; t < n; ++t) {
        const int p = n - 1 - t;
        a |= (u64)1 << p;
        const bool three = stacking == 0 ? (t % 2 == 1) : (t % 2 == 0);
        if (t >= 2 && three) b |= (u64)1 << p;
    }
    *m1 = a;
    *m2 = b;
}

// out-of-place ladder on buffer `src` into buffer `dst`
static int launch_ladder(qr_ctx* c, int src, int dst, int stacking) {
    u64 m1, m2;
    ladder_masks(c->n, 1 - stacking, &m1, &m2);
    QR_LAUNCH(k_ladder_gather, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[src], c->buf[dst], c->N, m1, m2);
    KERNEL_CHECK();
    return 0;
}

extern "C" int qr_apply_cnot_ladder(qr_ctx* c, int stacking, int periodic) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (stacking != 0 && stacking != 1) return fail(QR_EINVAL, "Invalid stacking %d", stacking);
    if (periodic && (c->n % 2 != 0))
        return fail(QR_EINVAL, "CNOT gates in a ladder structure with periodic boundaries are ambiguous for uneven number of qubits.");
    QR_TRY(use_device(c));
    if (!periodic) {
        if (c->n >= 2) {
            const int dst = other_buf(c, c->psi);
            QR_TRY(ensure_buf(c, dst));
            QR_TRY(launch_ladder(c, c->psi, dst, stacking));
            c->psi = dst;
        }
    } else {
        // state.py:235-241 with the wrap-around CNOT(n-1 -> 0) in the odd-start group
        const int n = c->n;
        for (int phase = 0; phase < 2; ++phase) {
            const int start = (stacking == 0) ? (phase == 0 ? 1 : 0) : (phase == 0 ? 0 : 1);
            for (int i = start; i < n; i += 2) QR_TRY(launch_cnot(c, c->buf[c->psi], i, (i + 1) % n));
        }
    }
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_apply_x_summed(qr_ctx* c) {
    if (!c) return fail(QR_EINVAL, "null context");
    QR_TRY(use_device(c));
    const int dst = other_buf(c, c->psi);
    QR_TRY(ensure_buf(c, dst));
    QR_LAUNCH(k_x_summed, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->buf[dst], c->N, c->n);
    KERNEL_CHECK();
    c->psi = dst;
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_norm2(qr_ctx* c, double* out) {
    if (!c || !out) return fail(QR_EINVAL, "null argument");
    QR_TRY(use_device(c));
    const int grid = grid_for(c, c->N);
    QR_TRY(ensure_scratch(c, grid));
    QR_LAUNCH(k_norm2, grid, QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N, c->d_scratch);
    KERNEL_CHECK();
    return reduce_to_host(c, grid, 1, out);
}

// ------------------------------------------------------------------------------------------
// observable
// ------------------------------------------------------------------------------------------
extern "C" int qr_obs_create(int n, int n_terms, const int32_t* kind, const int32_t* qi, const int32_t* qj,
                             const double* w, qr_obs** out) {
    if (!out) return fail(QR_EINVAL, "null output");
    *out = nullptr;
    if (n < 1 || n > 34) return fail(QR_EINVAL, "qubit_number must be in [1, 34], got %d", n);
    if (n_terms < 0 || (n_terms > 0 && (!kind || !qi || !w))) return fail(QR_EINVAL, "bad term arrays");
    qr_obs* o = new qr_obs();
    o->n = n;
    o->diagonal = true;
    for (int k = 0; k < n_terms; ++k) {
        ObsTerm t;
        t.kind = kind[k];
        t.pad = 0;
        t.w = w[k];
        if (t.kind < 0 || t.kind > 3) { delete o; return fail(QR_EINVAL, "Unknown key for projector %d.", t.kind); }
        if (qi[k] < 0 || qi[k] >= n) { delete o; return fail(QR_EINVAL, "observable qubit index %d out of range", qi[k]); }
        t.bit_i = n - 1 - qi[k];
        t.bit_j = 0;
        if (t.kind == QR_TERM_ZZ) {
            if (!qj || qj[k] <= qi[k] || qj[k] >= n) {
                delete o;
                return fail(QR_EINVAL, "zz of observable should be a upper triangular %d by %d matrix.", n, n);
            }
            t.bit_j = n - 1 - qj[k];
        }
        if (t.kind == QR_TERM_X || t.kind == QR_TERM_Y) o->diagonal = false;
        o->terms.push_back(t);
    }
    *out = o;
    return 0;
}

extern "C" int qr_obs_destroy(qr_obs* o) {
    delete o;
    return 0;
}

static int check_obs(qr_ctx* c, const qr_obs* o) {
    if (!c || !o) return fail(QR_EINVAL, "null argument");
    if (o->n != c->n) return fail(QR_EINVAL, "observable is for %d qubits, state has %d", o->n, c->n);
    return 0;
}

// put the observable's terms at the start of d_small; returns device pointer
static int upload_terms(qr_ctx* c, const std::vector<ObsTerm>& terms, const ObsTerm** d_terms) {
    const size_t bytes = std::max<size_t>(terms.size(), 1) * sizeof(ObsTerm);
    QR_TRY(ensure_small(c, bytes));
    QR_TRY(ensure_pin(c, bytes));
    if (!terms.empty()) QR_TRY(upload_small(c, 0, terms.data(), terms.size() * sizeof(ObsTerm), 0));
    *d_terms = (const ObsTerm*)c->d_small;
    return 0;
}

// dst = O src (dst < 0: expectation only); E = Re<src|O|src>
static int observable_pass(qr_ctx* c, const qr_obs* o, int src, int dst, double* e_out) {
    const ObsTerm* d_terms;
    QR_TRY(upload_terms(c, o->terms, &d_terms));
    const int grid = grid_for(c, c->N);
    QR_TRY(ensure_scratch(c, grid));
    QR_LAUNCH(k_apply_obs, grid, QR_BLOCK, 0, c->stream, (const double2*)c->buf[src],
              dst >= 0 ? c->buf[dst] : (double2*)nullptr, c->N, d_terms, (int)o->terms.size(), c->d_scratch, (u64)0, 64,
              PeerTable());
    KERNEL_CHECK();
    return reduce_to_host(c, grid, 1, e_out);
}

extern "C" int qr_apply_observable(qr_ctx* c, const qr_obs* o) {
    QR_TRY(check_obs(c, o));
    QR_TRY(use_device(c));
    const int dst = other_buf(c, c->psi);
    QR_TRY(ensure_buf(c, dst));
    double e;
    QR_TRY(observable_pass(c, o, c->psi, dst, &e));
    c->psi = dst;
    return 0;
}

extern "C" int qr_expec_val(qr_ctx* c, const qr_obs* o, double* out) {
    QR_TRY(check_obs(c, o));
    if (!out) return fail(QR_EINVAL, "null output");
    QR_TRY(use_device(c));
    return observable_pass(c, o, c->psi, -1, out);
}

extern "C" int qr_term_expecs(qr_ctx* c, const qr_obs* o, double* out) {
    QR_TRY(check_obs(c, o));
    if (!out && !o->terms.empty()) return fail(QR_EINVAL, "null output");
    QR_TRY(use_device(c));
    const ObsTerm* d_terms;
    QR_TRY(upload_terms(c, o->terms, &d_terms));
    const int grid = grid_for(c, c->N);
    QR_TRY(ensure_scratch(c, (size_t)grid * QR_TERMS_PER_LAUNCH));
    const int K = (int)o->terms.size();
    for (int k0 = 0; k0 < K; k0 += QR_TERMS_PER_LAUNCH) {
        const int nk = std::min(QR_TERMS_PER_LAUNCH, K - k0);
        QR_LAUNCH(k_term_expecs, grid, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], c->N, d_terms + k0, nk,
                  c->d_scratch);
        KERNEL_CHECK();
        double vals[QR_TERMS_PER_LAUNCH];
        QR_TRY(reduce_to_host(c, grid, QR_TERMS_PER_LAUNCH, vals));
        for (int k = 0; k < nk; ++k) out[k0 + k] = vals[k];
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// diagonal Hamiltonian
// ------------------------------------------------------------------------------------------
extern "C" int qr_ham_load(qr_ctx* c, const qr_obs* o) {
    QR_TRY(check_obs(c, o));
    QR_TRY(use_device(c));
    if (!c->d_ham) {
        cudaError_t e = cudaMalloc((void**)&c->d_ham, c->N * sizeof(double));
        if (e != cudaSuccess) { c->d_ham = nullptr; return fail(QR_ENOMEM, "cannot allocate the Hamiltonian table: %s", cudaGetErrorString(e)); }
    }
    const ObsTerm* d_terms;
    QR_TRY(upload_terms(c, o->terms, &d_terms));
    QR_LAUNCH(k_ham_build, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->d_ham, c->N, d_terms, (int)o->terms.size(), (u64)0);
    KERNEL_CHECK();
    // integer weights => H takes at most 2 sum|w| + 1 integer values: phases come from a small table
    double wsum = 0.0;
    bool integral = true;
    for (const ObsTerm& t : o->terms)
        if (t.kind >= 2) { wsum += std::fabs(t.w); if (t.w != std::floor(t.w)) integral = false; }
    c->ham_integer = false;
    if (integral && 2.0 * wsum + 1.0 <= (double)QR_LUT_MAX) {
        if (!c->d_hidx) {
            cudaError_t e = cudaMalloc((void**)&c->d_hidx, c->N * sizeof(short));
            if (e != cudaSuccess) { c->d_hidx = nullptr; cudaGetLastError(); }
        }
        if (c->d_hidx) {
            c->ham_min = -wsum;
            c->ham_range = (int)(2.0 * wsum + 1.0);
            QR_LAUNCH(k_ham_index, grid_for(c, c->N), QR_BLOCK, 0, c->stream, (const double*)c->d_ham, c->d_hidx, c->N, c->ham_min);
            KERNEL_CHECK();
            c->ham_integer = true;
        }
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->ham_loaded = true;
    return 0;
}

static int need_ham(qr_ctx* c) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (!c->ham_loaded) return fail(QR_ESTATE, "classical Hamiltonian not loaded (call qr_ham_load / Gates.add_classical_ham first)");
    return 0;
}

extern "C" int qr_ham_download(qr_ctx* c, double* out, size_t n_amps) {
    QR_TRY(need_ham(c));
    if (!out || n_amps != c->N) return fail(QR_EINVAL, "bad output buffer");
    QR_TRY(use_device(c));
    CUDA_TRY(cudaMemcpyAsync(out, c->d_ham, c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qr_apply_exp_ham(qr_ctx* c, double angle) {
    QR_TRY(need_ham(c));
    QR_TRY(use_device(c));
    QR_LAUNCH(k_exp_ham, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], (const double*)c->d_ham, c->N, angle);
    KERNEL_CHECK();
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_apply_exp_ham_component(qr_ctx* c, const qr_obs* o, int k, double angle) {
    QR_TRY(check_obs(c, o));
    // components are the diagonal terms in order (state.py:277-294)
    int idx = -1, seen = 0;
    for (size_t t = 0; t < o->terms.size(); ++t)
        if (o->terms[t].kind >= 2) { if (seen == k) { idx = (int)t; break; } ++seen; }
    if (k < 0 || idx < 0) return fail(QR_EINVAL, "classical Hamiltonian component %d out of range", k);
    QR_TRY(use_device(c));
    // vec *= exp(-i angle w z..z): a diagonal with two values = an Rz-like phase by parity; reuse
    // the generic 1-qubit kernel for z, and a two-step parity phase for zz via CNOT-free masks.
    const ObsTerm t = o->terms[idx];
    const double cs = std::cos(angle * t.w), sn = std::sin(angle * t.w);
    if (t.kind == 2) {
        Mat2 m;
        m.m00 = make_double2(cs, -sn); m.m01 = make_double2(0, 0); m.m10 = make_double2(0, 0); m.m11 = make_double2(cs, sn);
        QR_TRY(launch_1q(c, c->buf[c->psi], t.bit_i, m));
    } else {
        // exp(-i a w Z_i Z_j) = CNOT(i->j) . exp(-i a w Z_j) . CNOT(i->j)
        const int qi = c->n - 1 - t.bit_i, qj = c->n - 1 - t.bit_j;
        Mat2 m;
        m.m00 = make_double2(cs, -sn); m.m01 = make_double2(0, 0); m.m10 = make_double2(0, 0); m.m11 = make_double2(cs, sn);
        QR_TRY(launch_cnot(c, c->buf[c->psi], qi, qj));
        QR_TRY(launch_1q(c, c->buf[c->psi], t.bit_j, m));
        QR_TRY(launch_cnot(c, c->buf[c->psi], qi, qj));
    }
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_apply_ham(qr_ctx* c, int mode) {
    QR_TRY(need_ham(c));
    if (mode != 0 && mode != 1) return fail(QR_EINVAL, "bad mode %d", mode);
    QR_TRY(use_device(c));
    QR_LAUNCH(k_mul_ham, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], (const double*)c->d_ham, c->N, mode);
    KERNEL_CHECK();
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

// ------------------------------------------------------------------------------------------
// pass planner
// ------------------------------------------------------------------------------------------
// Rounds of a pass whose gate bits are the local bits [first, k): the top group first (its
// threads' lanes cover the lowest local bits -> coalesced global loads), then ascending groups;
// every gate bit is assigned to the first group that covers it.
static void plan_rounds(PassPlan& pp, int first, int R) {
    const int k = pp.k;
    int assigned[QR_MAX_TILE_BITS];
    for (int i = 0; i < QR_MAX_TILE_BITS; ++i) assigned[i] = -1;
    int nr = 0;
    const int top = std::max(k - R, 0);
    pp.g[nr++] = top;
    for (int g = first; g < top && nr < QR_MAXROUNDS + 4; g += R) {
        if (nr < QR_MAXROUNDS) pp.g[nr] = g;
        ++nr;
    }
    pp.nrounds = nr;
    for (int i = 0; i < QR_GATE_SLOTS; ++i) pp.gbit[i] = -1;
    for (int r = 0; r < std::min(nr, QR_MAXROUNDS); ++r)
        for (int b = 0; b < R; ++b) {
            const int lb = pp.g[r] + b;
            if (lb < first || lb >= k || assigned[lb] >= 0) continue;
            assigned[lb] = r;
            pp.gbit[r * R + b] = lb < pp.c ? lb : pp.h + (lb - pp.c);
        }
}

// Lean plan of a k = 12 pass whose gate bits are the local bits [first, 12): fixed groups
// G0..G3 = local bits 0-2, 3-5, 6-8, 9-11, visited G3 [G0] [G1] [G2]; gradient slot = local bit.
static void plan_lean(PassPlan& pp, int first) {
    for (int i = 0; i < QR_GATE_SLOTS; ++i) pp.gbit[i] = -1;
    const int K = pp.k;
    const Geo12 geo = {pp.c, pp.h, pp.m1, pp.h2, K, 0};
    for (int lb = first; lb < K; ++lb) {
        u64 gidx = geo12_local(geo, (u64)1 << lb);
        int gb = 0;
        while (!((gidx >> gb) & 1)) ++gb;
        pp.gbit[lb] = gb;
    }
    // chain of register groups (Tile12X::ngroups): L = K-3 | [0] | [3] | [6], or L | 2 | 5 for K = 11 with 64 B rows
    if (K == 11 && first == 2) pp.ngroups = 5;
    else pp.ngroups = first < 3 ? 4 : (first < 6 ? 3 : (first < K - 3 ? 2 : 1));
    pp.nrounds = pp.ngroups;
    pp.g[0] = K - 3;
    pp.lean = true;
}

// QR_OPT_TILE_BITS = 0 (auto): half-size tiles (two backward CTAs per SM) where they cost neither a pass nor row
// width against 12-bit tiles -- n = 13..19 and 22..26 (measured: +14 % batched 14x14, +3..5 % at n = 22..26,
// neutral at 20, -10 % at 27: profiles/README.md); 12-bit tiles otherwise.
static int pick_tile_bits(const qr_ctx* c, int n) {
    if (c->opt_tile_bits != 0) return (int)c->opt_tile_bits;
    const bool lean_on = c->opt_tile_bits_x == 0 && c->opt_min_row_bits == 3;
    if (lean_on && ((n >= 13 && n <= 20) || (n >= 22 && n <= 26))) return 11;
    return QR_MAX_TILE_BITS;
}
// n = 20 in auto mode: 11 | 9 with 64 B rows (the 16 MiB state is L2 resident, where narrow rows cost nothing): +5 %
static int pick_min_row_bits(const qr_ctx* c, int n) {
    if (c->opt_tile_bits == 0 && n == 20 && pick_tile_bits(c, n) == 11) return 2;
    return (int)c->opt_min_row_bits;
}

// Pass 0 is the contiguous low-k tile; the remaining n - k index bits are split evenly over strided passes whose
// tiles keep >= 2^min_row_bits contiguous amplitudes per row.  Tiles of 12 or 11 bits run k_tile12 (slot = local
// bit), smaller ones the generic kernel.
static int make_plan(int n, int tile_bits, LayerPlan* lp, int tile_bits_x = 0, int min_row_bits = 3) {
    if (n < 4) return fail(QR_EINVAL, "fused path needs at least 4 qubits");
    const int R = 3;
    const int k = std::min(n, tile_bits);
    lp->n = n;
    lp->k = k;
    lp->R = R;
    int np = 0;
    PassPlan& p0 = lp->pass[np++];
    p0.k = k; p0.c = k; p0.h = k; p0.lean = false; p0.ngroups = 0; p0.m1 = 0; p0.h2 = k; p0.gx = false; p0.zmask = 0;
    const bool lean_k = (k == 12 || k == 11) && (tile_bits_x == 0 || tile_bits_x == k) && (k == 12 || min_row_bits >= 2);
    if (lean_k) plan_lean(p0, 0); else plan_rounds(p0, 0, R);
    const int rem = n - k;
    if (rem > 0) {
        // strided passes: tile of kx bits = c contiguous low bits (rows of 2^c amplitudes) + m gate bits
        const int kx = tile_bits_x > 0 ? std::min(n, tile_bits_x) : k;
        const int umax = std::max(kx - std::min(min_row_bits, kx - 1), 1);
        const int nx = (rem + umax - 1) / umax;
        int h = k;
        for (int i = 0; i < nx; ++i) {
            const int m = rem / nx + (i < rem % nx ? 1 : 0);
            PassPlan& pp = lp->pass[np++];
            pp.k = kx; pp.c = kx - m; pp.h = h; pp.lean = false; pp.ngroups = 0; pp.gx = false; pp.zmask = 0;
            pp.m1 = m; pp.h2 = h + m;
            if (lean_k) plan_lean(pp, pp.c); else plan_rounds(pp, pp.c, R);
            h += m;
        }
    }
    lp->npasses = np;
    for (int i = 0; i < np; ++i)
        if (!lp->pass[i].lean && (lp->pass[i].nrounds > QR_MAXROUNDS || lp->pass[i].nrounds * R > QR_GATE_SLOTS))
            return fail(QR_EINVAL, "internal: pass needs %d rounds", lp->pass[i].nrounds);
    return 0;
}


// ---- axis-aware plans (QR_OPT_AXIS_PLAN) ----------------------------------------------------------------------------
// The static plan gives every index bit a tile bit in some pass.  A diagonal gate does not need one: k_tile12_g applies
// an Rz whose index bit lies outside the tile as a phase factor of the tile and takes its gradient from the tile's total
// of Im(conj(lambda) psi).  So the strided passes of a layer only need tile bits for the layer's X / Y rotations above
// the contiguous tile: with k of them in a pass the tile keeps 2^(K-k) contiguous amplitudes per row (K = 12, k = 6:
// 1 KiB rows and ONE shared-memory exchange instead of 128 B rows and two), and a register with few high bits may need
// fewer passes.  Pass 0 (contiguous tile, ladder gather) is unchanged.
// Estimated cost of a strided pass with k X / Y gate bits (two-vector pass at n = 30, ms; measured, profiles/README.md):
#define QR_AXIS_MAX_ROW_BITS 7   // rows of 2 KiB; 8 KiB rows (c = 9, no exchange at all) ran at 15 ms against 13.3 ms for c = 6..7
static double axis_pass_cost(int k, int K) {
    // measured back to back at n = 30 (QR_TRACE_PASSES, profiles/r2_axis_trace.log): one exchange round costs ~3.4 ms in
    // the two-vector pass whatever the row width -- k <= 6: 13.2-14.0 ms; k = 7, 8 (256 / 512 B rows): 17.0-17.5 ms;
    // k = 9 (128 B rows, staged loads like the static plan): 15.7-16.9 ms
    (void)K;
    return k <= 6 ? 13.5 : (k <= 8 ? 17.2 : 20.0);   // k = 9: 128 B rows on 9 scattered high bits ran at 18-22 ms
}

// cheapest composition of k X / Y bits into m parts of at most maxk (parts descending in split[]); 1e30 if none
static double axis_best_split(int k, int m, int maxk, int K, int* split_out) {
    double best = 1e30;
    int split[8];
    std::function<void(int, int, double)> rec = [&](int i, int left, double cost) {
        if (i == m - 1) {
            if (left > maxk) return;
            split[i] = left;
            const double cst = cost + axis_pass_cost(left, K);
            if (cst < best) { best = cst; for (int j = 0; j < m; ++j) split_out[j] = split[j]; }
            return;
        }
        for (int a = std::min(left, maxk); a >= 0; --a) {
            if ((m - 1 - i) * maxk < left - a) break;
            split[i] = a;
            rec(i + 1, left - a, cost + axis_pass_cost(a, K));
        }
    };
    rec(0, k, 0.0);
    if (best < 1e29) std::sort(split_out, split_out + m, [](int a, int b) { return a > b; });
    return best;
}

// Strided axis-aware passes for the X / Y bits nz[0..k) (ascending) and the Rz bits zs[0..nzs) of a register with nloc
// local index bits: m passes of split[i] X / Y bits each; fillers are taken from [fill_lo, nloc).  false: the Rz capacity
// (fillers + slots of the row bits) does not suffice.
static bool axis_build_strided(int K, int nloc, int fill_lo, const int* nz, int k, const int* zs, int nzs, int m, const int* split, PassPlan* out) {
    (void)k;
    const int H = nloc - fill_lo;
    bool used_z[64] = {};
    int nzpos = 0, tbits[8][QR_MAX_TILE_BITS], tgate[8][QR_MAX_TILE_BITS], tn[8], cs[8];
    for (int i = 0; i < m; ++i) {
        cs[i] = std::max(std::min(std::min(K - 3, QR_AXIS_MAX_ROW_BITS), K - split[i]), K - H);   // small registers: the tile takes every high bit
        if (cs[i] > K - 3 || cs[i] > fill_lo) return false;
        tn[i] = 0;
        for (int j = 0; j < split[i]; ++j) { tgate[i][tn[i]] = 1; tbits[i][tn[i]++] = nz[nzpos++]; }
    }
    // fillers (tiles with few X / Y bits): unassigned Rz bits first (applied in the tile), then any other high bit
    for (int i = 0; i < m; ++i) {
        const int want = K - cs[i];
        for (int z = 0; z < nzs && tn[i] < want; ++z)
            if (!used_z[z] && zs[z] >= fill_lo) { used_z[z] = true; tgate[i][tn[i]] = 1; tbits[i][tn[i]++] = zs[z]; }
        for (int b = fill_lo; b < nloc && tn[i] < want; ++b) {
            bool in = false;
            for (int j = 0; j < tn[i]; ++j) in = in || tbits[i][j] == b;
            if (!in) { tgate[i][tn[i]] = 0; tbits[i][tn[i]++] = b; }
        }
        if (tn[i] != want) return false;
    }
    // remaining Rz bits -> the slots of the row bits, spread evenly (the bit must lie outside the pass's tile)
    int zbits[8][QR_GX_ZSLOTS], zn[8];
    for (int i = 0; i < m; ++i) zn[i] = 0;
    for (int z = 0; z < nzs; ++z) {
        if (used_z[z]) continue;
        int at = -1;
        for (int i = 0; i < m; ++i) {
            bool in = zs[z] < cs[i];
            for (int j = 0; j < tn[i]; ++j) in = in || tbits[i][j] == zs[z];
            if (!in && zn[i] < std::min(cs[i], QR_GX_ZSLOTS) && (at < 0 || zn[i] < zn[at])) at = i;
        }
        if (at < 0) return false;
        zbits[at][zn[at]++] = zs[z];
    }
    for (int i = 0; i < m; ++i) {
        PassPlan& pp = out[i];
        pp = PassPlan();
        pp.k = K; pp.c = cs[i]; pp.h = K; pp.m1 = K - cs[i]; pp.h2 = pp.h + pp.m1;
        pp.lean = true; pp.gx = true; pp.zmask = 0;
        for (int j = 0; j < QR_MAXROUNDS; ++j) pp.g[j] = 0;
        for (int j = 0; j < QR_GATE_SLOTS; ++j) pp.gbit[j] = -1;
        for (int b = 0; b < QR_MAX_TILE_BITS; ++b) pp.lbit[b] = b;
        // local order above the rows: the bits without an X / Y gate first (they need no round), then the X / Y bits
        const int ki = split[i];
        int lb = cs[i];
        for (int j = ki; j < tn[i]; ++j, ++lb) {
            pp.lbit[lb] = tbits[i][j];
            if (tgate[i][j]) pp.gbit[lb] = tbits[i][j];
        }
        const int first = ki > 0 ? lb : K;
        for (int j = 0; j < ki; ++j, ++lb) { pp.lbit[lb] = tbits[i][j]; pp.gbit[lb] = tbits[i][j]; }
        for (int j = 0; j < zn[i]; ++j) { pp.gbit[j] = zbits[i][j]; pp.zmask |= 1u << j; }
        pp.ngroups = first < 3 ? 4 : (first < 6 ? 3 : (first < K - 3 ? 2 : 1));
        pp.nrounds = pp.ngroups;
        pp.g[0] = K - 3;
    }
    return true;
}

// plan of one layer from its axes (ax[q], qubit q <-> index bit n-1-q); false: keep the static plan.
// Strided passes: tile = rows + the X / Y bits given to the pass (+ fillers); the Rz gates of index bits outside the
// tile go to the slots of the row bits.  Pass 0 keeps the contiguous tile unless a strided pass would need a second
// exchange round for one or two bits: then pass 0 trades Rz-only bits of [7, K) (never row bits of a strided tile) for
// the lowest high X / Y bits (allow_absorb); its ladder gather works on any tile by linearity of the ladder map.
static bool plan_axis_layer(const LayerPlan& base, const int32_t* ax, LayerPlan* out, int allow_absorb) {
    const int n = base.n, K = base.k;
    if (base.npasses < 2 || !base.pass[0].lean || (K != 12 && K != 11)) return false;
    const int H = n - K, maxk = K - 3, maxm = base.npasses - 1;
    if (H < 3) return false;   // every strided tile has at least 3 high bits (the register bits at load time)
    int nz[64], zs[64], k = 0, nzs = 0;
    for (int b = K; b < n; ++b) {
        if (ax[n - 1 - b] == 2) zs[nzs++] = b; else nz[k++] = b;
    }
    int lowz[16], nlowz = 0;   // Rz-only bits of pass 0 that it may trade away, highest first
    for (int b = K - 1; b >= QR_AXIS_MAX_ROW_BITS && allow_absorb; --b)
        if (ax[n - 1 - b] == 2) lowz[nlowz++] = b;
    // The contiguous pass runs the chain L | 0 | 3 (two exchanges instead of three, ~2.5 ms) when the tile positions 6-8
    // can be given to bits without an X / Y gate, i.e. when at most 4 of its bits above bit 4 carry one (allow_absorb).
    int cnt5 = 0;
    for (int b = 5; b < K; ++b) cnt5 += ax[n - 1 - b] != 2 ? 1 : 0;
    auto three_rounds = [&](int e) { return allow_absorb && K == 12 && cnt5 + e <= 4; };
    // choose the number of absorbed bits e and the pass count m by estimated cost
    double best_total = 1e30;
    int best_e = -1, best_m = 0, best_split[8];
    for (int e = 0; e <= std::min(nlowz, k); ++e)
        for (int m = 1; m <= maxm; ++m) {
            if (m * maxk < k - e) continue;
            int split[8];
            const double cst = axis_best_split(k - e, m, maxk, K, split) + (e ? 0.6 + 0.1 * e : 0.0) - (three_rounds(e) ? 2.5 : 0.0);
            if (cst < best_total - 1e-9) { best_total = cst; best_e = e; best_m = m; for (int j = 0; j < m; ++j) best_split[j] = split[j]; }
        }
    if (best_e < 0 || best_total >= 16.0 * maxm) return false;   // the static plan's strided passes run at ~16 ms each (layers that keep it also keep the four-round contiguous pass)
    // try the chosen (e, m); when the Rz capacity does not suffice fall back to e = 0 with growing m
    for (int attempt = 0; attempt < 1 + maxm; ++attempt) {
        int e = best_e, m = best_m, splitv[8];
        if (attempt == 0) { for (int j = 0; j < m; ++j) splitv[j] = best_split[j]; }
        else {
            e = 0; m = attempt;
            if (m * maxk < k || axis_best_split(k, m, maxk, K, splitv) > 1e29) continue;
        }
        int zall[64], nzall = 0;   // Rz bits the strided passes take care of: the high ones and those pass 0 traded away
        for (int z = 0; z < nzs; ++z) zall[nzall++] = zs[z];
        for (int j = 0; j < e; ++j) zall[nzall++] = lowz[j];
        PassPlan strided[8];
        if (!axis_build_strided(K, n, K, nz + e, k - e, zall, nzall, m, splitv, strided)) continue;   // nz[0..e) go to pass 0
        *out = base;
        out->npasses = 1 + m;
        for (int i = 0; i < m; ++i) out->pass[1 + i] = strided[i];
        if (e > 0 || three_rounds(e)) {   // pass 0 on a general tile: bits [0, K) without the traded Rz bits, plus the e lowest high X / Y bits
            PassPlan& p0 = out->pass[0];
            p0.gx = true; p0.zmask = 0;
            u64 traded = 0;
            for (int j = 0; j < e; ++j) traded |= (u64)1 << lowz[j];
            int lb = 0;
            if (!three_rounds(e)) {
                for (int b = 0; b < K; ++b)
                    if (!((traded >> b) & 1)) { p0.lbit[lb] = b; p0.gbit[lb] = b; ++lb; }
                for (int j = 0; j < e; ++j, ++lb) { p0.lbit[lb] = nz[j]; p0.gbit[lb] = nz[j]; }
                p0.ngroups = 4; p0.nrounds = 4;
            } else {
                // local 0-4 = index bits 0-4; the X / Y bits above go to the positions 9, 10, 11, 5 (register groups L and 3),
                // Rz-only bits to 6, 7, 8 and to what is left
                int xy[8], nxy = 0, zz[8], nzz = 0;
                for (int b = 5; b < K; ++b) {
                    if ((traded >> b) & 1) continue;
                    if (ax[n - 1 - b] != 2) xy[nxy++] = b; else zz[nzz++] = b;
                }
                for (int j = 0; j < e; ++j) xy[nxy++] = nz[j];
                for (int b = 0; b < 5; ++b) p0.lbit[b] = b;
                const int xy_pos[4] = {9, 10, 11, 5};
                int pos_used[QR_MAX_TILE_BITS] = {1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
                for (int j = 0; j < nxy; ++j) { p0.lbit[xy_pos[j]] = xy[j]; pos_used[xy_pos[j]] = 1; }
                int zi = 0;
                for (int pos = 6; pos < K; ++pos) {           // 6, 7, 8 first, then the free ones of 9-11
                    if (pos_used[pos]) continue;
                    p0.lbit[pos] = zz[zi++]; pos_used[pos] = 1;
                }
                if (!pos_used[5]) p0.lbit[5] = zz[zi++];
                if (zi != nzz || nxy > 4) return false;   // cannot happen: 7 positions, cnt5 + e <= 4
                for (int b = 0; b < K; ++b) p0.gbit[b] = p0.lbit[b];
                p0.ngroups = 6; p0.nrounds = 3;
            }
            p0.c = 0;
            while (p0.c < K && p0.lbit[p0.c] == p0.c) ++p0.c;   // leading index bits in place
        }
        return true;
    }
    return false;
}

// global index bits of a tile-local index in a gx pass
static u64 gx_local(const PassPlan& pp, u64 l) {
    u64 m = 0;
    for (int b = 0; b < pp.k; ++b)
        if ((l >> b) & 1) m |= (u64)1 << pp.lbit[b];
    return m;
}

// gate table entries of one (layer, pass): QR_MAXROUNDS*QR_R GateP, from per-qubit (axis, cos, sin)
template <class F>
static void fill_gates(const LayerPlan& lp, int pass, GateP* out, F gate_of_qubit) {
    const PassPlan& pp = lp.pass[pass];
    for (int i = 0; i < QR_GATE_SLOTS; ++i) {
        GateP g;
        g.c = 1.0; g.s = 0.0; g.axis = -1; g.pad = 0;
        if (pp.gbit[i] >= 0) g = gate_of_qubit(lp.n - 1 - pp.gbit[i]);
        if (pp.gx && ((pp.zmask >> i) & 1)) g.pad = 1 + pp.gbit[i];   // Rz applied through the tile's own index bit
        out[i] = g;
    }
}

// every pass runs k_tile12 on tiles of the same size => every backward pass of a gradient uses the same grid
// (precondition of the deferred reduction, which reads the same number of per-CTA partials for every pass)
static bool uniform_lean_plan(const LayerPlan& lp) {
    for (int i = 0; i < lp.npasses; ++i)
        if (!lp.pass[i].lean || lp.pass[i].k != lp.pass[0].k) return false;
    return true;
}

typedef void (*tile_fn)(const TilePass);

struct PassIO {
    const double2* src0; const double2* src1; double2* dst0; double2* dst1;
};

struct LadderSpec { u64 M1, M2, src_xor; };   // explicit gather map (sharded states)

// optional launch parameters (sliced passes of sharded registers: qr_shard.cuh)
struct PassExtra {
    int hole = 0;                   // index bits [9, 9+hole) are fixed to tile_or >> 9 instead of enumerated
    u64 tile_or = 0;
    int max_sms = 0;                // > 0: use at most this many SMs (a concurrent pass runs on the others)
    double* partials = nullptr;     // per-CTA partials region (default: d_scratch)
    unsigned* counter = nullptr;    // arrival counter of the fused final reduction (default: d_counter)
    cudaStream_t stream = nullptr;  // default: the context's stream
    const short* hidx = nullptr;    // integer Hamiltonian index table of the pass's layout (default: d_hidx)
};

// opt in to > 48 KiB of dynamic shared memory: the attribute is PER DEVICE, so remember it per (kernel, device)
static int ensure_smem_attr(qr_ctx* c, const void* fn, int slot) {
    static bool done[64][44] = {};   // [device][kernel slot]
    const int dev = c->device & 63;
    if (slot < 0 || slot >= 44) return fail(QR_EINVAL, "internal: kernel slot %d", slot);
    if (!done[dev][slot]) {
        CUDA_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * (int)sizeof(double2) << QR_MAX_TILE_BITS));
        done[dev][slot] = true;
    }
    return 0;
}

// launch one tile pass; returns the number of partial units written (backward) through *units
static int launch_pass(qr_ctx* c, const LayerPlan& lp, int pass, int nv, const PassIO& io, const GateP* d_gates,
                       int gate_stride, int ladder_stacking /* -1 none, else gather of ladder(stacking) */,
                       i64 batch, i64 state_stride, int flush_per_tile, const double* ham, int pre_phase,
                       double angle_pre, int post_phase, double angle_post, int* units, const LadderSpec* spec = nullptr,
                       double* final_out = nullptr, const double2* lut = nullptr, double* partials_at = nullptr,
                       const PassExtra* ex = nullptr) {
    const PassPlan& pp = lp.pass[pass];
    cudaStream_t stream = (ex && ex->stream) ? ex->stream : c->stream;
    const int sms = (ex && ex->max_sms > 0) ? std::min(ex->max_sms, c->sm_count) : c->sm_count;
    TilePass tp;
    memset(&tp, 0, sizeof(tp));
    tp.k = pp.k; tp.c = pp.c; tp.h = pp.h; tp.nrounds = pp.nrounds;
    tp.m1 = pp.m1; tp.h2 = pp.h2;
    for (int r = 0; r < QR_MAXROUNDS; ++r) tp.g[r] = r < pp.nrounds ? pp.g[r] : 0;
    tp.ladder = 0;
    if (spec) {
        tp.ladder = 1;
        tp.M1 = spec->M1; tp.M2 = spec->M2; tp.src_xor = spec->src_xor;
    } else if (ladder_stacking >= 0) {
        tp.ladder = 1;
        ladder_masks(lp.n, 1 - ladder_stacking, &tp.M1, &tp.M2);
    }
    tp.hole = ex ? ex->hole : 0;
    tp.tile_or = ex ? ex->tile_or : 0;
    tp.tiles_log2 = lp.n - pp.k - tp.hole;
    tp.num_tiles = batch << tp.tiles_log2;
    tp.state_stride = state_stride;
    tp.src0 = io.src0; tp.src1 = io.src1; tp.dst0 = io.dst0; tp.dst1 = io.dst1;
    tp.gates = d_gates; tp.gate_stride = gate_stride;
    tp.ham = ham; tp.pre_phase = pre_phase; tp.post_phase = post_phase;
    if (lut && c->ham_integer) { tp.hidx = (ex && ex->hidx) ? ex->hidx : c->d_hidx; tp.lut = lut; tp.lut_size = c->ham_range; tp.hmin = c->ham_min; }
    tp.angle_pre = angle_pre; tp.angle_post = angle_post;
    tp.flush_per_tile = flush_per_tile;
    tp.prefetch = nv == 2 ? (int)(c->opt_prefetch & 3) : (int)((c->opt_prefetch >> 2) & 3);   // tiles ahead: bits 0-1 backward, bits 2-3 forward
    const int R = lp.R;
    const size_t tile_bytes = sizeof(double2) << pp.k;
    tp.final_out = (nv == 2 && !flush_per_tile) ? final_out : nullptr;
    tp.done_counter = (ex && ex->counter) ? ex->counter : c->d_counter;
    if (ex && ex->partials) partials_at = ex->partials;
    if (pp.lean) {   // k_tile12 (k = 12: 512 threads; k = 11: 256 threads)
        typedef void (*lean_fn)(const TilePass, const Tile12X);
        const int ph = (pre_phase || post_phase) ? 1 : 0;
        // staged (asynchronous shared-memory copies of the next tile) vs direct loads + L2 prefetch, per pass:
        // measured at n = 30 (profiles/README.md) the L2 prefetch wins for the contiguous pass and for strides
        // below 32 MiB, and loses badly (22 vs 17 ms) when all gate bits are >= 21 -> auto mode (bit 2).
        int staged = (nv == 2 ? (c->opt_staged & 1) : (c->opt_staged & 2)) ? 1 : 0;
        {   // auto: strided backward passes most of whose gate bits are high-stride bits (two-run tiles: count them)
            int nbits = 0, nhigh = 0;
            for (int sb = pp.gx ? pp.c : 0; sb < QR_GATE_SLOTS; ++sb)
                if (pp.gbit[sb] >= 0) { ++nbits; if (pp.gbit[sb] >= c->opt_staged_min_bit) ++nhigh; }
            const bool single_run = !pp.gx && pp.h2 == pp.h + pp.m1;
            const bool high = single_run ? pp.h >= c->opt_staged_min_bit : (nbits > 0 && 2 * nhigh > nbits);
            if (nv == 2 && (c->opt_staged & 4) && pp.c < pp.k && high) staged = 1;
        }
        const int K = pp.k;   // 12, or 11 (half-size tiles, direct loads only)
        if (pp.gx) staged = 0;   // staged loads measured at 20-25 ms per backward pass on these tiles (13-17.5 ms direct)
        if (K == 11) staged = 0;
        lean_fn lfn;
        if (pp.gx) {
            if (ph || tp.hole || spec) return fail(QR_ESTATE, "internal: axis-aware pass with a phase, a slice or a sharded gather");
            if (K == 11) lfn = nv == 1 ? k_tile12_g<1, 0, 11> : k_tile12_g<2, 0, 11>;
            else lfn = nv == 1 ? k_tile12_g<1, 0> : k_tile12_g<2, 0>;
        } else if (K == 11) lfn = nv == 1 ? (ph ? k_tile12<1, true, 0, 11> : k_tile12<1, false, 0, 11>) : (ph ? k_tile12<2, true, 0, 11> : k_tile12<2, false, 0, 11>);
        else if (staged) lfn = nv == 1 ? (ph ? k_tile12<1, true, 1> : k_tile12<1, false, 1>) : (ph ? k_tile12<2, true, 1> : k_tile12<2, false, 1>);
        else lfn = nv == 1 ? (ph ? k_tile12<1, true, 0> : k_tile12<1, false, 0>) : (ph ? k_tile12<2, true, 0> : k_tile12<2, false, 0>);
        QR_TRY(ensure_smem_attr(c, (const void*)lfn, pp.gx ? 32 + ((K - 11) * 2 + (nv - 1)) * 2 : ((K - 11) * 2 + (nv - 1)) * 6 + ph * 3 + staged));
        if (staged == 1) tp.prefetch = 0;
        Tile12X x;
        memset(&x, 0, sizeof(x));
        x.ngroups = pp.ngroups;
        if (pp.gx && (c->opt_axis_plan & 4) && x.ngroups == 1 && K == 12) x.ngroups = 2;   // experiment: keep the warps of a tile together
        if (pp.gx && (c->opt_axis_plan & 8) && x.ngroups == 2 && K == 12 && !staged) {
            // the halves of the CTA synchronise separately (measured 13.3-14.2 -> 12.2-12.8 ms per backward pass at n = 30)
            lfn = nv == 1 ? k_tile12_gs<1> : k_tile12_gs<2>;
            QR_TRY(ensure_smem_attr(c, (const void*)lfn, 40 + (nv - 1)));
        }
        const Geo12 geo = {pp.c, pp.h, pp.m1, pp.h2, K, 0};
        x.last_group = x.ngroups == 1 ? K - 3 : (x.ngroups == 5 ? 5 : (x.ngroups == 6 ? 3 : 6));
        for (int r = 0; r < 8; ++r) {
            const u64 lf = (u64)r << (K - 3);
            const u64 ll = (u64)r << x.last_group;
            x.droff_first[r] = pp.gx ? gx_local(pp, lf) : geo12_local(geo, lf);
            x.roff_first[r] = tp.ladder ? ladder_map(x.droff_first[r], tp.M1, tp.M2) : x.droff_first[r];
            x.roff_last[r] = pp.gx ? gx_local(pp, ll) : geo12_local(geo, ll);
        }
        if (pp.gx) {   // general geometry: index bit of every local bit, and of every tile-index bit (the rest, ascending)
            u64 in_tile = 0;
            for (int b = 0; b < K; ++b) { x.lpos[b] = (unsigned char)pp.lbit[b]; in_tile |= (u64)1 << pp.lbit[b]; }
            int j = 0;
            for (int b = 0; b < lp.n; ++b)
                if (!((in_tile >> b) & 1)) x.tpos[j++] = (unsigned char)b;
            if (j != lp.n - K || j > 24) return fail(QR_ESTATE, "internal: axis-aware tile geometry");
            for (; j < 24; ++j) x.tpos[j] = 63;   // never selected: tile indices have n - K bits
        }
        const long long lctas = K == 11 ? (nv == 1 ? 4 : 2) : ((nv == 1 && !staged) ? std::min<long long>(2, c->opt_ctas_fwd) : 1);
        const i64 lgrid = std::min<i64>(tp.num_tiles, (i64)sms * lctas);
        if (nv == 2) {
            const i64 nunits = flush_per_tile ? tp.num_tiles : lgrid;
            QR_TRY(ensure_scratch(c, (size_t)nunits * QR_SLOTS));
            *units = (int)nunits;
        }
        tp.partials = partials_at ? partials_at : c->d_scratch;   // partials_at: a slice of d_scratch the caller has sized (deferred reduction)
        const bool strided_pass = pp.c < K;
        // L2 prefetch of the next tile: opt_prefetch bit 4 = contiguous passes only
        if ((c->opt_prefetch & 16) && strided_pass) tp.prefetch = 0;
        // axis-aware strided passes: the prefetch buys nothing (12.2-13.2 ms with, 12.2-13.6 ms without) and costs 11-15 %
        // more DRAM reads (38.2-39.5 GB per launch against 34.36 GB: lines evicted again before their tile is loaded)
        if (pp.gx && pp.ngroups < 4) tp.prefetch = 0;
        const size_t lsmem = staged ? (size_t)(nv + 1) * tile_bytes : (x.ngroups > 1 ? (size_t)nv * tile_bytes : 0);
        // programmatic dependent launch: the next pass's CTAs queue up while this one drains.  Auto (1): only where a
        // pass is short enough for the launch ramp to matter (states that fit in L2); 2: every pass.
        // The first pass after the gate / phase tables were written is launched fully serialized: k_tile12 reads the
        // tables BEFORE griddepcontrol.wait, which is only safe once a serialized launch separates it from their writer.
        const bool pdl = (c->opt_pdl == 2 || (c->opt_pdl == 1 && lp.n <= QR_PDL_AUTO_MAX_QUBITS)) && !c->tables_fresh;
        c->tables_fresh = false;
        if (pdl) {
            CUDA_TRY(QR_LAUNCH_EX(lfn, (unsigned)lgrid, 1u << (K - 3), lsmem, stream, 1u, pdl, tp, x));
        } else {
            QR_LAUNCH(lfn, (unsigned)lgrid, 1 << (K - 3), lsmem, stream, tp, x);
        }
        KERNEL_CHECK();
        c->perf.kernel_launches++;
        return 0;
    }
    // generic kernel (tiles of fewer than 11 bits)
    const int threads = 1 << (pp.k - R);
    const int full = 1 << (QR_MAX_TILE_BITS - R);
    const long long ctas = nv == 1 ? c->opt_ctas_fwd : c->opt_ctas_bwd;
    const long long per_sm = std::min<long long>(16, ctas * std::max(1, full / threads));
    const i64 grid = std::min<i64>(tp.num_tiles, (i64)c->sm_count * per_sm);
    const size_t smem = pp.nrounds > 1 ? (size_t)nv * tile_bytes : 0;
    if (nv == 2) {
        const i64 nunits = flush_per_tile ? tp.num_tiles : grid;
        QR_TRY(ensure_scratch(c, (size_t)nunits * QR_SLOTS));
        *units = (int)nunits;
    }
    tp.partials = partials_at ? partials_at : c->d_scratch;
    tile_fn fn = nv == 1 ? k_tile_pass<1, 3> : k_tile_pass<2, 3>;
    QR_TRY(ensure_smem_attr(c, (const void*)fn, 24 + nv));
    QR_LAUNCH(fn, (unsigned)grid, threads, smem, c->stream, tp);
    KERNEL_CHECK();
    c->perf.kernel_launches++;
    return 0;
}

// ------------------------------------------------------------------------------------------
// McClean (circuit_logic/mc_clean.py)
// ------------------------------------------------------------------------------------------
static int check_mcclean_args(qr_ctx* c, int L, const int32_t* axes, const double* angles, const qr_obs* o) {
    QR_TRY(check_obs(c, o));
    if (L < 0) return fail(QR_EINVAL, "layer_number must be >= 0");
    if (L > 0 && (!axes || !angles)) return fail(QR_EINVAL, "null axes/angles");
    for (i64 i = 0; i < (i64)L * c->n; ++i)
        if (axes[i] < 0 || axes[i] > 2) return fail(QR_EINVAL, "Invalid axis %d", axes[i]);   // mc_clean.py:392
    return 0;
}

static void perf_reset(qr_ctx* c) { memset(&c->perf, 0, sizeof(c->perf)); }

// product state prod_q Ry_q(pi/4)|0>: amplitude cos(pi/8)^(n-w) sin(pi/8)^w, w = popcount(j)
static int init_mcclean_product(qr_ctx* c) {
    double table[40];
    const double cs = std::cos(M_PI / 8.0), sn = std::sin(M_PI / 8.0);
    for (int w = 0; w <= c->n; ++w) {
        double v = 1.0;
        for (int q = 0; q < c->n; ++q) v *= (q < c->n - w) ? cs : sn;
        table[w] = v;
    }
    const size_t off = c->small_cap - 512;   // tail of d_small, away from terms / gate tables
    QR_TRY(upload_small(c, off, table, sizeof(double) * (c->n + 1), c->pin_cap - 512));
    QR_LAUNCH(k_init_product, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N,
              (const double*)((char*)c->d_small + off), c->N - 1);
    KERNEL_CHECK();
    c->perf.kernel_launches++;
    return 0;
}

// --- un-fused McClean (gate-at-a-time kernels): QR_OPT_FUSION=0, n < 4, or cross-checks ---
// init_mode (McClean: 0 / 1; layered circuits, qr_layered_grad: 2 / 3):
//   0  Ry(pi/4)^n |0..0>            1  Ry(pi/4)^n applied to the current state (ini_state)
//   2  |0..0>, no Ry layer          3  the current state, no Ry layer
// lad_flags (null = every layer): layer i starts with the CNOT ladder iff lad_flags[i] != 0
static int mcclean_unfused(qr_ctx* c, int L, const int32_t* axes, const double* angles, const qr_obs* o,
                           int use_current, double* e_out, double* grad, const unsigned char* lad_flags = nullptr) {
    const int n = c->n;
    auto has_ladder = [&](int i) { return n >= 2 && (!lad_flags || lad_flags[i]); };
    if (use_current == 0 || use_current == 2) {
        QR_LAUNCH(k_init_basis, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N, 0, 0.0);
        KERNEL_CHECK();
    }
    if (use_current < 2)
        for (int q = 0; q < n; ++q) QR_TRY(launch_1q(c, c->buf[c->psi], n - 1 - q, rot_matrix(1, M_PI / 4.0)));
    for (int i = 0; i < L; ++i) {
        if (has_ladder(i)) {
            const int dst = other_buf(c, c->psi);
            QR_TRY(ensure_buf(c, dst));
            QR_TRY(launch_ladder(c, c->psi, dst, 0));
            c->psi = dst;
        }
        for (int q = 0; q < n; ++q)
            QR_TRY(launch_1q(c, c->buf[c->psi], n - 1 - q, rot_matrix(axes[i * n + q], angles[i * n + q])));
    }
    if (!grad) return observable_pass(c, o, c->psi, -1, e_out);
    int lam = other_buf(c, c->psi);
    QR_TRY(ensure_buf(c, lam));
    QR_TRY(observable_pass(c, o, c->psi, lam, e_out));
    const u64 npairs = c->N >> 1;
    const int grid = grid_for(c, npairs);
    QR_TRY(ensure_scratch(c, grid));
    for (int i = L - 1; i >= 0; --i) {
        for (int q = 0; q < n; ++q) {
            QR_LAUNCH(k_pauli_inner, grid, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi],
                      (const double2*)c->buf[lam], npairs, n - 1 - q, (int)axes[i * n + q], c->d_scratch);
            KERNEL_CHECK();
            QR_TRY(reduce_to_host(c, grid, 1, &grad[i * n + q]));
        }
        for (int q = 0; q < n; ++q) {
            const Mat2 m = rot_matrix(axes[i * n + q], -angles[i * n + q]);
            QR_TRY(launch_1q(c, c->buf[c->psi], n - 1 - q, m));
            QR_TRY(launch_1q(c, c->buf[lam], n - 1 - q, m));
        }
        if (has_ladder(i)) {
            const int d1 = other_buf(c, c->psi, lam);
            const int d2 = other_buf(c, c->psi, lam, d1);
            QR_TRY(ensure_buf(c, d1));
            QR_TRY(launch_ladder(c, lam, d1, 1));
            lam = d1;
            if (i > 0) {   // psi is only needed while gradients remain
                QR_TRY(ensure_buf(c, d2));
                QR_TRY(launch_ladder(c, c->psi, d2, 1));
                c->psi = d2;
            }
        }
    }
    c->psi = lam;   // mc_clean.py leaves the back-propagated co-state in state.vec
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

struct GradLayout {   // where the per-(layer, pass) slot sums land in d_result
    int slots_per_layer;
};

// fused McClean forward (+ backward when grad != null) for `batch` parameter sets laid out
// contiguously in the state buffers (state b at offset b * N)
// Device-resident parameters (optimiser loop, qr_mcclean_optimize): the angles are read from device memory, the gate
// tables are built on the device, and nothing is read back -- E and the slot sums stay in d_result
// ([E][L][P][QR_SLOTS]) for the update kernel that is enqueued next.
struct DevParams {
    const double* d_angles;      // [L * n] on the device
    int* slot_qubit_out;         // host [L * P * QR_GATE_SLOTS]: qubit of each backward gradient slot (or -1), per layer
    int* passes_out;             // host: P
};

static int mcclean_fused(qr_ctx* c, i64 batch, int L, const int32_t* axes, const double* angles, const qr_obs* o,
                         int use_current, double* e_out, double* grad, const DevParams* dev = nullptr,
                         const unsigned char* lad_flags = nullptr) {
    const int n = c->n;
    c->tables_fresh = true;
    LayerPlan lpf, lp;   // forward / backward plans: same tile geometry, different register blocking
    // axis-aware plans apply to single circuits whose axes the host knows; with them 12-bit tiles win from n = 23 on
    // (measured, profiles/r2_axis_ab.log: n = 24: 30.1 -> 25.3 ms per 30 layers, n = 26: 127 -> 118 ms; n = 22: 4.45 vs 4.57 ms)
    const bool axis_ok = c->opt_axis_plan && batch == 1 && (n >= 20 || c->axis_plan_forced);
    int tile_bits = pick_tile_bits(c, n);
    if (axis_ok && c->opt_tile_bits == 0 && n >= 23 && n <= 26 && tile_bits == 11) tile_bits = 12;
    QR_TRY(make_plan(n, tile_bits, &lpf, (int)c->opt_tile_bits_x, tile_bits == pick_tile_bits(c, n) ? pick_min_row_bits(c, n) : (int)c->opt_min_row_bits));
    lp = lpf;
    const int P = lp.npasses;
    const int GS = QR_GATE_SLOTS;                 // gate entries per (layer, pass)
    const bool want_grad = grad != nullptr;
    const bool ry_layer = use_current == 1;       // ini_state given: apply the Ry(pi/4) layer as gates (init modes: mcclean_unfused)
    auto has_ladder = [&](int i) { return n >= 2 && (!lad_flags || lad_flags[i]); };
    // ---- gate tables: [batch][ (ry layer) + L forward + L backward ][P][GS] ----
    const int nlay_tab = (ry_layer ? 1 : 0) + L + (want_grad ? L : 0);
    const size_t per_batch = (size_t)nlay_tab * P * GS;
    const size_t tab_bytes = (size_t)batch * per_batch * sizeof(GateP);
    const size_t terms_bytes = (o->terms.size() + 1) * sizeof(ObsTerm);
    const size_t tab_off = (terms_bytes + 255) & ~(size_t)255;
    const size_t raw_off = (tab_off + tab_bytes + 1024 + 255) & ~(size_t)255;   // batched: raw axes/angles/qmap staging
    const bool dev_tables = batch > 1 || dev != nullptr;   // build the gate tables on the device from raw axes / angles
    // per-layer plans from the axes (QR_OPT_AXIS_PLAN): pass counts may differ between layers (<= P), the table and
    // result layouts keep P entries per layer
    std::vector<LayerPlan> plans((size_t)L, lpf);
    int n_axis_layers = 0;
    if (axis_ok && uniform_lean_plan(lpf))
        for (int i = 0; i < L; ++i) n_axis_layers += plan_axis_layer(lpf, axes + (size_t)i * n, &plans[i], (c->opt_axis_plan & 2) ? 1 : 0) ? 1 : 0;
    const bool qmap_per_layer = dev != nullptr;   // one circuit: slot -> qubit maps per layer (axis-aware plans)
    const size_t qmap_ints = 2 * (size_t)(qmap_per_layer ? L : 1) * P * GS;
    const size_t raw_bytes = dev_tables ? (size_t)batch * L * n * (sizeof(double) + sizeof(int32_t)) + qmap_ints * sizeof(int) + 1024 : 0;
    QR_TRY(ensure_small(c, raw_off + raw_bytes + 1024));
    QR_TRY(ensure_pin(c, dev_tables ? std::max((size_t)1 << 16, (size_t)batch * (1 + (size_t)L * P * QR_SLOTS) * sizeof(double) + 8192) + tab_off + qmap_ints * sizeof(int) + (dev ? (size_t)L * n * sizeof(int32_t) + 64 : 0)
                                    : tab_off + tab_bytes + 1024));
    if (dev_tables) {
        // device-side table build: upload raw parameters + the slot->qubit maps of both directions
        char* d_raw = (char*)c->d_small + raw_off;
        double* d_angles = (double*)d_raw;
        int* d_axes = (int*)(d_raw + (size_t)batch * L * n * sizeof(double));
        int* d_qmap = d_axes + (size_t)batch * L * n;
        int* qmap = (int*)(c->h_pin + tab_off);   // own staging slot: offset 0 is reused for the observable terms below
        for (int dir = 0; dir < 2; ++dir)
            for (int i = 0; i < (qmap_per_layer ? L : 1); ++i)
                for (int p = 0; p < P; ++p)
                    for (int s2 = 0; s2 < GS; ++s2) {
                        const LayerPlan& pl = qmap_per_layer ? plans[i] : lp;
                        const int gb = p < pl.npasses ? pl.pass[p].gbit[s2] : -1;
                        const int pad = (gb >= 0 && pl.pass[p].gx && ((pl.pass[p].zmask >> s2) & 1)) ? 1 + gb : 0;
                        qmap[(((size_t)dir * (qmap_per_layer ? L : 1) + i) * P + p) * GS + s2] = gb < 0 ? -1 : (n - 1 - gb) + 64 * pad;
                    }
        CUDA_TRY(cudaMemcpyAsync(d_qmap, qmap, qmap_ints * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        if (!dev) CUDA_TRY(cudaMemcpyAsync(d_angles, angles, (size_t)batch * L * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        if (dev) {   // through pinned memory: the step may be captured into a graph (device optimiser loop)
            int32_t* ax_pin = (int32_t*)(c->h_pin + tab_off + qmap_ints * sizeof(int) + (dev ? (size_t)L * n * sizeof(int32_t) + 64 : 0));
            memcpy(ax_pin, axes, (size_t)L * n * sizeof(int32_t));
            CUDA_TRY(cudaMemcpyAsync(d_axes, ax_pin, (size_t)L * n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        } else CUDA_TRY(cudaMemcpyAsync(d_axes, axes, (size_t)batch * L * n * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        const i64 total = (i64)batch * per_batch;
        QR_LAUNCH(k_build_gates, grid_for(c, (u64)total), QR_BLOCK, 0, c->stream, (const int*)d_axes, dev ? dev->d_angles : (const double*)d_angles,
                  (const int*)d_qmap, (GatePOut*)((char*)c->d_small + tab_off), batch, L, n, P, GS, want_grad ? 2 : 1, qmap_per_layer ? 1 : 0);
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    } else {
        GateP* tab = (GateP*)(c->h_pin + tab_off);
        for (i64 b = 0; b < batch; ++b) {
            GateP* tb = tab + b * per_batch;
            int lay = 0;
            if (ry_layer) {
                const double cs = std::cos(M_PI / 8.0), sn = std::sin(M_PI / 8.0);
                for (int p = 0; p < P; ++p)
                    fill_gates(lpf, p, tb + ((size_t)lay * P + p) * GS, [&](int) { GateP g; g.c = cs; g.s = sn; g.axis = 1; g.pad = 0; return g; });
                ++lay;
            }
            for (int dir = 0; dir < (want_grad ? 2 : 1); ++dir)
                for (int i = 0; i < L; ++i, ++lay) {
                    const int32_t* ax = axes + ((size_t)b * L + i) * n;
                    const double* an = angles + ((size_t)b * L + i) * n;
                    const double sgn = dir == 0 ? 1.0 : -1.0;
                    for (int p = 0; p < plans[i].npasses; ++p)
                        fill_gates(plans[i], p, tb + ((size_t)lay * P + p) * GS, [&](int q) {
                            GateP g; g.c = std::cos(0.5 * an[q]); g.s = sgn * std::sin(0.5 * an[q]); g.axis = ax[q]; g.pad = 0; return g; });
                }
        }
        CUDA_TRY(cudaMemcpyAsync((char*)c->d_small + tab_off, tab, tab_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    const GateP* d_tab = (const GateP*)((char*)c->d_small + tab_off);
    const int gate_stride = (int)per_batch;
    const i64 stride = (i64)c->N;
    const int flush = batch > 1 ? 1 : 0;

    // QR_TRACE_PASSES=1 (environment): one CUDA event per tile pass, per-pass device times on stderr after the run
    static const bool trace_env = getenv("QR_TRACE_PASSES") != nullptr;
    const bool trace = trace_env && !c->capturing && dev == nullptr;
    struct TraceRec { cudaEvent_t ev; int layer, pass, nv, c, ng, gx; };
    std::vector<TraceRec> trace_recs;
    auto trace_mark = [&](int layer, int pass, int nv, const PassPlan* pp) {
        if (!trace) return;
        TraceRec r;
        cudaEventCreate(&r.ev);
        cudaEventRecord(r.ev, c->stream);
        r.layer = layer; r.pass = pass; r.nv = nv; r.c = pp ? pp->c : 0; r.ng = pp ? pp->ngroups : 0; r.gx = pp ? (int)pp->gx : 0;
        trace_recs.push_back(r);
    };
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    // ---- forward ----
    int lay = 0;
    if (use_current == 2) {          // layered circuits: |0..0>, no Ry layer
        QR_LAUNCH(k_init_basis, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N, 0, 0.0);
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    } else if (use_current == 3) {   // layered circuits: the current state as it is
    } else if (!ry_layer) {
        if (batch == 1) QR_TRY(init_mcclean_product(c));
        else {
            // same product state for every batch element: N*batch amplitudes, popcount of the low n bits
            double table[40];
            const double cs = std::cos(M_PI / 8.0), sn = std::sin(M_PI / 8.0);
            for (int w = 0; w <= n; ++w) { double v = 1.0; for (int q = 0; q < n; ++q) v *= (q < n - w) ? cs : sn; table[w] = v; }
            const size_t off = c->small_cap - 512;
            QR_TRY(upload_small(c, off, table, sizeof(double) * (n + 1), c->pin_cap - 512));
            QR_LAUNCH(k_init_product, grid_for(c, c->N * batch), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N * (u64)batch,
                      (const double*)((char*)c->d_small + off), c->N - 1);
            KERNEL_CHECK();
            c->perf.kernel_launches++;
        }
    } else {
        for (int p = 0; p < P; ++p) {
            PassIO io = {c->buf[c->psi], nullptr, c->buf[c->psi], nullptr};
            QR_TRY(launch_pass(c, lpf, p, 1, io, d_tab + ((size_t)lay * P + p) * GS, gate_stride, -1, batch, stride, 0,
                               nullptr, 0, 0, 0, 0, nullptr));
        }
        ++lay;
    }
    int n_fwd_pass = ry_layer ? P : 0;
    for (int i = 0; i < L; ++i, ++lay) {
        for (int p = 0; p < plans[i].npasses; ++p, ++n_fwd_pass) {
            int dst = c->psi;
            int lad = -1;
            if (p == 0 && has_ladder(i)) { dst = other_buf(c, c->psi); QR_TRY(ensure_buf(c, dst)); lad = 0; }
            PassIO io = {c->buf[c->psi], nullptr, c->buf[dst], nullptr};
            trace_mark(i, p, 1, &plans[i].pass[p]);
            QR_TRY(launch_pass(c, plans[i], p, 1, io, d_tab + ((size_t)lay * P + p) * GS, gate_stride, lad, batch, stride, 0,
                               nullptr, 0, 0, 0, 0, nullptr));
            c->psi = dst;
        }
    }
    trace_mark(-1, -1, 0, nullptr);
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    // ---- observable: lambda = O psi, E = Re<psi|lambda> ----
    const ObsTerm* d_terms;
    QR_TRY(upload_small(c, 0, o->terms.data(), o->terms.size() * sizeof(ObsTerm), 0));
    d_terms = (const ObsTerm*)c->d_small;
    int lam = -1;
    if (want_grad) { lam = other_buf(c, c->psi); QR_TRY(ensure_buf(c, lam)); }
    const int ogrid = grid_for(c, c->N);
    // deferred second-stage reduction: each backward pass keeps its per-CTA partials in its own slice of d_scratch
    const bool defer = want_grad && batch == 1 && c->opt_defer_reduce && L > 0 && uniform_lean_plan(lp);
    const size_t unit_cap = (size_t)c->sm_count * 16;   // upper bound on the CTAs of a pass (launch_pass: per_sm <= 16)
    QR_TRY(ensure_scratch(c, std::max<size_t>((size_t)ogrid * (size_t)std::max<i64>(batch, 1), defer ? (size_t)L * P * unit_cap * QR_SLOTS : (size_t)0)));
    QR_TRY(ensure_result(c, (size_t)batch * (1 + (want_grad ? (size_t)L * P * QR_SLOTS : 0)) + 16));
    {
        QR_LAUNCH(k_apply_obs, dim3((unsigned)ogrid, (unsigned)batch), QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi],
                  want_grad ? c->buf[lam] : (double2*)nullptr, c->N, d_terms, (int)o->terms.size(), c->d_scratch, (u64)0, 64,
                  PeerTable());
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    }
    // result layout: [batch] E values, then [batch][L][P][QR_SLOTS] slot sums
    QR_LAUNCH(k_reduce_partials_grouped, (unsigned)batch, QR_BLOCK, 0, c->stream, (const double*)c->d_scratch, ogrid, 1,
              c->d_result, 1);
    KERNEL_CHECK();
    c->perf.kernel_launches++;
    CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    // ---- backward ----
    double* d_slots = c->d_result + batch;
    const int slots_per_state = L * P * QR_SLOTS;
    int n_bwd_pass = 0;
    int defer_units = -1;
    if (want_grad) {
        for (int i = L - 1; i >= 0; --i, ++lay) {
            // table index of backward layer i: forward tables come first, backward tables are stored in layer order
            const int tlay = (ry_layer ? 1 : 0) + L + i;
            for (int p = plans[i].npasses; p < P && defer; ++p)   // a pass this layer does not need: its slice of partials reads as zero
                CUDA_TRY(cudaMemsetAsync(c->d_scratch + ((size_t)i * P + p) * unit_cap * QR_SLOTS, 0, unit_cap * QR_SLOTS * sizeof(double), c->stream));
            for (int p = 0; p < plans[i].npasses; ++p) {
                int dpsi = c->psi, dlam = lam, lad = -1;
                if (p == 0 && i < L - 1 && has_ladder(i + 1)) {
                    dpsi = other_buf(c, c->psi, lam);
                    dlam = other_buf(c, c->psi, lam, dpsi);
                    QR_TRY(ensure_buf(c, dpsi));
                    QR_TRY(ensure_buf(c, dlam));
                    lad = 1;   // inverse ladder of layer i+1, folded into this load (mc_clean.py:77)
                }
                PassIO io = {c->buf[c->psi], c->buf[lam], c->buf[dpsi], c->buf[dlam]};
                int units = 0;
                trace_mark(i, p, 2, &plans[i].pass[p]);
                QR_TRY(launch_pass(c, plans[i], p, 2, io, d_tab + ((size_t)tlay * P + p) * GS, gate_stride, lad, batch, stride,
                                   flush, nullptr, 0, 0, 0, 0, &units, nullptr,
                                   (batch == 1 && !defer) ? d_slots + ((size_t)i * P + p) * QR_SLOTS : nullptr, nullptr,
                                   defer ? c->d_scratch + ((size_t)i * P + p) * unit_cap * QR_SLOTS : nullptr));
                if (defer) {
                    if (defer_units >= 0 && units != defer_units) return fail(QR_ESTATE, "internal: passes of one gradient use different grids");
                    defer_units = units;
                }
                c->psi = dpsi;
                lam = dlam;
                ++n_bwd_pass;
                if (batch == 1) continue;   // second-stage reduction is fused into the pass (last CTA)
                {
                    const int tiles_per_state = 1 << (n - plans[i].pass[p].k);
                    QR_LAUNCH(k_reduce_partials_grouped, (unsigned)batch, 32, 0, c->stream, (const double*)c->d_scratch,
                              tiles_per_state, QR_SLOTS, d_slots + ((size_t)i * P + p) * QR_SLOTS, slots_per_state);
                }
                KERNEL_CHECK();
                c->perf.kernel_launches++;
            }
        }
        if (defer) {
            QR_LAUNCH(k_reduce_slots_strided, (unsigned)(L * P), 32 * 16, 0, c->stream, (const double*)c->d_scratch, defer_units,
                      (u64)unit_cap * QR_SLOTS, d_slots, QR_SLOTS);
            KERNEL_CHECK();
            c->perf.kernel_launches++;
        }
        if (c->opt_final_ladder && L > 0 && batch == 1 && has_ladder(0)) {   // mc_clean.py:77 for layer 0
            const int d = other_buf(c, c->psi, lam);
            QR_TRY(ensure_buf(c, d));
            QR_TRY(launch_ladder(c, lam, d, 1));
            c->perf.kernel_launches++;
            lam = d;
        }
    }
    trace_mark(-1, -1, 0, nullptr);
    CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    if (trace) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (size_t t = 0; t + 1 < trace_recs.size(); ++t) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, trace_recs[t].ev, trace_recs[t + 1].ev);
            if (trace_recs[t].nv)
                fprintf(stderr, "[qr trace] layer %d pass %d nv %d %s rows 2^%d groups %d: %.3f ms\n", trace_recs[t].layer, trace_recs[t].pass,
                        trace_recs[t].nv, trace_recs[t].gx ? "axis" : "static", trace_recs[t].c, trace_recs[t].ng, ms);
        }
        for (auto& r : trace_recs) cudaEventDestroy(r.ev);
    }
    if (dev) {   // results stay on the device; tell the caller how to read the slot sums
        for (int i = 0; i < L; ++i)
            for (int p = 0; p < P; ++p)
                for (int s2 = 0; s2 < GS; ++s2) {
                    const int gb = p < plans[i].npasses ? plans[i].pass[p].gbit[s2] : -1;
                    dev->slot_qubit_out[((size_t)i * P + p) * GS + s2] = gb < 0 ? -1 : n - 1 - gb;
                }
        *dev->passes_out = P;
        c->psi = lam;
        return 0;
    }
    // ---- results ----
    const size_t nres = (size_t)batch * (1 + (want_grad ? slots_per_state : 0));
    QR_TRY(ensure_pin(c, nres * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(c->h_pin, c->d_result, nres * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const double* res = (const double*)c->h_pin;
    for (i64 b = 0; b < batch; ++b) e_out[b] = res[b];
    if (want_grad) {
        const double* slots = res + batch;
        // (slot offset within a layer, qubit) of the n gradient slots, once per call: the scatter below runs
        // batch * L * n times (1.6 M for BASELINE config 4)
        int soff[64], sq[64], ns = 0;
        auto slot_map = [&](const LayerPlan& pl) {
            ns = 0;
            for (int p = 0; p < pl.npasses; ++p)
                for (int s = 0; s < GS; ++s) {
                    const int gb = pl.pass[p].gbit[s];
                    if (gb < 0 || ns >= 64) continue;
                    soff[ns] = p * QR_SLOTS + s;
                    sq[ns] = n - 1 - gb;
                    ++ns;
                }
        };
        slot_map(lp);
        for (i64 b = 0; b < batch; ++b)
            for (int i = 0; i < L; ++i) {
                if (n_axis_layers) slot_map(plans[i]);
                const double* src = slots + (size_t)b * slots_per_state + (size_t)i * P * QR_SLOTS;
                double* dst = grad + ((size_t)b * L + i) * n;
                for (int k = 0; k < ns; ++k) dst[sq[k]] = src[soff[k]];
            }
        c->psi = lam;   // state.vec = back-propagated co-state (mc_clean.py:68-77)
    }
    // ---- perf ----
    float ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->perf.ms_forward = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->perf.ms_observable = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->perf.ms_backward = ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); c->perf.ms_total = ms;
    const double amps = (double)c->N * (double)batch;
    c->perf.passes_per_layer = P;
    c->perf.tile_bits = lp.k;
    c->perf.fwd_pass_bytes = 32.0 * amps;
    c->perf.bwd_pass_bytes = 64.0 * amps;
    c->perf.fwd_pass_ms_avg = n_fwd_pass ? c->perf.ms_forward / n_fwd_pass : 0.0;
    c->perf.bwd_pass_ms_avg = n_bwd_pass ? c->perf.ms_backward / n_bwd_pass : 0.0;
    // B_sched: init write 16, forward 32 per pass, observable 32 (+16 read-only if no grad), backward 64 per pass
    c->perf.algorithmic_bytes = amps * (16.0 + 32.0 * n_fwd_pass + (want_grad ? 32.0 : 16.0) + 64.0 * n_bwd_pass +
                                        (want_grad && c->opt_final_ladder && batch == 1 && L > 0 && has_ladder(0) ? 32.0 : 0.0));
    return 0;
}

static bool use_fused(qr_ctx* c) { return c->opt_fusion && c->n >= 4; }

extern "C" int qr_mcclean_expec(qr_ctx* c, int L, const int32_t* axes, const double* angles, const qr_obs* o,
                                int use_current_state, double* e_out) {
    QR_TRY(check_mcclean_args(c, L, axes, angles, o));
    if (!e_out) return fail(QR_EINVAL, "null output");
    QR_TRY(use_device(c));
    perf_reset(c);
    if (use_fused(c)) return mcclean_fused(c, 1, L, axes, angles, o, use_current_state, e_out, nullptr);
    return mcclean_unfused(c, L, axes, angles, o, use_current_state, e_out, nullptr);
}

extern "C" int qr_mcclean_grad(qr_ctx* c, int L, const int32_t* axes, const double* angles, const qr_obs* o,
                               int use_current_state, double* e_out, double* grad_out) {
    QR_TRY(check_mcclean_args(c, L, axes, angles, o));
    if (!e_out || (!grad_out && L > 0)) return fail(QR_EINVAL, "null output");
    QR_TRY(use_device(c));
    perf_reset(c);
    double dummy;
    if (use_fused(c)) return mcclean_fused(c, 1, L, axes, angles, o, use_current_state, e_out, grad_out ? grad_out : &dummy);
    return mcclean_unfused(c, L, axes, angles, o, use_current_state, e_out, grad_out ? grad_out : &dummy);
}

// Layered circuit = McClean's layer structure with the ladder optional per layer and without the Ry(pi/4) layer:
// the engine behind MeynardClassifier (tutorials/meynard-classifier.ipynb cells 3, 11, 14), whose layers are
// [ladder] Rx Ry Rz on every qubit, i.e. three ladder-free / ladder-first "sub-layers" of one rotation per qubit.
extern "C" int qr_layered_grad(qr_ctx* c, int L, const int32_t* axes, const double* angles, const unsigned char* ladder_before,
                               int use_current_state, const qr_obs* o, double* e_out, double* grad_out) {
    QR_TRY(check_mcclean_args(c, L, axes, angles, o));
    if (!e_out) return fail(QR_EINVAL, "null output");
    if (L > 0 && !ladder_before) return fail(QR_EINVAL, "null ladder flags");
    QR_TRY(use_device(c));
    perf_reset(c);
    const int mode = use_current_state ? 3 : 2;
    if (use_fused(c)) return mcclean_fused(c, 1, L, axes, angles, o, mode, e_out, grad_out, nullptr, ladder_before);
    return mcclean_unfused(c, L, axes, angles, o, mode, e_out, grad_out, ladder_before);
}

extern "C" int qr_mcclean_grad_batch(qr_ctx* c, int batch, int L, const int32_t* axes, const double* angles,
                                     const qr_obs* o, double* e_out, double* grad_out) {
    QR_TRY(check_obs(c, o));
    if (batch < 1) return fail(QR_EINVAL, "batch must be >= 1");
    if (!axes || !angles || !e_out || !grad_out) return fail(QR_EINVAL, "null argument");
    if (c->n < 4) return fail(QR_EINVAL, "batched path needs at least 4 qubits");
    for (i64 i = 0; i < (i64)batch * L * c->n; ++i)
        if (axes[i] < 0 || axes[i] > 2) return fail(QR_EINVAL, "Invalid axis %d", axes[i]);
    QR_TRY(use_device(c));
    perf_reset(c);
    // chunk the batch so the four ping-pong buffers of a chunk stay L2-friendly / within memory
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
    const u64 per_state = c->N * sizeof(double2) * QR_NBUF;
    const u64 chunk_bytes = c->opt_batch_chunk_mb > 0 ? (u64)c->opt_batch_chunk_mb << 20 : (u64)1 << 29;   // per buffer (measured: 512 MiB 81.5 ms, 256 MiB 84.4 ms, 128 MiB 88.9 ms per 8192 14x14 gradients)
    u64 chunk = std::max<u64>(1, std::min<u64>((u64)batch, chunk_bytes / std::max<u64>(per_state / QR_NBUF, 1)));
    // (re)allocate the buffers for `chunk` states
    if (c->buf_amps < c->N * chunk) {
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < QR_NBUF; ++i) if (c->buf_base[i]) { cudaFree(c->buf_base[i]); c->buf[i] = nullptr; c->buf_base[i] = nullptr; }
        c->buf_amps = c->N * chunk;
        c->psi = 0;
        QR_TRY(ensure_buf(c, 0));
    }
    qr_perf total;
    memset(&total, 0, sizeof(total));
    for (i64 b0 = 0; b0 < batch; b0 += (i64)chunk) {
        const i64 nb = std::min<i64>((i64)chunk, batch - b0);
        QR_TRY(mcclean_fused(c, nb, L, axes + (size_t)b0 * L * c->n, angles + (size_t)b0 * L * c->n, o, 0, e_out + b0,
                             grad_out + (size_t)b0 * L * c->n));
        total.ms_total += c->perf.ms_total; total.ms_forward += c->perf.ms_forward;
        total.ms_observable += c->perf.ms_observable; total.ms_backward += c->perf.ms_backward;
        total.algorithmic_bytes += c->perf.algorithmic_bytes; total.kernel_launches += c->perf.kernel_launches;
        total.passes_per_layer = c->perf.passes_per_layer; total.tile_bits = c->perf.tile_bits;
        total.bwd_pass_ms_avg = c->perf.bwd_pass_ms_avg; total.bwd_pass_bytes = c->perf.bwd_pass_bytes;
        total.fwd_pass_ms_avg = c->perf.fwd_pass_ms_avg; total.fwd_pass_bytes = c->perf.fwd_pass_bytes;
    }
    c->perf = total;
    // leave a valid single state in psi (state 0 of the last chunk)
    return 0;
}

// Optimiser loop on the device (optimization.py:41-91 McCleanOpt.step repeated `steps` times):
// gradient -> parameter update -> gate tables -> next gradient, all stream ordered, one synchronisation at the end.
// Device optimiser loops: `steps` identical launch sequences.  The first step runs eagerly (it settles every allocation and
// kernel attribute), the second is captured into a CUDA graph and steps 2..N are replays of it -- one launch per step instead
// of ~4 per layer, which is what a small register's step costs on the host.  Any failure while capturing falls back to eager
// launches.  body() must enqueue on c->stream only, read host data only from pinned staging it wrote itself, and not
// depend on the step index (the history index lives on the device, OptDev::step).
template <class F>
static int run_steps_graphed(qr_ctx* c, int steps, F body) {
#ifndef QR_HOST_EMUL
    bool use_graph = c->opt_loop_graph != 0 && steps > 2;
    cudaGraphExec_t exec = nullptr;
    int rc = 0;
    for (int it = 0; it < steps && rc == 0; ++it) {
        if (it == 0 || !use_graph) { rc = body(it); continue; }
        if (!exec) {
            c->capturing = true;
            cudaGraph_t graph = nullptr;
            const cudaError_t e0 = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
            const int rc2 = e0 == cudaSuccess ? body(it) : -1;
            const cudaError_t e1 = e0 == cudaSuccess ? cudaStreamEndCapture(c->stream, &graph) : e0;
            c->capturing = false;
            if (rc2 == 0 && e1 == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                cudaGraphDestroy(graph);
            } else {   // not capturable here: eager launches for the rest
                if (graph) cudaGraphDestroy(graph);
                exec = nullptr;
                use_graph = false;
                cudaGetLastError();
                g_err.clear();
                rc = body(it);
                continue;
            }
        }
        if (cudaGraphLaunch(exec, c->stream) != cudaSuccess) rc = fail(QR_ECUDA, "optimiser loop: graph launch failed");
    }
    if (exec) { cudaStreamSynchronize(c->stream); cudaGraphExecDestroy(exec); }
    return rc;
#else
    int rc = 0;
    for (int it = 0; it < steps && rc == 0; ++it) rc = body(it);
    return rc;
#endif
}

extern "C" int qr_mcclean_optimize(qr_ctx* c, int L, const int32_t* axes, double* angles, const qr_obs* o, int rule,
                                   double* hyper, int* iter_inout, double* m_inout, double* v_inout, int steps,
                                   double* cost_history, double* param_history) {
    QR_TRY(check_mcclean_args(c, L, axes, angles, o));
    if (!use_fused(c)) return fail(QR_EINVAL, "the device optimiser loop needs the fused path (>= 4 qubits, QR_OPT_FUSION)");
    if (rule < 0 || rule > 2 || !hyper || !iter_inout || steps < 0 || (steps > 0 && !cost_history)) return fail(QR_EINVAL, "bad optimiser arguments");
    if (rule == 0 && (!m_inout || !v_inout)) return fail(QR_EINVAL, "Adam needs the moment arrays");
    if (steps == 0 || L == 0) return 0;
    QR_TRY(use_device(c));
    perf_reset(c);
    const size_t np_ = (size_t)L * c->n;
    // device block: params | m | v | grad | cost history | parameter history | slot map | state
    const size_t doubles = 4 * np_ + (size_t)steps + (param_history ? (size_t)steps * np_ : 0);
    const size_t map_ints = (size_t)L * 16 * QR_GATE_SLOTS;
    char* d_blk = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_blk, doubles * sizeof(double) + map_ints * sizeof(int) + sizeof(OptDev) + 64));
    double* d_params = (double*)d_blk;
    double *d_m = d_params + np_, *d_v = d_m + np_, *d_grad = d_v + np_, *d_cost = d_grad + np_;
    double* d_hist = param_history ? d_cost + steps : nullptr;
    int* d_map = (int*)(d_params + doubles);
    OptDev* d_st = (OptDev*)(d_map + map_ints);
    OptDev st;
    memset(&st, 0, sizeof(st));
    st.rule = rule; st.iter = *iter_inout;
    st.step_size = hyper[0]; st.beta1 = hyper[1]; st.beta2 = hyper[2]; st.eps = hyper[3];
    st.plateau_length = (int)hyper[4]; st.decay_rate = hyper[5]; st.cost = hyper[6]; st.plateau_counter = (int)hyper[7];
    int rc = 0;
    std::vector<int> slot_q((size_t)L * 16 * QR_GATE_SLOTS, -1);
    int P = 0;
    do {
        cudaError_t e;
        if ((e = cudaMemcpyAsync(d_params, angles, np_ * sizeof(double), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess ||
            (e = cudaMemsetAsync(d_m, 0, 3 * np_ * sizeof(double), c->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_st, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) {
            rc = fail(QR_ECUDA, "optimiser setup: %s", cudaGetErrorString(e));
            break;
        }
        if (rule == 0) {
            cudaMemcpyAsync(d_m, m_inout, np_ * sizeof(double), cudaMemcpyHostToDevice, c->stream);
            cudaMemcpyAsync(d_v, v_inout, np_ * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        }
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { rc = fail(QR_ECUDA, "optimiser setup: %s", cudaGetErrorString(e)); break; }
        double dummy_grad = 0.0, dummy_e = 0.0;
        long long launches = 0;
        rc = run_steps_graphed(c, steps, [&](int it) -> int {
            const long long before = c->perf.kernel_launches;
            DevParams dev = {d_params, slot_q.data(), &P};
            QR_TRY(mcclean_fused(c, 1, L, axes, nullptr, o, 0, &dummy_e, &dummy_grad, &dev));
            if (it == 0) {   // the slot map is the same for every step
                if (P > 16) return fail(QR_EINVAL, "internal: too many passes");
                CUDA_TRY(cudaMemcpyAsync(d_map, slot_q.data(), (size_t)L * P * QR_GATE_SLOTS * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                CUDA_TRY(cudaStreamSynchronize(c->stream));
            }
            QR_LAUNCH(k_opt_step, 1, QR_BLOCK, 0, c->stream, (const double*)c->d_result, (const int*)d_map, L, c->n, P, QR_GATE_SLOTS, QR_SLOTS,
                      d_params, d_m, d_v, d_grad, d_st, d_cost, d_hist);
            KERNEL_CHECK();
            c->perf.kernel_launches++;
            launches = c->perf.kernel_launches - before;
            return 0;
        });
        c->perf.kernel_launches = launches * steps;   // replayed steps launch the same kernels from the graph
        if (rc) break;
        cudaMemcpyAsync(angles, d_params, np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(cost_history, d_cost, (size_t)steps * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (param_history) cudaMemcpyAsync(param_history, d_hist, (size_t)steps * np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (rule == 0) {
            cudaMemcpyAsync(m_inout, d_m, np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
            cudaMemcpyAsync(v_inout, d_v, np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        }
        cudaMemcpyAsync(&st, d_st, sizeof(st), cudaMemcpyDeviceToHost, c->stream);
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { rc = fail(QR_ECUDA, "optimiser loop: %s", cudaGetErrorString(e)); break; }
        *iter_inout = st.iter;
        hyper[0] = st.step_size; hyper[6] = st.cost; hyper[7] = (double)st.plateau_counter;
    } while (0);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_blk);
    return rc;
}

// ------------------------------------------------------------------------------------------
// QAOA (circuit_logic/qaoa.py)
// ------------------------------------------------------------------------------------------
static int qaoa_unfused(qr_ctx* c, int p, const double* betas, const double* gammas, int use_current, double* e_out,
                        double* grad) {
    const int n = c->n;
    if (!use_current) {
        QR_LAUNCH(k_init_basis, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N, 1, std::pow(2.0, -0.5 * n));
        KERNEL_CHECK();
    }
    const int g = grid_for(c, c->N);
    for (int i = 0; i < p; ++i) {
        QR_LAUNCH(k_exp_ham, g, QR_BLOCK, 0, c->stream, c->buf[c->psi], (const double*)c->d_ham, c->N, gammas[i]);
        KERNEL_CHECK();
        for (int q = 0; q < n; ++q) QR_TRY(launch_1q(c, c->buf[c->psi], n - 1 - q, rot_matrix(0, betas[i])));
    }
    QR_TRY(ensure_scratch(c, g));
    int lam = -1;
    if (grad) { lam = other_buf(c, c->psi); QR_TRY(ensure_buf(c, lam)); }
    QR_LAUNCH(k_ham_costate, g, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], grad ? c->buf[lam] : (double2*)nullptr,
              (const double*)c->d_ham, c->N, c->d_scratch);
    KERNEL_CHECK();
    QR_TRY(reduce_to_host(c, g, 1, e_out));
    if (!grad) return 0;
    const u64 npairs = c->N >> 1;
    const int gp = grid_for(c, npairs);
    for (int i = p - 1; i >= 0; --i) {
        double gb = 0.0;
        for (int q = 0; q < n; ++q) {
            QR_LAUNCH(k_pauli_inner, gp, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], (const double2*)c->buf[lam],
                      npairs, n - 1 - q, 0, c->d_scratch);
            KERNEL_CHECK();
            double v;
            QR_TRY(reduce_to_host(c, gp, 1, &v));
            gb += v;
        }
        grad[2 * i] = gb;
        for (int q = 0; q < n; ++q) {
            const Mat2 m = rot_matrix(0, -betas[i]);
            QR_TRY(launch_1q(c, c->buf[c->psi], n - 1 - q, m));
            QR_TRY(launch_1q(c, c->buf[lam], n - 1 - q, m));
        }
        QR_LAUNCH(k_ham_inner, g, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], (const double2*)c->buf[lam],
                  (const double*)c->d_ham, c->N, c->d_scratch);
        KERNEL_CHECK();
        double v;
        QR_TRY(reduce_to_host(c, g, 1, &v));
        grad[2 * i + 1] = 2.0 * v;
        QR_LAUNCH(k_exp_ham, g, QR_BLOCK, 0, c->stream, c->buf[c->psi], (const double*)c->d_ham, c->N, -gammas[i]);
        KERNEL_CHECK();
        QR_LAUNCH(k_exp_ham, g, QR_BLOCK, 0, c->stream, c->buf[lam], (const double*)c->d_ham, c->N, -gammas[i]);
        KERNEL_CHECK();
    }
    c->psi = lam;   // qaoa.py leaves the back-propagated co-state in state.vec
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// Device-resident parameters (optimiser loop, qr_qaoa_optimize): rows (beta_i, gamma_i) in device memory; the gate tables
// and the phase look-up tables are built on the device (integer-valued Hamiltonians only: a general H takes its angle as a
// kernel argument), nothing is read back -- E and the slot sums stay in d_result for the update kernel.
struct QaoaDev {
    const double* d_params;      // [p][2] on the device
    int* slot_qubit_out;         // host [P * QR_GATE_SLOTS]: qubit of each gradient slot (or -1)
    int* passes_out;             // host: P
};

static int qaoa_fused(qr_ctx* c, int p, const double* betas, const double* gammas, int use_current, double* e_out,
                      double* grad, const QaoaDev* dev = nullptr) {
    const int n = c->n;
    c->tables_fresh = true;
    LayerPlan lpf, lp;
    QR_TRY(make_plan(n, pick_tile_bits(c, n), &lpf, (int)c->opt_tile_bits_x, pick_min_row_bits(c, n)));
    lp = lpf;
    const int P = lp.npasses;
    const int GS = QR_GATE_SLOTS;
    const bool want_grad = grad != nullptr;
    const int nlay_tab = p * (want_grad ? 2 : 1);
    const size_t tab_bytes = (size_t)nlay_tab * P * GS * sizeof(GateP);
    const size_t lut_space = (size_t)2 * p * QR_LUT_MAX * sizeof(double2) + 2048;   // reserved up front: no realloc mid-stream
    QR_TRY(ensure_small(c, tab_bytes + 1024 + lut_space + (dev ? (size_t)P * GS * sizeof(int) + 2048 : 0)));
    QR_TRY(ensure_pin(c, std::max(tab_bytes + 1024 + lut_space, (size_t)(1 + (size_t)p * P * QR_SLOTS) * sizeof(double) + 1024)));
    if (dev && !(c->ham_integer && c->opt_ham_lut)) return fail(QR_EINVAL, "the device optimiser loop needs an integer-valued Hamiltonian (phase look-up tables)");
    if (dev) {
        // tables from the device-resident parameters: slot map up, one kernel for gate tables + look-up tables
        const size_t lut_off = (tab_bytes + 1024 + 255) & ~(size_t)255;
        const size_t map_off = (lut_off + lut_space + 255) & ~(size_t)255;
        int* qmap = (int*)c->h_pin;
        for (int q = 0; q < P; ++q)
            for (int s2 = 0; s2 < GS; ++s2) {
                const int gb = lp.pass[q].gbit[s2];
                qmap[q * GS + s2] = gb < 0 ? -1 : n - 1 - gb;
                dev->slot_qubit_out[q * GS + s2] = qmap[q * GS + s2];
            }
        *dev->passes_out = P;
        CUDA_TRY(cudaMemcpyAsync((char*)c->d_small + map_off, qmap, (size_t)P * GS * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        const i64 total = (i64)nlay_tab * P * GS + (i64)2 * p * c->ham_range;
        QR_LAUNCH(k_qaoa_tables, grid_for(c, (u64)total), QR_BLOCK, 0, c->stream, dev->d_params, (const int*)((char*)c->d_small + map_off), p, P, GS,
                  want_grad ? 2 : 1, c->ham_min, c->ham_range, (GatePOut*)c->d_small, (double2*)((char*)c->d_small + lut_off));
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    } else {
        GateP* tab = (GateP*)c->h_pin;
        int lay = 0;
        for (int dir = 0; dir < (want_grad ? 2 : 1); ++dir)
            for (int i = 0; i < p; ++i, ++lay) {
                const double cs = std::cos(0.5 * betas[i]), sn = (dir == 0 ? 1.0 : -1.0) * std::sin(0.5 * betas[i]);
                for (int q = 0; q < P; ++q)
                    fill_gates(dir == 0 ? lpf : lp, q, tab + ((size_t)lay * P + q) * GS, [&](int) { GateP g; g.c = cs; g.s = sn; g.axis = 0; g.pad = 0; return g; });
            }
        CUDA_TRY(cudaMemcpyAsync(c->d_small, tab, tab_bytes, cudaMemcpyHostToDevice, c->stream));
    }
    const GateP* d_tab = (const GateP*)c->d_small;
    const i64 stride = (i64)c->N;
    // phase look-up tables exp(-i gamma_i h) (forward) and exp(+i gamma_i h) (backward), host libm values
    const bool lut_on = c->ham_integer && c->opt_ham_lut;
    const double2* d_lut = nullptr;
    if (dev) d_lut = (const double2*)((char*)c->d_small + ((tab_bytes + 1024 + 255) & ~(size_t)255));
    else if (lut_on) {
        const size_t lut_off = (tab_bytes + 1024 + 255) & ~(size_t)255;
        const size_t lut_bytes = (size_t)2 * p * c->ham_range * sizeof(double2);
        double2* lut = (double2*)(c->h_pin + lut_off);
        for (int dir = 0; dir < 2; ++dir)
            for (int i = 0; i < p; ++i)
                for (int v = 0; v < c->ham_range; ++v) {
                    const double ang = (dir == 0 ? gammas[i] : -gammas[i]) * (c->ham_min + v);
                    lut[((size_t)dir * p + i) * c->ham_range + v] = make_double2(std::cos(ang), -std::sin(ang));
                }
        CUDA_TRY(cudaMemcpyAsync((char*)c->d_small + lut_off, lut, lut_bytes, cudaMemcpyHostToDevice, c->stream));
        d_lut = (const double2*)((char*)c->d_small + lut_off);
    }
    CUDA_TRY(cudaEventRecord(c->ev[0], c->stream));
    if (!use_current) {
        QR_LAUNCH(k_init_basis, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N, 1, std::pow(2.0, -0.5 * n));
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    }
    for (int i = 0; i < p; ++i)
        for (int q = 0; q < P; ++q) {
            PassIO io = {c->buf[c->psi], nullptr, c->buf[c->psi], nullptr};
            QR_TRY(launch_pass(c, lpf, q, 1, io, d_tab + ((size_t)i * P + q) * GS, 0, -1, 1, stride, 0, c->d_ham, q == 0 ? 1 : 0,
                               dev ? 0.0 : gammas[i], 0, 0.0, nullptr, nullptr, nullptr, d_lut ? d_lut + (size_t)i * c->ham_range : nullptr));
        }
    CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    const int g = grid_for(c, c->N);
    const bool defer = want_grad && c->opt_defer_reduce && p > 0 && uniform_lean_plan(lp);   // see mcclean_fused
    const size_t unit_cap = (size_t)c->sm_count * 16;
    QR_TRY(ensure_scratch(c, std::max<size_t>((size_t)g, defer ? (size_t)p * P * unit_cap * QR_SLOTS : (size_t)0)));
    QR_TRY(ensure_result(c, 1 + (size_t)p * P * QR_SLOTS + 16));
    int lam = -1;
    if (want_grad) { lam = other_buf(c, c->psi); QR_TRY(ensure_buf(c, lam)); }
    QR_LAUNCH(k_ham_costate, g, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], want_grad ? c->buf[lam] : (double2*)nullptr,
              (const double*)c->d_ham, c->N, c->d_scratch);
    KERNEL_CHECK();
    QR_LAUNCH(k_reduce_partials, 1, QR_BLOCK, 0, c->stream, (const double*)c->d_scratch, g, 1, c->d_result);
    KERNEL_CHECK();
    c->perf.kernel_launches += 2;
    CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    double* d_slots = c->d_result + 1;
    int n_bwd_pass = 0;
    int defer_units = -1;
    if (want_grad) {
        for (int i = p - 1; i >= 0; --i)
            for (int q = 0; q < P; ++q) {
                PassIO io = {c->buf[c->psi], c->buf[lam], c->buf[c->psi], c->buf[lam]};
                int units = 0;
                QR_TRY(launch_pass(c, lp, q, 2, io, d_tab + ((size_t)(p + i) * P + q) * GS, 0, -1, 1, stride, 0, c->d_ham, 0, 0.0,
                                   q == P - 1 ? 1 : 0, dev ? 0.0 : -gammas[i], &units, nullptr, defer ? nullptr : d_slots + ((size_t)i * P + q) * QR_SLOTS,
                                   d_lut ? d_lut + (size_t)(p + i) * c->ham_range : nullptr,
                                   defer ? c->d_scratch + ((size_t)i * P + q) * unit_cap * QR_SLOTS : nullptr));
                if (defer) {
                    if (defer_units >= 0 && units != defer_units) return fail(QR_ESTATE, "internal: passes of one gradient use different grids");
                    defer_units = units;
                }
                ++n_bwd_pass;
            }
        if (defer) {
            QR_LAUNCH(k_reduce_slots_strided, (unsigned)(p * P), 32 * 16, 0, c->stream, (const double*)c->d_scratch, defer_units,
                      (u64)unit_cap * QR_SLOTS, d_slots, QR_SLOTS);
            KERNEL_CHECK();
            c->perf.kernel_launches++;
        }
    }
    CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    if (dev) {   // results stay on the device
        if (want_grad) c->psi = lam;
        return 0;
    }
    const size_t nres = 1 + (want_grad ? (size_t)p * P * QR_SLOTS : 0);
    QR_TRY(ensure_pin(c, nres * sizeof(double)));
    CUDA_TRY(cudaMemcpyAsync(c->h_pin, c->d_result, nres * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const double* res = (const double*)c->h_pin;
    *e_out = res[0];
    if (want_grad) {
        for (int i = 0; i < p; ++i) {
            double gb = 0.0;
            for (int q = 0; q < P; ++q)
                for (int s = 0; s < GS; ++s)
                    if (lp.pass[q].gbit[s] >= 0) gb += res[1 + ((size_t)i * P + q) * QR_SLOTS + s];
            grad[2 * i] = gb;                                                              // qaoa.py:62-63
            grad[2 * i + 1] = 2.0 * res[1 + ((size_t)i * P + (P - 1)) * QR_SLOTS + (QR_SLOTS - 1)];   // qaoa.py:67-68
        }
        c->psi = lam;
    }
    float ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]); c->perf.ms_forward = ms;
    cudaEventElapsedTime(&ms, c->ev[1], c->ev[2]); c->perf.ms_observable = ms;
    cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]); c->perf.ms_backward = ms;
    cudaEventElapsedTime(&ms, c->ev[0], c->ev[3]); c->perf.ms_total = ms;
    const double amps = (double)c->N;
    c->perf.passes_per_layer = P;
    c->perf.tile_bits = lp.k;
    c->perf.fwd_pass_bytes = 32.0 * amps;
    c->perf.bwd_pass_bytes = 64.0 * amps;
    c->perf.fwd_pass_ms_avg = p ? c->perf.ms_forward / (p * P) : 0.0;
    c->perf.bwd_pass_ms_avg = n_bwd_pass ? c->perf.ms_backward / n_bwd_pass : 0.0;
    // H table reads: 8 B/amp in the first forward pass and the last backward pass of every layer
    c->perf.algorithmic_bytes = amps * (16.0 + (32.0 * P + 8.0) * p + (want_grad ? 40.0 : 24.0) + (want_grad ? (64.0 * P + 8.0) * p : 0.0));
    return 0;
}

static int check_qaoa_args(qr_ctx* c, int p, const double* betas, const double* gammas) {
    QR_TRY(need_ham(c));
    if (p < 0) return fail(QR_EINVAL, "layer_number must be >= 0");
    if (p > 0 && (!betas || !gammas)) return fail(QR_EINVAL, "null parameters");
    return 0;
}

extern "C" int qr_qaoa_expec(qr_ctx* c, int p, const double* betas, const double* gammas, int use_current_state, double* e_out) {
    QR_TRY(check_qaoa_args(c, p, betas, gammas));
    if (!e_out) return fail(QR_EINVAL, "null output");
    QR_TRY(use_device(c));
    perf_reset(c);
    if (use_fused(c)) return qaoa_fused(c, p, betas, gammas, use_current_state, e_out, nullptr);
    return qaoa_unfused(c, p, betas, gammas, use_current_state, e_out, nullptr);
}

extern "C" int qr_qaoa_grad(qr_ctx* c, int p, const double* betas, const double* gammas, int use_current_state, double* e_out,
                            double* grad_out) {
    QR_TRY(check_qaoa_args(c, p, betas, gammas));
    if (!e_out || (!grad_out && p > 0)) return fail(QR_EINVAL, "null output");
    QR_TRY(use_device(c));
    perf_reset(c);
    double dummy[2];
    if (use_fused(c)) return qaoa_fused(c, p, betas, gammas, use_current_state, e_out, grad_out ? grad_out : dummy);
    return qaoa_unfused(c, p, betas, gammas, use_current_state, e_out, grad_out ? grad_out : dummy);
}

// QaoaOpt.step x steps on the device (optimization.py:113-129): params = rows (beta_i, gamma_i), same rules and state
// hand-over as qr_mcclean_optimize
extern "C" int qr_qaoa_optimize(qr_ctx* c, int p, double* params, int rule, double* hyper, int* iter_inout, double* m_inout,
                                double* v_inout, int steps, double* cost_history, double* param_history) {
    QR_TRY(check_qaoa_args(c, p, params, params));
    if (!use_fused(c)) return fail(QR_EINVAL, "the device optimiser loop needs the fused path (>= 4 qubits, QR_OPT_FUSION)");
    if (rule < 0 || rule > 2 || !hyper || !iter_inout || steps < 0 || (steps > 0 && !cost_history)) return fail(QR_EINVAL, "bad optimiser arguments");
    if (rule == 0 && (!m_inout || !v_inout)) return fail(QR_EINVAL, "Adam needs the moment arrays");
    if (!(c->ham_integer && c->opt_ham_lut)) return fail(QR_EINVAL, "the device optimiser loop needs an integer-valued Hamiltonian (phase look-up tables)");
    if (steps == 0 || p == 0) return 0;
    QR_TRY(use_device(c));
    perf_reset(c);
    const size_t np_ = (size_t)2 * p;
    const size_t doubles = 4 * np_ + (size_t)steps + (param_history ? (size_t)steps * np_ : 0);
    const size_t map_ints = 16 * QR_GATE_SLOTS;
    char* d_blk = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d_blk, doubles * sizeof(double) + map_ints * sizeof(int) + sizeof(OptDev) + 64));
    double* d_params = (double*)d_blk;
    double *d_m = d_params + np_, *d_v = d_m + np_, *d_grad = d_v + np_, *d_cost = d_grad + np_;
    double* d_hist = param_history ? d_cost + steps : nullptr;
    int* d_map = (int*)(d_params + doubles);
    OptDev* d_st = (OptDev*)(d_map + map_ints);
    OptDev st;
    memset(&st, 0, sizeof(st));
    st.rule = rule; st.iter = *iter_inout;
    st.step_size = hyper[0]; st.beta1 = hyper[1]; st.beta2 = hyper[2]; st.eps = hyper[3];
    st.plateau_length = (int)hyper[4]; st.decay_rate = hyper[5]; st.cost = hyper[6]; st.plateau_counter = (int)hyper[7];
    int rc = 0;
    std::vector<int> slot_q(16 * QR_GATE_SLOTS, -1);
    int P = 0;
    do {
        cudaError_t e;
        if ((e = cudaMemcpyAsync(d_params, params, np_ * sizeof(double), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess ||
            (e = cudaMemsetAsync(d_m, 0, 3 * np_ * sizeof(double), c->stream)) != cudaSuccess ||
            (e = cudaMemcpyAsync(d_st, &st, sizeof(st), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) {
            rc = fail(QR_ECUDA, "optimiser setup: %s", cudaGetErrorString(e));
            break;
        }
        if (rule == 0) {
            cudaMemcpyAsync(d_m, m_inout, np_ * sizeof(double), cudaMemcpyHostToDevice, c->stream);
            cudaMemcpyAsync(d_v, v_inout, np_ * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        }
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { rc = fail(QR_ECUDA, "optimiser setup: %s", cudaGetErrorString(e)); break; }
        double dummy_grad[2] = {0.0, 0.0}, dummy_e = 0.0;
        long long launches = 0;
        rc = run_steps_graphed(c, steps, [&](int it) -> int {
            const long long before = c->perf.kernel_launches;
            QaoaDev dev = {d_params, slot_q.data(), &P};
            QR_TRY(qaoa_fused(c, p, nullptr, nullptr, 0, &dummy_e, dummy_grad, &dev));
            if (it == 0) {   // the slot map is the same for every step
                if (P > 16) return fail(QR_EINVAL, "internal: too many passes");
                CUDA_TRY(cudaMemcpyAsync(d_map, slot_q.data(), (size_t)P * QR_GATE_SLOTS * sizeof(int), cudaMemcpyHostToDevice, c->stream));
                CUDA_TRY(cudaStreamSynchronize(c->stream));
            }
            QR_LAUNCH(k_qaoa_opt_step, 1, QR_BLOCK, 0, c->stream, (const double*)c->d_result, (const int*)d_map, p, P, QR_GATE_SLOTS, QR_SLOTS,
                      d_params, d_m, d_v, d_grad, d_st, d_cost, d_hist);
            KERNEL_CHECK();
            c->perf.kernel_launches++;
            launches = c->perf.kernel_launches - before;
            return 0;
        });
        c->perf.kernel_launches = launches * steps;
        if (rc) break;
        cudaMemcpyAsync(params, d_params, np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(cost_history, d_cost, (size_t)steps * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (param_history) cudaMemcpyAsync(param_history, d_hist, (size_t)steps * np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (rule == 0) {
            cudaMemcpyAsync(m_inout, d_m, np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
            cudaMemcpyAsync(v_inout, d_v, np_ * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        }
        cudaMemcpyAsync(&st, d_st, sizeof(st), cudaMemcpyDeviceToHost, c->stream);
        if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) { rc = fail(QR_ECUDA, "optimiser loop: %s", cudaGetErrorString(e)); break; }
        *iter_inout = st.iter;
        hyper[0] = st.step_size; hyper[6] = st.cost; hyper[7] = (double)st.plateau_counter;
    } while (0);
    cudaStreamSynchronize(c->stream);
    cudaFree(d_blk);
    return rc;
}

// ------------------------------------------------------------------------------------------
// sampling
// ------------------------------------------------------------------------------------------
extern "C" int qr_sample_bitstrings(qr_ctx* c, int S, const double* uniforms, int64_t* out_idx) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (S < 0 || (S > 0 && (!uniforms || !out_idx))) return fail(QR_EINVAL, "bad sample arguments");
    if (S == 0) return 0;
    QR_TRY(use_device(c));
    const u64 nchunks = (c->N + QR_SCAN_CHUNK - 1) / QR_SCAN_CHUNK;
    QR_TRY(ensure_scratch(c, nchunks + 2 * (size_t)S + 16));
    double* d_cdf = c->d_scratch;
    double* d_u = c->d_scratch + nchunks;
    i64* d_idx = (i64*)(c->d_scratch + nchunks + S);
    QR_TRY(ensure_pin(c, (size_t)S * 16));
    memcpy(c->h_pin, uniforms, (size_t)S * sizeof(double));
    CUDA_TRY(cudaMemcpyAsync(d_u, c->h_pin, (size_t)S * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int grid = (int)std::min<u64>(nchunks, (u64)c->sm_count * 8);
    QR_LAUNCH(k_prob_chunk_sums, grid, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], c->N, d_cdf);
    KERNEL_CHECK();
    QR_LAUNCH(k_scan_inclusive_single, 1, 1024, 0, c->stream, d_cdf, nchunks);
    KERNEL_CHECK();
    QR_LAUNCH(k_sample_search, S, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], c->N, (const double*)d_cdf, nchunks,
              (const double*)d_u, d_idx);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemcpyAsync(c->h_pin, d_idx, (size_t)S * sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out_idx, c->h_pin, (size_t)S * sizeof(i64));
    return 0;
}

// ---- state permutation (sampling in the order of the sorted eigenvalues of a diagonal observable) ----
extern "C" int qr_perm_load(qr_ctx* c, const int64_t* perm, size_t n_amps) {
    if (!c || !perm) return fail(QR_EINVAL, "null argument");
    if (n_amps != c->N) return fail(QR_EINVAL, "permutation must have 2^n = %llu entries", (unsigned long long)c->N);
    std::vector<bool> seen(c->N, false);
    for (u64 k = 0; k < c->N; ++k) {
        if (perm[k] < 0 || (u64)perm[k] >= c->N || seen[(size_t)perm[k]]) return fail(QR_EINVAL, "not a permutation of 0..2^n-1 (entry %llu)", (unsigned long long)k);
        seen[(size_t)perm[k]] = true;
    }
    QR_TRY(use_device(c));
    if (!c->d_perm) {
        cudaError_t e = cudaMalloc((void**)&c->d_perm, c->N * sizeof(i64));
        if (e != cudaSuccess) { c->d_perm = nullptr; return fail(QR_ENOMEM, "cannot allocate the permutation table: %s", cudaGetErrorString(e)); }
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_perm, perm, c->N * sizeof(i64), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qr_state_permute(qr_ctx* c) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (!c->d_perm) return fail(QR_ESTATE, "no permutation loaded (qr_perm_load)");
    QR_TRY(use_device(c));
    const int dst = other_buf(c, c->psi);
    QR_TRY(ensure_buf(c, dst));
    QR_LAUNCH(k_permute_gather, grid_for(c, c->N), QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi], (const i64*)c->d_perm, c->buf[dst], c->N);
    KERNEL_CHECK();
    c->psi = dst;
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

// Dense basis change (observables with x / y terms are measured in the eigenbasis of the dense 2^n x 2^n observable,
// mc_clean.py:221-224, 255-256): the caller supplies M = V^dagger of numpy.linalg.eigh, row major, interleaved re/im.
#define QR_DENSE_MAX_QUBITS 12
extern "C" int qr_dense_load(qr_ctx* c, const double* m_re_im, size_t dim) {
    if (!c || !m_re_im) return fail(QR_EINVAL, "null argument");
    if (c->n > QR_DENSE_MAX_QUBITS) return fail(QR_EINVAL, "dense basis changes are limited to %d qubits (2^n x 2^n matrix)", QR_DENSE_MAX_QUBITS);
    if (dim != c->N) return fail(QR_EINVAL, "matrix must be 2^n x 2^n with 2^n = %llu, got %llu", (unsigned long long)c->N, (unsigned long long)dim);
    QR_TRY(use_device(c));
    if (!c->d_dense) {
        cudaError_t e = cudaMalloc((void**)&c->d_dense, c->N * c->N * sizeof(double2));
        if (e != cudaSuccess) { c->d_dense = nullptr; return fail(QR_ENOMEM, "cannot allocate the dense matrix: %s", cudaGetErrorString(e)); }
    }
    CUDA_TRY(cudaMemcpyAsync(c->d_dense, m_re_im, c->N * c->N * sizeof(double2), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

extern "C" int qr_state_apply_dense(qr_ctx* c) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (!c->d_dense) return fail(QR_ESTATE, "no dense matrix loaded (qr_dense_load)");
    QR_TRY(use_device(c));
    const int dst = other_buf(c, c->psi);
    QR_TRY(ensure_buf(c, dst));
    const int grid = (int)std::min<u64>((c->N * 32 + QR_BLOCK - 1) / QR_BLOCK, (u64)c->sm_count * 8);
    QR_LAUNCH(k_dense_matvec, grid, QR_BLOCK, 0, c->stream, (const double2*)c->d_dense, (const double2*)c->buf[c->psi], c->buf[dst], c->N);
    KERNEL_CHECK();
    c->psi = dst;
    return 0;   // stream ordered: the next call that returns data to the host synchronises
}

extern "C" int qr_ham_gather(qr_ctx* c, int n, const int64_t* idx, double* out) {
    QR_TRY(need_ham(c));
    if (n < 0 || (n > 0 && (!idx || !out))) return fail(QR_EINVAL, "bad gather arguments");
    if (n == 0) return 0;
    for (int i = 0; i < n; ++i)
        if (idx[i] < 0 || (u64)idx[i] >= c->N) return fail(QR_EINVAL, "index %lld out of range", (long long)idx[i]);
    QR_TRY(use_device(c));
    QR_TRY(ensure_scratch(c, 2 * (size_t)n + 16));
    QR_TRY(ensure_pin(c, (size_t)n * 16));
    i64* d_idx = (i64*)c->d_scratch;
    double* d_out = c->d_scratch + n;
    memcpy(c->h_pin, idx, (size_t)n * sizeof(i64));
    CUDA_TRY(cudaMemcpyAsync(d_idx, c->h_pin, (size_t)n * sizeof(i64), cudaMemcpyHostToDevice, c->stream));
    QR_LAUNCH(k_gather_f64, (n + QR_BLOCK - 1) / QR_BLOCK, QR_BLOCK, 0, c->stream, (const double*)c->d_ham, (const i64*)d_idx, n, d_out);
    KERNEL_CHECK();
    CUDA_TRY(cudaMemcpyAsync(c->h_pin, d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    memcpy(out, c->h_pin, (size_t)n * sizeof(double));
    return 0;
}


#include "qr_shard.cuh"

// ------------------------------------------------------------------------------------------
// Sharded state vector: one shard per rank, top log2(G) qubits = rank bits.
//   * CNOT ladder: GF(2)-linear and banded towards more significant bits, so a whole destination
//     shard reads exactly one source shard -> pure relabelling of shards (pi[rank] = logical shard
//     id) plus the local gather with a rank-dependent XOR on the top local bits.  No data moves.
//   * rotations on global qubits: k_global_gates (exchange + gates fused over peer memory).
//   * the host sequences the steps; a cross-rank barrier is required after every step.
// ------------------------------------------------------------------------------------------
struct ShardRun {
    int L = 0, P = 0;
    bool want_grad = false;
    std::vector<int32_t> axes;
    std::vector<double> angles;
    std::vector<ObsTerm> terms;
    LayerPlan lpf, lpb;
    std::vector<int> pi;          // pi[physical rank] = logical shard id
    int lam = -1;
    size_t tab_off = 0;
    int n_steps = 0;
    size_t res_local = 0, res_global = 0, res_total = 0;   // offsets (doubles) in d_result
};

static void shard_release(qr_ctx* c) {
    for (int r = 0; r < QR_MAX_RANKS; ++r)
        for (int b = 0; b < QR_NBUF; ++b)
            if (c->peer_mapped[r][b] && c->peer[r][b]) { cudaIpcCloseMemHandle(c->peer[r][b]); c->peer[r][b] = nullptr; c->peer_mapped[r][b] = false; }
    for (int r = 0; r < QR_MAX_RANKS; ++r)
        if (c->flags_mapped[r] && c->peer_flags[r]) { cudaIpcCloseMemHandle(c->peer_flags[r]); c->peer_flags[r] = nullptr; c->flags_mapped[r] = false; }
    delete c->run;
    c->run = nullptr;
    delete c->srun;
    c->srun = nullptr;
}

extern "C" int qr_shard_create(int n_total, int log2_world, int rank, int device, qr_ctx** out) {
    if (!out) return fail(QR_EINVAL, "null output");
    if (log2_world < 1 || log2_world > 4) return fail(QR_EINVAL, "log2(world size) must be in [1, 4]");
    if (rank < 0 || rank >= (1 << log2_world)) return fail(QR_EINVAL, "rank %d out of range", rank);
    const int nl = n_total - log2_world;
    if (nl < 4 || nl < log2_world) return fail(QR_EINVAL, "sharded registers need at least %d local qubits", std::max(4, log2_world));
    QR_TRY(qr_ctx_create(nl, device, out));
    qr_ctx* c = *out;
    c->n_total = n_total;
    c->g = log2_world;
    c->rank = rank;
    for (int b = 0; b < QR_NBUF; ++b) {
        int rc = ensure_buf(c, b);
        if (rc) { qr_ctx_destroy(c); *out = nullptr; return rc; }
        c->peer[rank][b] = c->buf[b];
    }
    if (cudaMalloc((void**)&c->d_flags, QR_FLAG_WORDS * sizeof(unsigned long long)) != cudaSuccess) {
        qr_ctx_destroy(c); *out = nullptr;
        return fail(QR_ENOMEM, "flag allocation failed");
    }
    cudaMemsetAsync(c->d_flags, 0, QR_FLAG_WORDS * sizeof(unsigned long long), c->stream);
    cudaStreamSynchronize(c->stream);
    c->peer_flags[rank] = c->d_flags;
    return 0;
}

static int need_shard(qr_ctx* c) {
    if (!c) return fail(QR_EINVAL, "null context");
    if (c->g == 0) return fail(QR_ESTATE, "context is not sharded (use qr_shard_create)");
    return 0;
}

extern "C" int qr_shard_ipc_handle(qr_ctx* c, int buf, void* handle64) {
    QR_TRY(need_shard(c));
    if (buf < 0 || buf > QR_NBUF || !handle64) return fail(QR_EINVAL, "bad buffer index");
    QR_TRY(use_device(c));
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, buf == QR_NBUF ? (void*)c->d_flags : c->buf_base[buf]));
    static_assert(sizeof(h) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int qr_shard_ipc_open(qr_ctx* c, int peer_rank, int buf, const void* handle64) {
    QR_TRY(need_shard(c));
    if (peer_rank < 0 || peer_rank >= (1 << c->g) || buf < 0 || buf > QR_NBUF || !handle64) return fail(QR_EINVAL, "bad peer/buffer");
    if (peer_rank == c->rank) return 0;
    QR_TRY(use_device(c));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    if (buf == QR_NBUF) {   // the peer's ordering flags
        c->peer_flags[peer_rank] = (unsigned long long*)p;
        c->flags_mapped[peer_rank] = true;
        return 0;
    }
    c->peer[peer_rank][buf] = (double2*)p;
    c->peer_mapped[peer_rank][buf] = true;
    return 0;
}

// same-process peers (one process driving several devices, or the CPU test tier)
extern "C" int qr_shard_set_peer_ptr(qr_ctx* c, int peer_rank, int buf, void* ptr, int peer_device) {
    QR_TRY(need_shard(c));
    if (peer_rank < 0 || peer_rank >= (1 << c->g) || buf < 0 || buf > QR_NBUF || !ptr) return fail(QR_EINVAL, "bad peer/buffer");
    if (peer_rank == c->rank) return 0;
    QR_TRY(use_device(c));
    if (peer_device != c->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
        if (e != cudaSuccess) cudaGetLastError();   // already enabled is fine
    }
    if (buf == QR_NBUF) { c->peer_flags[peer_rank] = (unsigned long long*)ptr; return 0; }
    c->peer[peer_rank][buf] = (double2*)ptr;
    return 0;
}

extern "C" int qr_shard_buffer_ptr(qr_ctx* c, int buf, void** out) {
    QR_TRY(need_shard(c));
    if (buf < 0 || buf > QR_NBUF || !out) return fail(QR_EINVAL, "bad buffer index");
    *out = buf == QR_NBUF ? (void*)c->d_flags : (void*)c->buf[buf];
    return 0;
}

// full-register ladder map helpers
static inline u64 full_map(u64 d, u64 m1, u64 m2) { return ladder_map(d, m1, m2); }

// relabel after a ladder gather with masks (m1, m2): logical dest r reads logical source
// sigma(r) = full_map(r << nl) >> nl; returns this rank's new logical id and the XOR carry.
static void shard_relabel(qr_ctx* c, ShardRun* run, u64 m1, u64 m2, LadderSpec* spec) {
    const int nl = c->n, G = 1 << c->g;
    std::vector<int> inv_sigma(G);
    std::vector<u64> carry(G);
    for (int r = 0; r < G; ++r) {
        const u64 img = full_map((u64)r << nl, m1, m2);
        inv_sigma[(int)(img >> nl)] = r;
        carry[r] = img & (((u64)1 << nl) - 1);
    }
    const int r_new = inv_sigma[run->pi[c->rank]];
    spec->M1 = m1 & (((u64)1 << nl) - 1);
    spec->M2 = m2 & (((u64)1 << nl) - 1);
    spec->src_xor = carry[r_new];
    for (int rho = 0; rho < G; ++rho) run->pi[rho] = inv_sigma[run->pi[rho]];
}

// plan a swap-engine run, upload its tables, reset the result block
static int shard_swap_begin(qr_ctx* c, const SwapCircuit& circ, const qr_obs* o, bool want_grad, int* n_steps) {
    const int G = 1 << c->g;
    for (int r = 0; r < G; ++r)
        if (!c->peer_flags[r]) return fail(QR_ESTATE, "the ordering flags of rank %d are not mapped", r);
    delete c->srun;
    SwapRun* sr = c->srun = new SwapRun();
    sr->lockstep = c->opt_shard_lockstep != 0;
    sr->slices = (int)c->opt_shard_slices;
    sr->gen_base = c->flag_gen;
    std::vector<GateP> tab;
    QR_TRY(swap_build(c, sr, circ, o, want_grad, &tab));
    c->flag_gen = sr->gen_base;
    // observable terms (remapped to the final layout of the forward sweep) at offset 0, gate tables behind them
    const std::vector<ObsTerm>* terms = nullptr;
    for (const XOp& op : sr->ops) if (op.kind == 2) terms = &op.terms;
    const size_t terms_bytes = ((terms ? terms->size() : 0) + 1) * sizeof(ObsTerm);
    sr->tab_off = (terms_bytes + 255) & ~(size_t)255;
    const size_t tab_bytes = tab.size() * sizeof(GateP);
    const bool lut_on = circ.kind == 1 && c->ham_integer && c->opt_ham_lut;
    const size_t lut_bytes = lut_on ? (size_t)2 * circ.L * c->ham_range * sizeof(double2) : 0;
    const size_t lut_off = (sr->tab_off + tab_bytes + 255) & ~(size_t)255;
    QR_TRY(ensure_small(c, lut_off + lut_bytes + 2048));
    QR_TRY(ensure_pin(c, std::max(lut_off + lut_bytes + 2048, (sr->n_results + 16) * sizeof(double))));
    QR_TRY(ensure_result(c, sr->n_results + 16));
    QR_TRY(ensure_scratch(c, std::max<size_t>(2 * (size_t)c->sm_count * 16 * QR_SLOTS, (size_t)grid_for(c, c->N))));
    memcpy(c->h_pin + sr->tab_off, tab.data(), tab_bytes);
    if (lut_on) {   // exp(-i gamma_i h) (forward) and exp(+i gamma_i h) (backward), host libm values (as qaoa_fused)
        double2* lut = (double2*)(c->h_pin + lut_off);
        for (int dir = 0; dir < 2; ++dir)
            for (int i = 0; i < circ.L; ++i)
                for (int v = 0; v < c->ham_range; ++v) {
                    const double ang = (dir == 0 ? circ.gammas[i] : -circ.gammas[i]) * (c->ham_min + v);
                    lut[((size_t)dir * circ.L + i) * c->ham_range + v] = make_double2(std::cos(ang), -std::sin(ang));
                }
        sr->lut_off = lut_off;
    }
    CUDA_TRY(cudaMemcpyAsync((char*)c->d_small + sr->tab_off, c->h_pin + sr->tab_off, lut_off + lut_bytes - sr->tab_off, cudaMemcpyHostToDevice, c->stream));
    if (terms && !terms->empty()) QR_TRY(upload_small(c, 0, terms->data(), terms->size() * sizeof(ObsTerm), 0));
    CUDA_TRY(cudaMemsetAsync(c->d_result, 0, (sr->n_results + 16) * sizeof(double), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    perf_reset(c);
    c->perf.passes_per_layer = sr->sweeps;
    c->perf.tile_bits = 12;
    c->tables_fresh = true;
    *n_steps = (int)sr->step_first.size();
    return 0;
}

extern "C" int qr_shard_mcclean_begin(qr_ctx* c, int L, const int32_t* axes, const double* angles, const qr_obs* o,
                                      int want_grad, int* n_steps) {
    QR_TRY(need_shard(c));
    if (!o || o->n != c->n_total) return fail(QR_EINVAL, "observable must be defined on all %d qubits", c->n_total);
    if (L < 0 || (L > 0 && (!axes || !angles)) || !n_steps) return fail(QR_EINVAL, "bad arguments");
    const int nt = c->n_total, nl = c->n, G = 1 << c->g;
    for (i64 i = 0; i < (i64)L * nt; ++i)
        if (axes[i] < 0 || axes[i] > 2) return fail(QR_EINVAL, "Invalid axis %d", axes[i]);
    for (int r = 0; r < G; ++r)
        for (int b = 0; b < QR_NBUF; ++b)
            if (!c->peer[r][b]) return fail(QR_ESTATE, "peer buffers of rank %d are not mapped", r);
    QR_TRY(use_device(c));
    delete c->run;
    c->run = nullptr;
    delete c->srun;
    c->srun = nullptr;
    for (int r = 0; r < G; ++r)
        for (int b2 = 0; b2 < QR_NBUF; ++b2)
            if (!c->peer[r][b2]) return fail(QR_ESTATE, "peer buffers of rank %d are not mapped", r);
    const bool swap_ok = swap_engine_ok(nl, c->g) && c->opt_fusion && c->opt_tile_bits == 0 && c->opt_tile_bits_x == 0 && c->opt_min_row_bits == 3;
    if (c->opt_shard_mode == 2 && !swap_ok)
        return fail(QR_EINVAL, "the swap engine needs 1..3 rank bits, at least %d local qubits and default tile options", 12 + c->g);
    if (c->opt_shard_mode == 2 || (c->opt_shard_mode == 0 && swap_ok)) {
        SwapCircuit circ = {0, L, axes, angles, nullptr, nullptr};
        return shard_swap_begin(c, circ, o, want_grad != 0, n_steps);
    }
    ShardRun* run = c->run = new ShardRun();
    run->L = L;
    run->want_grad = want_grad != 0;
    run->axes.assign(axes, axes + (size_t)L * nt);
    run->angles.assign(angles, angles + (size_t)L * nt);
    run->terms = o->terms;
    QR_TRY(make_plan(nl, pick_tile_bits(c, nl), &run->lpf, (int)c->opt_tile_bits_x, pick_min_row_bits(c, nl)));
    run->lpb = run->lpf;
    const int P = run->P = run->lpb.npasses;
    run->pi.resize(G);
    for (int r = 0; r < G; ++r) run->pi[r] = r;
    c->psi = 0;
    // gate tables of the local passes: [forward L][P][GS] then [backward L][P][GS]
    const int GS = QR_GATE_SLOTS;
    const size_t terms_bytes = (o->terms.size() + 1) * sizeof(ObsTerm);
    run->tab_off = (terms_bytes + 255) & ~(size_t)255;
    const size_t tab_entries = (size_t)(run->want_grad ? 2 : 1) * L * P * GS;
    const size_t tab_bytes = tab_entries * sizeof(GateP);
    run->res_local = 1;
    run->res_global = 1 + (size_t)L * P * QR_SLOTS;
    run->res_total = run->res_global + (size_t)L * 4;
    QR_TRY(ensure_small(c, run->tab_off + tab_bytes + 1024));
    QR_TRY(ensure_pin(c, std::max(run->tab_off + tab_bytes + 1024, run->res_total * sizeof(double))));
    QR_TRY(ensure_result(c, run->res_total + 16));
    QR_TRY(ensure_scratch(c, (size_t)c->sm_count * 16 * QR_SLOTS));
    GateP* tab = (GateP*)(c->h_pin + run->tab_off);
    int lay = 0;
    for (int dir = 0; dir < (run->want_grad ? 2 : 1); ++dir)
        for (int i = 0; i < L; ++i, ++lay) {
            const int32_t* ax = axes + (size_t)i * nt;
            const double* an = angles + (size_t)i * nt;
            const double sgn = dir == 0 ? 1.0 : -1.0;
            for (int p = 0; p < P; ++p)
                fill_gates(dir == 0 ? run->lpf : run->lpb, p, tab + ((size_t)lay * P + p) * GS, [&](int qloc) {
                    const int q = qloc + c->g;
                    GateP g; g.c = std::cos(0.5 * an[q]); g.s = sgn * std::sin(0.5 * an[q]); g.axis = ax[q]; g.pad = 0; return g; });
        }
    CUDA_TRY(cudaMemcpyAsync((char*)c->d_small + run->tab_off, tab, tab_bytes, cudaMemcpyHostToDevice, c->stream));
    if (!o->terms.empty()) QR_TRY(upload_small(c, 0, o->terms.data(), o->terms.size() * sizeof(ObsTerm), 0));
    CUDA_TRY(cudaMemsetAsync(c->d_result, 0, (run->res_total + 16) * sizeof(double), c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    perf_reset(c);
    run->n_steps = run->want_grad ? 4 * L + 1 : 2 * L + 1;
    *n_steps = run->n_steps;
    return 0;
}

static int shard_global_step(qr_ctx* c, ShardRun* run, int layer, int nv) {
    const int G = 1 << c->g, nt = c->n_total;
    GlobalGates gg;
    memset(&gg, 0, sizeof(gg));
    const double sgn = nv == 2 ? -1.0 : 1.0;
    const int me = run->pi[c->rank];          // logical shard id held by this rank
    // X / Y rotations form the exchange group; Rz gates are diagonal (a constant phase per subgroup, no exchange)
    int active[4], ga = 0;
    unsigned zmask = 0;
    gg.zphase = make_double2(1.0, 0.0);
    for (int b = 0; b < c->g; ++b) {
        const int q = c->g - 1 - b;
        const double an = run->angles[(size_t)layer * nt + q];
        const int axis = run->axes[(size_t)layer * nt + q];
        const double cs = std::cos(0.5 * an), sn = sgn * std::sin(0.5 * an);
        if (axis == 2 && c->opt_shard_zskip) {
            const int v = (me >> b) & 1;       // (c - i s) on bit value 0, (c + i s) on bit value 1 (state.py:168-170)
            const double pr = cs, pi_ = v ? sn : -sn, zr = gg.zphase.x, zi = gg.zphase.y;
            gg.zphase = make_double2(zr * pr - zi * pi_, zr * pi_ + zi * pr);
            gg.zslot[gg.nz] = b;
            gg.zsign[gg.nz] = v ? -1.0 : 1.0;
            gg.nz++;
            zmask |= 1u << b;
        } else {
            gg.gate[ga].c = cs; gg.gate[ga].s = sn; gg.gate[ga].axis = axis;
            gg.slot[ga] = b;
            active[ga++] = b;
        }
    }
    auto member = [&](int shard) { int m = 0; for (int i = 0; i < ga; ++i) m |= ((shard >> active[i]) & 1) << i; return m; };
    gg.ga = ga;
    gg.slice_len = c->N >> ga;
    gg.slice_off = (u64)member(me) * gg.slice_len;
    for (int rho = 0; rho < G; ++rho) {
        const int t = run->pi[rho];
        if ((t & zmask) != (me & zmask)) continue;   // another subgroup
        gg.psi[member(t)] = c->peer[rho][c->psi];
        if (nv == 2) gg.lam[member(t)] = c->peer[rho][run->lam];
    }
    const int grid = (int)std::min<u64>((gg.slice_len + 255) / 256, (u64)c->sm_count * 4);
    gg.partials = c->d_scratch;
    typedef void (*gfn)(const GlobalGates);
    static const gfn fns[2][5] = {
        {k_global_gates<1, 0>, k_global_gates<1, 1>, k_global_gates<1, 2>, k_global_gates<1, 3>, k_global_gates<1, 4>},
        {k_global_gates<2, 0>, k_global_gates<2, 1>, k_global_gates<2, 2>, k_global_gates<2, 3>, k_global_gates<2, 4>}};
    gfn fn = fns[nv - 1][ga];
    c->perf.link_bytes += (double)nv * 2.0 * 16.0 * (double)c->N * ((1 << ga) - 1) / (double)(1 << ga);   // in + out, this rank
    QR_LAUNCH(fn, grid, 256, 0, c->stream, gg);
    KERNEL_CHECK();
    c->perf.kernel_launches++;
    if (nv == 2) {
        QR_LAUNCH(k_reduce_partials, 1, QR_BLOCK, 0, c->stream, (const double*)c->d_scratch, grid, 4,
                  c->d_result + run->res_global + (size_t)layer * 4);
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    }
    return 0;
}

// H over this shard's local index for the two layouts of a sharded QAOA run (natural / swapped); rebuilt when the terms change
static int shard_build_ham_tables(qr_ctx* c, const qr_obs* o, const SwapRun* sr) {
    bool same = c->d_ham_ly[0] && c->ham_ly_terms.size() == o->terms.size();
    for (size_t k = 0; same && k < o->terms.size(); ++k)
        same = memcmp(&c->ham_ly_terms[k], &o->terms[k], sizeof(ObsTerm)) == 0;
    if (same) return 0;
    double wsum = 0.0;
    bool integral = true;
    for (const ObsTerm& t : o->terms) {
        if (t.kind < 2) return fail(QR_EINVAL, "the classical Hamiltonian of a QAOA circuit has z / zz terms only");
        wsum += std::fabs(t.w);
        if (t.w != std::floor(t.w)) integral = false;
    }
    c->ham_integer = integral && 2.0 * wsum + 1.0 <= (double)QR_LUT_MAX;
    c->ham_min = -wsum;
    c->ham_range = (int)(2.0 * wsum + 1.0);
    for (int ly = 0; ly < 2; ++ly) {
        if (sr->lay[ly].nl == 0) continue;   // (a run without exchange passes never reaches the swapped layout)
        if (!c->d_ham_ly[ly]) {
            if (cudaMalloc((void**)&c->d_ham_ly[ly], c->N * sizeof(double)) != cudaSuccess) { c->d_ham_ly[ly] = nullptr; return fail(QR_ENOMEM, "cannot allocate the Hamiltonian table"); }
            if (cudaMalloc((void**)&c->d_hidx_ly[ly], c->N * sizeof(short)) != cudaSuccess) { c->d_hidx_ly[ly] = nullptr; return fail(QR_ENOMEM, "cannot allocate the Hamiltonian index table"); }
        }
        std::vector<ObsTerm> terms = o->terms;
        QR_TRY(layout_remap_terms(sr->lay[ly], terms, nullptr));
        const ObsTerm* d_terms;
        QR_TRY(upload_terms(c, terms, &d_terms));
        const u64 off = (u64)layout_shard_value(sr->lay[ly], c->rank) << c->n;
        QR_LAUNCH(k_ham_build, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->d_ham_ly[ly], c->N, d_terms, (int)terms.size(), off);
        KERNEL_CHECK();
        if (c->ham_integer) {
            QR_LAUNCH(k_ham_index, grid_for(c, c->N), QR_BLOCK, 0, c->stream, (const double*)c->d_ham_ly[ly], c->d_hidx_ly[ly], c->N, c->ham_min);
            KERNEL_CHECK();
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    c->ham_ly_terms = o->terms;
    return 0;
}

// Qaoa.run_expec_val / grad_run (qaoa.py:23-70) on a sharded register: swap engine without a ladder; the diagonal phase
// exp(-i gamma H) rides on the first pass of a layer, 2 Im<lambda|H|psi> and the un-phase on the backward exchange pass
extern "C" int qr_shard_qaoa_begin(qr_ctx* c, int p, const double* betas, const double* gammas, const qr_obs* o, int want_grad,
                                   int* n_steps) {
    QR_TRY(need_shard(c));
    if (!o || o->n != c->n_total) return fail(QR_EINVAL, "observable must be defined on all %d qubits", c->n_total);
    if (p < 0 || (p > 0 && (!betas || !gammas)) || !n_steps) return fail(QR_EINVAL, "bad arguments");
    const int G = 1 << c->g;
    for (int r = 0; r < G; ++r)
        for (int b = 0; b < QR_NBUF; ++b)
            if (!c->peer[r][b]) return fail(QR_ESTATE, "peer buffers of rank %d are not mapped", r);
    if (!(swap_engine_ok(c->n, c->g) && c->opt_fusion && c->opt_tile_bits == 0 && c->opt_tile_bits_x == 0 && c->opt_min_row_bits == 3))
        return fail(QR_EINVAL, "sharded QAOA needs 1..3 rank bits, at least %d local qubits and default tile options", 12 + c->g);
    QR_TRY(use_device(c));
    delete c->run;
    c->run = nullptr;
    c->psi = 0;
    // the layouts are fixed by the register shape: plan once without tables to learn them, build H, then plan for real
    SwapCircuit circ = {1, p, nullptr, nullptr, betas, gammas};
    {
        SwapRun probe;
        probe.lockstep = true; probe.slices = 1; probe.gen_base = 0;
        std::vector<GateP> tab;
        const double b1[1] = {0.0};
        SwapCircuit pc = {1, 1, nullptr, nullptr, b1, b1};
        QR_TRY(swap_build(c, &probe, pc, o, false, &tab));
        QR_TRY(shard_build_ham_tables(c, o, &probe));
    }
    return shard_swap_begin(c, circ, o, want_grad != 0, n_steps);
}

extern "C" int qr_shard_step(qr_ctx* c, int step) {
    if (c) c->tables_fresh = true;   // every step is stream-synchronised: no programmatic launch across steps
    QR_TRY(need_shard(c));
    if (c->srun) {   // swap engine
        SwapRun* sr = c->srun;
        if (step < 0 || step >= (int)sr->step_first.size()) return fail(QR_EINVAL, "step %d out of range", step);
        QR_TRY(use_device(c));
        const int first = sr->step_first[step];
        const int last = step + 1 < (int)sr->step_first.size() ? sr->step_first[step + 1] : (int)sr->ops.size();
        for (int k = first; k < last; ++k) QR_TRY(swap_launch_op(c, sr, sr->ops[k], k));
        if (sr->lockstep) CUDA_TRY(cudaStreamSynchronize(c->stream));   // asynchronous mode: device-side flags order the ranks
        return 0;
    }
    ShardRun* run = c->run;
    if (!run) return fail(QR_ESTATE, "no sharded run in progress");
    if (step < 0 || step >= run->n_steps) return fail(QR_EINVAL, "step %d out of range", step);
    QR_TRY(use_device(c));
    const int L = run->L, P = run->P, nl = c->n, nt = c->n_total, GS = QR_GATE_SLOTS;
    const GateP* d_tab = (const GateP*)((char*)c->d_small + run->tab_off);
    const i64 stride = (i64)c->N;
    if (step < 2 * L) {
        const int i = step / 2;
        if (step % 2 == 0) {   // ---- forward, local passes of layer i ----
            if (i == 0) {
                double table[48];
                const double cs = std::cos(M_PI / 8.0), sn = std::sin(M_PI / 8.0);
                for (int w = 0; w <= nt; ++w) { double v = 1.0; for (int q = 0; q < nt; ++q) v *= (q < nt - w) ? cs : sn; table[w] = v; }
                const size_t off = c->small_cap - 512;
                QR_TRY(upload_small(c, off, table, sizeof(double) * (nt + 1), c->pin_cap - 512));
                const int pop = __builtin_popcount((unsigned)run->pi[c->rank]);
                QR_LAUNCH(k_init_product, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[c->psi], c->N,
                          (const double*)((char*)c->d_small + off) + pop, c->N - 1);
                KERNEL_CHECK();
            }
            for (int p = 0; p < P; ++p) {
                int dst = c->psi;
                LadderSpec spec;
                const LadderSpec* sp = nullptr;
                if (p == 0) {
                    u64 m1, m2;
                    ladder_masks(nt, 1, &m1, &m2);   // gather of ladder(0) uses the masks of ladder(1)
                    shard_relabel(c, run, m1, m2, &spec);
                    sp = &spec;
                    dst = other_buf(c, c->psi);
                }
                PassIO io = {c->buf[c->psi], nullptr, c->buf[dst], nullptr};
                QR_TRY(launch_pass(c, run->lpf, p, 1, io, d_tab + ((size_t)i * P + p) * GS, 0, -1, 1, stride, 0, nullptr, 0, 0, 0, 0,
                                   nullptr, sp));
                c->psi = dst;
            }
        } else {
            QR_TRY(shard_global_step(c, run, i, 1));
        }
    } else if (step == 2 * L) {   // ---- observable ----
        PeerTable peers;
        memset(&peers, 0, sizeof(peers));
        const int G = 1 << c->g;
        for (int rho = 0; rho < G; ++rho) peers.p[run->pi[rho]] = c->peer[rho][c->psi];
        run->lam = other_buf(c, c->psi);
        const int ogrid = grid_for(c, c->N);
        QR_TRY(ensure_scratch(c, ogrid));
        QR_LAUNCH(k_apply_obs, ogrid, QR_BLOCK, 0, c->stream, (const double2*)c->buf[c->psi],
                  run->want_grad ? c->buf[run->lam] : (double2*)nullptr, c->N, (const ObsTerm*)c->d_small, (int)run->terms.size(),
                  c->d_scratch, (u64)run->pi[c->rank] << nl, nl, peers);
        KERNEL_CHECK();
        QR_LAUNCH(k_reduce_partials, 1, QR_BLOCK, 0, c->stream, (const double*)c->d_scratch, ogrid, 1, c->d_result);
        KERNEL_CHECK();
    } else {
        const int t = step - (2 * L + 1);
        const int i = L - 1 - t / 2;
        if (t % 2 == 0) {   // ---- backward, local passes of layer i ----
            for (int p = 0; p < P; ++p) {
                int dpsi = c->psi, dlam = run->lam;
                LadderSpec spec;
                const LadderSpec* sp = nullptr;
                if (p == 0 && i < L - 1) {
                    u64 m1, m2;
                    ladder_masks(nt, 0, &m1, &m2);   // inverse ladder of layer i+1 (mc_clean.py:77)
                    shard_relabel(c, run, m1, m2, &spec);
                    sp = &spec;
                    dpsi = other_buf(c, c->psi, run->lam);
                    dlam = other_buf(c, c->psi, run->lam, dpsi);
                }
                PassIO io = {c->buf[c->psi], c->buf[run->lam], c->buf[dpsi], c->buf[dlam]};
                int units = 0;
                QR_TRY(launch_pass(c, run->lpb, p, 2, io, d_tab + ((size_t)(L + i) * P + p) * GS, 0, -1, 1, stride, 0, nullptr, 0, 0, 0,
                                   0, &units, sp, c->d_result + run->res_local + ((size_t)i * P + p) * QR_SLOTS));
                c->psi = dpsi;
                run->lam = dlam;
            }
        } else {
            QR_TRY(shard_global_step(c, run, i, 2));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

// partial sums of this rank: E and dE/d angles[L * n_total]; the caller adds them over ranks
extern "C" int qr_shard_mcclean_finish(qr_ctx* c, double* e_partial, double* grad_partial) {
    QR_TRY(need_shard(c));
    if (c->srun) {   // swap engine
        SwapRun* sr = c->srun;
        if (!e_partial) return fail(QR_EINVAL, "null output");
        QR_TRY(use_device(c));
        const int nt = c->n_total;
        unsigned long long* h_flags = (unsigned long long*)(c->h_pin + ((sr->n_results + 1) * sizeof(double) + 63) / 64 * 64);
        QR_TRY(ensure_pin(c, (sr->n_results + 16) * sizeof(double) + QR_FLAG_WORDS * sizeof(unsigned long long) + 128));
        h_flags = (unsigned long long*)(c->h_pin + ((sr->n_results + 1) * sizeof(double) + 63) / 64 * 64);
        CUDA_TRY(cudaMemcpyAsync(c->h_pin, c->d_result, sr->n_results * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(h_flags, c->d_flags, QR_FLAG_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int r = 0; r < (1 << c->g); ++r)
            if (h_flags[QR_FLAG_ERR * QR_MAX_RANKS + r]) {
                cudaMemsetAsync(c->d_flags + QR_FLAG_ERR * QR_MAX_RANKS, 0, QR_MAX_RANKS * sizeof(unsigned long long), c->stream);
                delete c->srun; c->srun = nullptr;
                return fail(QR_ESTATE, "sharded run: timed out waiting for rank %d (generation %llu)", r, h_flags[QR_FLAG_ERR * QR_MAX_RANKS + r]);
            }
        swap_collect_times(c, sr);
        c->psi = sr->final_psi;
        const double* res = (const double*)c->h_pin;
        *e_partial = res[0];
        if (sr->circuit == 1) {   // QAOA: grad[2 i] = sum of the X-generator slots of layer i, grad[2 i + 1] = 2 x the H-generator slot (qaoa.py:62-68)
            if (grad_partial) {
                for (int i = 0; i < 2 * sr->layers; ++i) grad_partial[i] = 0.0;
                for (const XOp& op : sr->ops)
                    if (op.kind == 1 && op.nv == 2 && op.layer >= 0)
                        for (int sl = 0; sl < op.nslices; ++sl) {
                            const double* r2 = res + op.res_off + (size_t)sl * QR_SLOTS;
                            for (int s2 = 0; s2 < QR_GATE_SLOTS; ++s2)
                                if (op.slot_qubit[s2] >= 0) grad_partial[2 * op.layer] += r2[s2];
                            if (op.post_phase) grad_partial[2 * op.layer + 1] += 2.0 * r2[QR_SLOTS - 1];
                        }
            }
            delete c->srun;
            c->srun = nullptr;
            return 0;
        }
        if (grad_partial) {
            for (size_t i = 0; i < (size_t)sr->layers * nt; ++i) grad_partial[i] = 0.0;
            for (const XOp& op : sr->ops)
                if (op.kind == 1 && op.nv == 2) {
                    for (int s2 = 0; s2 < QR_GATE_SLOTS; ++s2)
                        if (op.slot_qubit[s2] >= 0) {
                            double v = 0.0;
                            for (int sl = 0; sl < op.nslices; ++sl) v += res[op.res_off + (size_t)sl * QR_SLOTS + s2];
                            grad_partial[(size_t)op.layer * nt + op.slot_qubit[s2]] = v;
                        }
                }
        }
        delete c->srun;
        c->srun = nullptr;
        return 0;
    }
    ShardRun* run = c->run;
    if (!run || !e_partial) return fail(QR_ESTATE, "no sharded run in progress");
    QR_TRY(use_device(c));
    const int L = run->L, P = run->P, nt = c->n_total;
    CUDA_TRY(cudaMemcpyAsync(c->h_pin, c->d_result, run->res_total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const double* res = (const double*)c->h_pin;
    *e_partial = res[0];
    if (run->want_grad && grad_partial) {
        for (size_t i = 0; i < (size_t)L * nt; ++i) grad_partial[i] = 0.0;
        for (int i = 0; i < L; ++i) {
            for (int p = 0; p < P; ++p)
                for (int s = 0; s < QR_GATE_SLOTS; ++s) {
                    const int gb = run->lpb.pass[p].gbit[s];
                    if (gb >= 0) grad_partial[(size_t)i * nt + (nt - 1 - gb)] = res[run->res_local + ((size_t)i * P + p) * QR_SLOTS + s];
                }
            for (int b = 0; b < c->g; ++b) grad_partial[(size_t)i * nt + (c->g - 1 - b)] = res[run->res_global + (size_t)i * 4 + b];
        }
        c->psi = run->lam;
    }
    delete c->run;
    c->run = nullptr;
    return 0;
}

extern "C" int qr_shard_info(qr_ctx* c, int* n_total, int* log2_world, int* rank, int* logical_shard) {
    QR_TRY(need_shard(c));
    if (n_total) *n_total = c->n_total;
    if (log2_world) *log2_world = c->g;
    if (rank) *rank = c->rank;
    if (logical_shard) *logical_shard = c->run ? c->run->pi[c->rank] : c->rank;
    return 0;
}
