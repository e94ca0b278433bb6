// Sharded registers, "swap" engine: ONE NVLink crossing per exchange, fused into an ordinary tile pass.
//
// Included by qr_lib.cu (single translation unit; uses its planner, launch_pass and error helpers).
//
// A register of n qubits is sharded over G = 2^g ranks; rank rho holds 2^nl amplitudes (nl = n - g).  Which logical
// index bit a physical bit (local bit k, or rank bit b) holds is the LAYOUT, kept as a GF(2)-linear map Phi from the
// physical index x = (rho << nl) | l to the logical index j.  Every step of a McClean layer is then a tile pass whose
// SOURCE is given by a linear map M (destination physical index -> source physical index):
//
//   * CNOT ladder (mc_clean.py:39,77): gather through M0 = Phi^-1 G Phi, G the banded ladder map (state.py:229-241).
//     The rank-rank block D of M0 is absorbed into the layout (Phi' = Phi T, T = diag(1, D^-1): a relabelling of the
//     shards, no data moves).  In the natural layout the source rank never depends on local bits; in the swapped layout
//     it depends on one or two tile-index bits of the contiguous pass (the CNOT whose control is local and whose target
//     sits on a rank bit), i.e. some tiles are read from one other shard over NVLink, fused into the pass.
//   * rotations on local qubits: the same in-place tile passes as on one GPU.
//   * rotations on the qubits held by the rank bits: the EXCHANGE PASS.  Its destination tile is a strided tile of this
//     rank's own memory whose top g gate bits are local bits [sigma, sigma+g); its source is M = Phi^-1 Phi_new where
//     Phi_new swaps the rank bits with those local bits.  The three index bits a thread holds in registers at load time
//     then select the source SHARD, so every thread loads its 8 amplitudes per vector from (up to) 8 peers through a
//     per-register pointer table, applies the gates of the formerly global qubits (and of 9 - g local ones) like any other
//     pass -- gradient inner products included -- and stores to local memory.  Each exchanged amplitude crosses NVLink
//     once; the layout alternates between "qubits 0..g-1 on the rank bits" and "the qubits of [sigma, sigma+g) on the
//     rank bits"; a 30-qubit register on 8 GPUs needs 3 sweeps per layer (12 | 9 | exchange 9), like on one GPU.
//
// Cross-rank ordering.  A pass that reads peer memory starts a new STEP.  Lockstep mode (one process drives all shards,
// and the CPU test tier): the caller finishes a step on every rank before starting the next.  Asynchronous mode (one
// process per GPU): everything is enqueued at once and device-side flags in peer-mapped memory order the streams --
// READY(gen) is raised after the passes that produce the data a remote pass reads, DONE(gen) after the remote pass; a
// rank waits for READY from all ranks before its remote pass and for DONE before it overwrites a buffer peers read.
#pragma once

// ------------------------------------------------------------------------------------------
// GF(2) linear maps on index bits
// ------------------------------------------------------------------------------------------
#define QR_LIN_MAX 40
struct Lin {
    int n;
    u64 col[QR_LIN_MAX];   // col[k] = image of the unit vector e_k
};

static Lin lin_identity(int n) {
    Lin a;
    a.n = n;
    for (int k = 0; k < QR_LIN_MAX; ++k) a.col[k] = k < n ? (u64)1 << k : 0;
    return a;
}
static u64 lin_apply(const Lin& a, u64 x) {
    u64 y = 0;
    for (int k = 0; k < a.n; ++k)
        if ((x >> k) & 1) y ^= a.col[k];
    return y;
}
static Lin lin_mul(const Lin& a, const Lin& b) {   // x -> a(b(x))
    Lin c = lin_identity(a.n);
    for (int k = 0; k < a.n; ++k) c.col[k] = lin_apply(a, b.col[k]);
    return c;
}
static bool lin_inverse(const Lin& a, Lin* out) {
    const int n = a.n;
    u64 row[QR_LIN_MAX], inv[QR_LIN_MAX];   // row r of a / of the inverse, as bit masks over columns
    for (int r = 0; r < n; ++r) {
        row[r] = 0;
        for (int k = 0; k < n; ++k) row[r] |= ((a.col[k] >> r) & 1) << k;
        inv[r] = (u64)1 << r;
    }
    for (int k = 0; k < n; ++k) {
        int piv = -1;
        for (int r = k; r < n; ++r)
            if ((row[r] >> k) & 1) { piv = r; break; }
        if (piv < 0) return false;
        std::swap(row[k], row[piv]);
        std::swap(inv[k], inv[piv]);
        for (int r = 0; r < n; ++r)
            if (r != k && ((row[r] >> k) & 1)) { row[r] ^= row[k]; inv[r] ^= inv[k]; }
    }
    *out = lin_identity(n);
    for (int k = 0; k < n; ++k) {
        out->col[k] = 0;
        for (int r = 0; r < n; ++r) out->col[k] |= ((inv[r] >> k) & 1) << r;
    }
    return true;
}
// gather map of ladder(stacking) on n index bits: dest index -> source index (the masks of the inverse ladder)
static Lin lin_ladder_gather(int n, int stacking) {
    u64 m1, m2;
    ladder_masks(n, 1 - stacking, &m1, &m2);
    Lin g = lin_identity(n);
    for (int k = 0; k < n; ++k) g.col[k] = ladder_map((u64)1 << k, m1, m2);
    return g;
}
static int unit_bit(u64 v) {   // index of the single set bit, or -1
    if (v == 0 || (v & (v - 1))) return -1;
    int b = 0;
    while (!((v >> b) & 1)) ++b;
    return b;
}

// ------------------------------------------------------------------------------------------
// device-side flags
// ------------------------------------------------------------------------------------------
#define QR_FLAG_READY 0
#define QR_FLAG_DONE 1
#define QR_FLAG_ERR 2
#define QR_FLAG_WORDS (3 * QR_MAX_RANKS)

struct FlagPeers { unsigned long long* p[QR_MAX_RANKS]; };

// raise flag `kind` of rank `me` to `value` in every rank's flag array (remote stores over NVLink)
__global__ void k_flag_signal(FlagPeers peers, int nranks, int me, int kind, unsigned long long value) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
#ifndef QR_HOST_EMUL
    __threadfence_system();
    volatile unsigned long long* f = peers.p[r] + kind * QR_MAX_RANKS + me;
    *f = value;
    __threadfence_system();
#else
    __atomic_store_n(peers.p[r] + kind * QR_MAX_RANKS + me, value, __ATOMIC_SEQ_CST);
#endif
}

// spin until every rank's flag `kind` in the LOCAL flag array has reached `value` (bounded: a dead peer must not hang the GPU)
__global__ void k_flag_wait(unsigned long long* flags, int nranks, int kind, unsigned long long value) {
    const int r = threadIdx.x;
    if (r >= nranks) return;
#ifndef QR_HOST_EMUL
    volatile unsigned long long* f = flags + kind * QR_MAX_RANKS + r;
    volatile unsigned long long* err = flags + QR_FLAG_ERR * QR_MAX_RANKS;
    const long long t0 = clock64();
    while (*f < value) {
        bool dead = false;   // a wait that already timed out on this rank ends every later wait at once (reported at finish)
        for (int q = 0; q < nranks; ++q) dead = dead || err[q] != 0;
        if (dead) break;
        if (clock64() - t0 > 60000000000ll) { err[r] = value; break; }   // ~30 s
        __nanosleep(200);
    }
    __threadfence_system();
#else
    unsigned long long spins = 0;
    while (__atomic_load_n(flags + kind * QR_MAX_RANKS + r, __ATOMIC_SEQ_CST) < value) {
        if (++spins > 4000000000ull) { flags[QR_FLAG_ERR * QR_MAX_RANKS + r] = value; break; }
        emul_yield_cpu();
    }
#endif
}

// ------------------------------------------------------------------------------------------
// schedule
// ------------------------------------------------------------------------------------------
struct Layout {
    Lin phi;         // physical (rho << nl | l) -> logical index
    int nl, g;
};

struct XOp {
    int kind = 0;            // 0: initial state, 1: tile pass, 2: observable (lambda = O psi), 3: QAOA co-state (lambda = H psi)
    int nv = 1;
    int layer = 0;           // layer whose rotations the pass applies (gate table, result slots)
    PassPlan pp;             // destination tile geometry; gbit = destination local bit of each gate slot (or -1)
    bool xmap = false;       // general source map (k_tile12_x) instead of the banded ladder masks
    bool gather = false;     // plain path: out-of-place gather through `spec`
    LadderSpec spec = {0, 0, 0};
    u64 tcol[24], lcol[9], roff[8], src_const = 0;
    int sel_bit[2] = {-1, -1};   // selector bit k: tile-index bit ...
    int sel_pos[2] = {-1, -1};   //   (its destination local bit: the tile-index bit depends on the enumeration of a sliced launch)
    int thr_bit[2] = {-1, -1};   // ... or thread bit (tile-local bit < 9)
    Lin M;                       // source map (dest physical index -> source physical index)
    int nslices = 1;             // > 1: issued in slices of the index bits [9, 9 + log2 nslices) (last local pass + exchange pass of a layer)
    long long pair_gen = 0;      // sliced local pass: first generation of its exchange pass (raises READY slice by slice)
    int src_rank[4][8];      // physical source rank per (selector, register)
    bool remote = false;
    double remote_frac = 0.0;   // fraction of the source amplitudes read from peers
    int src_buf[2] = {-1, -1}, dst_buf[2] = {-1, -1};
    size_t gate_off = 0;     // entries into the gate table
    size_t res_off = 0;      // doubles into d_result (backward passes: QR_SLOTS sums)
    int slot_qubit[QR_GATE_SLOTS];
    // observable
    std::vector<ObsTerm> terms;
    u64 obs_base = 0;
    int obs_peer_rank[QR_MAX_RANKS];   // physical rank holding logical shard value u
    // init
    int init_pop = 0;
    int init_plus = 0;       // 1: |+..+> (QAOA) instead of the Ry(pi/4) product state
    // QAOA diagonal phase exp(-i angle H) before the gates (forward) / generator inner product + un-phase after them (backward)
    int pre_phase = 0, post_phase = 0;
    double angle_pre = 0.0, angle_post = 0.0;
    int ham_layout = -1;     // which of the two per-layout H tables (0 natural, 1 swapped) the pass / the co-state step uses
    int lut_index = -1;      // phase look-up table of this (direction, layer)
    // ordering
    bool new_step = false;   // a pass that reads peer memory starts a step (and so does the pass after it)
    long long wait_done = 0; // asynchronous mode: wait for DONE >= this generation before launching (0: none)
    long long gen = 0;       // generation of this remote pass (READY / DONE value)
};

struct SwapRun {
    std::vector<XOp> ops;
    std::vector<int> step_first;   // first op of each step
    int sigma = 0, h = 0, m = 0;
    int sweeps = 0, layers = 0;
    int slices = 1;
    int circuit = 0;               // 0 McClean, 1 QAOA
    int final_psi = 0;             // buffer that holds the state when the run ends
    Layout lay[2];                 // QAOA: the natural and the swapped layout (H tables are built per layout)
    size_t tab_off = 0, n_results = 0;
    size_t lut_off = 0;            // QAOA: phase look-up tables [2 L][ham_range] in d_small (0: none)
    bool lockstep = true;
    long long gen_base = 0;
    std::vector<cudaEvent_t> ev;   // two events per op (after its waits / after its kernels): per-kind device times at finish
    ~SwapRun() { for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e); }
};

// layout bookkeeping -------------------------------------------------------------------------------------

// logical bit held by physical local bit k (-1 if the column is not a unit vector)
static int layout_local_logical(const Layout& ly, int k) { return unit_bit(ly.phi.col[k]); }

// the g logical positions held by the rank bits, ascending
static void layout_held(const Layout& ly, int* held) {
    u64 used = 0;
    for (int k = 0; k < ly.nl; ++k) used |= ly.phi.col[k];
    int b = 0;
    for (int p = 0; p < ly.nl + ly.g; ++p)
        if (!((used >> p) & 1) && b < ly.g) held[b++] = p;
}

// logical value u (bit b = logical bit held[b]) of the shard held by physical rank rho
static int layout_shard_value(const Layout& ly, int rho) {
    int held[8];
    layout_held(ly, held);
    const u64 j = lin_apply(ly.phi, (u64)rho << ly.nl);
    int u = 0;
    for (int b = 0; b < ly.g; ++b) u |= (int)((j >> held[b]) & 1) << b;
    return u;
}

// observable / Hamiltonian terms with their bit positions translated from the logical index to (u << nl) | l, where l is
// this shard's local index in layout ly and u the logical value of the rank-held bits (bit b = logical bit held[b])
static int layout_remap_terms(const Layout& ly, std::vector<ObsTerm>& terms, bool* needs_peer) {
    int held[8];
    layout_held(ly, held);
    if (needs_peer) *needs_peer = false;
    for (ObsTerm& t : terms) {
        auto remap = [&](int p) -> int {
            for (int k = 0; k < ly.nl; ++k)
                if (layout_local_logical(ly, k) == p) return k;
            for (int b = 0; b < ly.g; ++b)
                if (held[b] == p) return ly.nl + b;
            return -1;
        };
        t.bit_i = remap(t.bit_i);
        if (t.kind == QR_TERM_ZZ) t.bit_j = remap(t.bit_j);
        if (t.bit_i < 0 || (t.kind == QR_TERM_ZZ && t.bit_j < 0)) return fail(QR_ESTATE, "internal: observable qubit not found in the layout");
        if (needs_peer && (t.kind == QR_TERM_X || t.kind == QR_TERM_Y) && t.bit_i >= ly.nl) *needs_peer = true;
    }
    return 0;
}

// fill the source description of a pass with destination geometry `pp` (k = 12) and source map M (dest -> source, physical)
static int xop_set_map(qr_ctx* c, XOp& op, const Lin& M, int nl, int g) {
    const int me = c->rank, G = 1 << g;
    const u64 lmask = ((u64)1 << nl) - 1;
    const PassPlan& pp = op.pp;
    const Geo12 geo = {pp.c, pp.h, pp.m1, pp.h2, 12, 0};
    const u64 cimg = lin_apply(M, (u64)me << nl);
    op.M = M;
    op.src_const = cimg & lmask;
    const int r0 = (int)(cimg >> nl);
    int rr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int nsel = 0, srank[2] = {0, 0};
    for (int b = 0; b < 12; ++b) {   // tile-local bits
        const u64 d = geo12_local(geo, (u64)1 << b);
        const u64 img = lin_apply(M, d);
        if (b < 9) {
            op.lcol[b] = img & lmask;
            if (img >> nl) {   // the source shard depends on a thread bit: a per-thread choice of the pointer table
                if (nsel >= 2) return fail(QR_ESTATE, "internal: source rank depends on more than two selector bits");
                op.thr_bit[nsel] = b;
                srank[nsel] = (int)(img >> nl);
                ++nsel;
            }
        } else {
            for (int r = 0; r < 8; ++r)
                if ((r >> (b - 9)) & 1) { op.roff[r] ^= img & lmask; rr[r] ^= (int)(img >> nl); }
        }
    }
    const int tiles_log2 = nl - 12;
    if (tiles_log2 > 24) return fail(QR_EINVAL, "sharded registers: at most 36 local qubits");
    for (int j = 0; j < 24; ++j) {
        op.tcol[j] = 0;
        if (j >= tiles_log2) continue;
        const u64 d = geo12_tile(geo, (u64)1 << j);
        const u64 img = lin_apply(M, d);
        op.tcol[j] = img & lmask;
        if (img >> nl) {
            if (nsel >= 2) return fail(QR_ESTATE, "internal: source rank depends on more than two selector bits");
            op.sel_bit[nsel] = j;
            op.sel_pos[nsel] = unit_bit(d);
            srank[nsel] = (int)(img >> nl);
            ++nsel;
        }
    }
    int nremote = 0;
    for (int s = 0; s < 4; ++s)
        for (int r = 0; r < 8; ++r) {
            const int rk = r0 ^ rr[r] ^ ((s & 1) ? srank[0] : 0) ^ ((s & 2) ? srank[1] : 0);
            if (rk < 0 || rk >= G) return fail(QR_ESTATE, "internal: source rank out of range");
            op.src_rank[s][r] = rk;
            const bool live = (s < (1 << nsel));
            if (live && rk != me) ++nremote;
        }
    op.remote = nremote > 0;
    op.remote_frac = (double)nremote / (double)(8 << nsel);
    // plain path: all local and banded (or identity)
    op.xmap = true;
    if (!op.remote) {
        u64 m1 = 0, m2 = 0;
        bool banded = true;
        for (int k = 0; k < nl && banded; ++k) {
            const u64 img = lin_apply(M, (u64)1 << k);
            if (img >> nl) { banded = false; break; }
            u64 rest = img ^ ((u64)1 << k);
            if (!((img >> k) & 1)) { banded = false; break; }
            if (k >= 1 && ((rest >> (k - 1)) & 1)) { m1 |= (u64)1 << (k - 1); rest ^= (u64)1 << (k - 1); }
            if (k >= 2 && ((rest >> (k - 2)) & 1)) { m2 |= (u64)1 << (k - 2); rest ^= (u64)1 << (k - 2); }
            if (rest) banded = false;
        }
        if (banded) {
            op.xmap = false;
            op.gather = (m1 | m2 | op.src_const) != 0;
            op.spec.M1 = m1; op.spec.M2 = m2; op.spec.src_xor = op.src_const;
        }
    }
    return 0;
}

// strided / contiguous destination tile for the local gate bits `bits` (ascending; at most two contiguous runs)
static int xop_set_geometry(XOp& op, const std::vector<int>& bits, bool contiguous) {
    PassPlan& pp = op.pp;
    memset(&pp, 0, sizeof(pp));
    pp.k = 12;
    if (contiguous) {
        pp.c = 12; pp.h = 12; pp.m1 = 0; pp.h2 = 12;
        plan_lean(pp, 0);
        // keep only the requested bits
        for (int s = 0; s < QR_GATE_SLOTS; ++s) {
            bool keep = false;
            for (int b : bits) keep = keep || pp.gbit[s] == b;
            if (!keep) pp.gbit[s] = -1;
        }
        return 0;
    }
    const int m = (int)bits.size();
    if (m < 1 || m > 9) return fail(QR_ESTATE, "internal: strided pass with %d gate bits", m);
    int run1 = 1;
    while (run1 < m && bits[run1] == bits[run1 - 1] + 1) ++run1;
    for (int i = run1 + 1; i < m; ++i)
        if (bits[i] != bits[i - 1] + 1) return fail(QR_ESTATE, "internal: gate bits of a strided pass form more than two runs");
    pp.c = 12 - m;
    pp.h = bits[0];
    pp.m1 = run1;
    pp.h2 = run1 < m ? bits[run1] : pp.h + run1;
    if (pp.h < pp.c) return fail(QR_ESTATE, "internal: strided gate bits overlap the tile rows");
    plan_lean(pp, pp.c);
    return 0;
}

static bool swap_engine_ok(int nl, int g) { return g >= 1 && g <= 3 && nl >= 12 + g && nl <= 36; }

// CNOT(control bit pc -> target bit pt) as an index map: j -> j ^ (j_pc ? e_pt : 0) (its own inverse)
static Lin lin_cnot(int n, int pc, int pt) {
    Lin f = lin_identity(n);
    f.col[pc] ^= (u64)1 << pt;
    return f;
}

// the CNOT of the ladder whose control is held by a local bit and whose target by the rank bits (at most one in the two
// layouts the engine uses); returns false if there is none
static bool layout_offending_cnot(const Layout& ly, int* pc_out, int* pt_out) {
    const int nt = ly.nl + ly.g;
    u64 local = 0;
    for (int k = 0; k < ly.nl; ++k) local |= ly.phi.col[k];
    int found = 0;
    for (int c = 0; c + 1 < nt; ++c) {
        const int pc = nt - 1 - c, pt = nt - 2 - c;
        if (((local >> pc) & 1) && !((local >> pt) & 1)) { *pc_out = pc; *pt_out = pt; ++found; }
    }
    return found == 1;
}

// does the gather G, run in layout ly, read only this rank's own shard (after the shard relabelling)?
static bool gather_is_local(const Layout& ly, const Lin& G) {
    Lin inv;
    if (!lin_inverse(ly.phi, &inv)) return false;
    const Lin M0 = lin_mul(inv, lin_mul(G, ly.phi));
    for (int k = 0; k < ly.nl; ++k)
        if (M0.col[k] >> ly.nl) return false;
    return true;
}

// Build the whole schedule of one gradient (or forward-only run).  Deterministic given (n, g, L): every rank builds the
// same sequence with the same buffer indices and generations.
//
// The CNOT of a ladder whose control is local and whose target sits on a rank bit (swapped layout only) would make the
// ladder pass read half of its tiles from another shard.  Where the ladder can be written with that CNOT acting first
// (forward) / last (backward), it is PEELED off and folded into the load addresses of the neighbouring exchange pass,
// which reads the peers anyway:
//   forward : the exchange pass that leaves the natural layout applies it on the way (both qubits are still local there);
//   backward: the exchange pass that returns to the natural layout applies it as a permutation of its source pointers and
//             also takes over the un-rotation of the CNOT's control qubit (which must follow the CNOT), so that pass's
//             tile holds the swapped bits, the control bit right above them and 8 - g other local bits.
// Every exchanged amplitude then crosses NVLink exactly once per layer and vector.
struct SwapCircuit {
    int kind;                                   // 0 McClean, 1 QAOA
    int L;
    const int32_t* axes; const double* angles;  // McClean [L * n]
    const double* betas; const double* gammas;  // QAOA [L]
};

static int swap_build(qr_ctx* c, SwapRun* sr, const SwapCircuit& circ, const qr_obs* o, bool want_grad,
                      std::vector<GateP>* gate_tab) {
    const int nl = c->n, g = c->g, nt = c->n_total, G = 1 << g, me = c->rank;
    const int L = circ.L;
    const int32_t* axes = circ.axes;
    const double* angles = circ.angles;
    const bool qaoa = circ.kind == 1;
    sr->circuit = circ.kind;
    const int m = 9 - g;                                   // local gate bits of an exchange pass
    // gate run of the exchange pass: [h, h+9) = [h, h+m) local + [sigma, sigma+g) swapped.  One local bit must stay above
    // it (the control of the peeled CNOT), and the peel needs that control to be an odd qubit (it then belongs to the
    // CNOT group of ladder(0) that acts first): pick h accordingly where the register is large enough.
    int h = std::min(12, nl - 9);
    bool peel_geometry = false;
    {
        const int hmax = std::min(13, nl - 10);
        for (int cand = hmax; cand >= std::max(4, hmax - 1); --cand) {
            const int ctrl_qubit = nt - 1 - (cand + 9);    // qubit held by local bit sigma + g in the natural layout
            if (ctrl_qubit >= 0 && (ctrl_qubit & 1) && cand + 9 < nl) { h = cand; peel_geometry = true; break; }
        }
    }
    const int sigma = h + m;
    sr->h = h; sr->m = m; sr->sigma = sigma; sr->layers = L;
    Layout ly;
    ly.phi = lin_identity(nt);
    ly.nl = nl; ly.g = g;
    std::vector<int> xbits;
    for (int k = h; k < h + 9; ++k) xbits.push_back(k);
    sr->ops.clear();
    gate_tab->clear();
    int psi = 0, lam = -1;
    long long gen = sr->gen_base;
    long long readable[QR_NBUF] = {0, 0, 0, 0};   // generation of the last remote pass in which peers read this buffer
    long long waited = sr->gen_base;              // DONE generation already waited for
    size_t res = 1;                               // d_result[0] = E
    int max_sweeps = 0;
    auto pick_free = [&](int a, int b, int d) { for (int i = 0; i < QR_NBUF; ++i) if (i != a && i != b && i != d) return i; return -1; };
    auto finish_op = [&](XOp& op, const Layout& dest_ly) -> int {
        // gate table entries and slot -> qubit map in the destination layout
        op.gate_off = gate_tab->size();
        const double sgn = op.nv == 2 ? -1.0 : 1.0;
        for (int s = 0; s < QR_GATE_SLOTS; ++s) {
            GateP gp; gp.c = 1.0; gp.s = 0.0; gp.axis = -1; gp.pad = 0;
            op.slot_qubit[s] = -1;
            if (op.pp.gbit[s] >= 0) {
                const int p = layout_local_logical(dest_ly, op.pp.gbit[s]);
                if (p < 0) return fail(QR_ESTATE, "internal: a gate bit does not hold a single logical qubit");
                const int q = nt - 1 - p;
                if (op.layer >= 0) {
                    const double an = qaoa ? circ.betas[op.layer] : angles[(size_t)op.layer * nt + q];
                    gp.c = std::cos(0.5 * an); gp.s = sgn * std::sin(0.5 * an); gp.axis = qaoa ? 0 : axes[(size_t)op.layer * nt + q];
                    op.slot_qubit[s] = q;
                    if (op.pp.gx && ((op.pp.zmask >> s) & 1)) gp.pad = 1 + op.pp.gbit[s];   // Rz applied through the tile's own index bit
                }
            }
            gate_tab->push_back(gp);
        }
        if (op.nv == 2) { op.res_off = res; res += (size_t)QR_SLOTS * op.nslices; }
        // ordering (a sliced exchange pass takes one generation per slice)
        if (op.remote) { op.new_step = true; op.gen = gen + 1; gen += op.nslices; }
        for (int v = 0; v < op.nv; ++v) {
            const int b = op.dst_buf[v];
            if (readable[b] > waited) { op.wait_done = std::max(op.wait_done, readable[b]); }
        }
        if (op.wait_done > waited) waited = op.wait_done; else op.wait_done = 0;
        if (op.remote)
            for (int v = 0; v < op.nv; ++v) readable[op.src_buf[v]] = op.gen + op.nslices - 1;
        return 0;
    };
    auto new_op = [&](int layer, int nv) {
        XOp op;
        op.kind = 1; op.nv = nv; op.layer = layer;
        memset(op.tcol, 0, sizeof(op.tcol)); memset(op.lcol, 0, sizeof(op.lcol)); memset(op.roff, 0, sizeof(op.roff));
        return op;
    };
    // One layer.  ladder_stacking >= 0: the gather of that ladder is folded into pass 0.  `pre` (forward): a CNOT already
    // applied by the previous exchange pass, to be taken out of this ladder.  `next_stacking` (forward): the ladder of the
    // next layer, whose offending CNOT this layer's exchange pass may apply in advance.
    std::vector<Lin> pre;   // 0 or 1 entries
    bool swapped = false;   // QAOA (no ladder, no relabelling): which of its two layouts is current
    sr->lay[0] = ly;
    sr->lay[1].nl = 0;
    auto add_layer = [&](int layer, int nv, int ladder_stacking, int next_stacking) -> int {
        bool owed = false;          // backward: a CNOT peeled off this layer's ladder, owed to this layer's exchange pass
        Lin owed_f = lin_identity(nt);
        // ---- the ladder in the current layout ----
        Lin M = lin_identity(nt);
        if (ladder_stacking >= 0 && nt >= 2 && !qaoa) {
            Lin Gl = lin_ladder_gather(nt, ladder_stacking);
            if (!pre.empty()) { Gl = lin_mul(pre[0], Gl); pre.clear(); }          // G = F G_rest  =>  G_rest = F G
            int pc, pt;
            if (nv == 2 && peel_geometry && !gather_is_local(ly, Gl) && layout_offending_cnot(ly, &pc, &pt)) {
                const Lin F = lin_cnot(nt, pc, pt);
                const Lin Grest = lin_mul(Gl, F);                                  // inverse ladder = rest first, then the CNOT
                int kc = -1;
                for (int k = 0; k < nl; ++k) if (layout_local_logical(ly, k) == pc) kc = k;
                if (gather_is_local(ly, Grest) && kc == sigma + g && sigma - 2 >= 3 && h - 1 >= 3) {
                    owed = true; owed_f = F; Gl = Grest;
                }
            }
            Lin inv;
            if (!lin_inverse(ly.phi, &inv)) return fail(QR_ESTATE, "internal: singular layout");
            const Lin M0 = lin_mul(inv, lin_mul(Gl, ly.phi));
            // absorb the rank-rank block into the layout: T = diag(1, D^-1)
            Lin D = lin_identity(g), Dinv;
            for (int cb = 0; cb < g; ++cb) D.col[cb] = (M0.col[nl + cb] >> nl) & (u64)(G - 1);
            if (!lin_inverse(D, &Dinv)) return fail(QR_ESTATE, "internal: singular shard relabelling");
            Lin T = lin_identity(nt);
            for (int cb = 0; cb < g; ++cb) T.col[nl + cb] = Dinv.col[cb] << nl;
            ly.phi = lin_mul(ly.phi, T);
            M = lin_mul(M0, T);
        } else pre.clear();
        // ---- gate bits of this layer's passes ----
        std::vector<int> xb;        // exchange pass (destination layout)
        if (!owed) xb = xbits;
        else {   // [h-1, sigma-2) + [sigma, sigma+g] : 8 - g local bits, the swapped bits, the control bit of the owed CNOT
            for (int k = h - 1; k < sigma - 2; ++k) xb.push_back(k);
            for (int k = sigma; k <= sigma + g; ++k) xb.push_back(k);
        }
        std::vector<int> low, high;
        for (int k = 0; k < nl; ++k) {
            bool in_x = false;
            for (int b : xb) in_x = in_x || b == k;
            // the exchange pass rotates what its destination tile holds AFTER the swap: the swapped bits [sigma, sigma+g)
            // still need their local rotation in the current layout
            if (in_x && !(k >= sigma && k < sigma + g)) continue;
            (k < 12 ? low : high).push_back(k);
        }
        std::vector<std::vector<int>> strided;
        std::vector<PassPlan> strided_gx;   // axis-aware plans of the strided local passes (McClean; qr_lib.cu: plan_axis_layer)
        if (!high.empty() && !qaoa && layer >= 0 && c->opt_axis_plan && (nl >= 20 || c->axis_plan_forced) && sr->slices <= 1 && nl - 12 >= 3) {
            int nzb[64], zsb[64], k = 0, nzs = 0;
            for (int kb : high) {
                const int p = layout_local_logical(ly, kb);
                if (p < 0) { k = -1; break; }
                if (axes[(size_t)layer * nt + (nt - 1 - p)] == 2) zsb[nzs++] = kb; else nzb[k++] = kb;
            }
            const int nx = ((int)high.size() + 8) / 9;
            for (int mm = 1; mm <= nx && k >= 0 && strided_gx.empty(); ++mm) {
                int split[8];
                if (mm * 9 < k) continue;
                const double cst = axis_best_split(k, mm, 9, 12, split);
                if (cst > 1e29 || (mm == nx && cst >= 16.0 * nx)) continue;
                PassPlan outp[8];
                if (axis_build_strided(12, nl, 12, nzb, k, zsb, nzs, mm, split, outp)) strided_gx.assign(outp, outp + mm);
            }
        }
        if (!high.empty() && strided_gx.empty()) {
            const int nx = ((int)high.size() + 8) / 9;
            size_t pos = 0;
            for (int i = 0; i < nx; ++i) {
                const int sz = (int)high.size() / nx + (i < (int)high.size() % nx ? 1 : 0);
                strided.emplace_back(high.begin() + pos, high.begin() + pos + sz);
                pos += sz;
            }
        }
        const size_t n_strided = strided_gx.empty() ? strided.size() : strided_gx.size();
        max_sweeps = std::max(max_sweeps, 2 + (int)n_strided);
        // ---- pass 0: contiguous tile, gather through the ladder ----
        if (layer >= 0) {
            XOp op = new_op(layer, nv);
            QR_TRY(xop_set_geometry(op, low, true));
            QR_TRY(xop_set_map(c, op, M, nl, g));
            if (qaoa && nv == 1 && layer >= 0) {   // exp(-i gamma H) before the mixer (qaoa.py:51), H in the current layout
                op.pre_phase = 1; op.angle_pre = circ.gammas[layer]; op.ham_layout = swapped ? 1 : 0; op.lut_index = layer;
            }
            const bool oop = op.xmap || op.gather;
            op.src_buf[0] = psi; op.src_buf[1] = lam;
            if (oop) {
                const int d0 = pick_free(psi, lam, -1);
                const int d1 = nv == 2 ? pick_free(psi, lam, d0) : -1;
                op.dst_buf[0] = d0; op.dst_buf[1] = d1;
            } else { op.dst_buf[0] = psi; op.dst_buf[1] = lam; }
            QR_TRY(finish_op(op, ly));
            psi = op.dst_buf[0]; if (nv == 2) lam = op.dst_buf[1];
            sr->ops.push_back(op);
        }
        // ---- strided local passes, in place ----
        // The last one and the exchange pass can be issued in slices of the index bits [9, 12) when those bits are tile-index
        // bits of both (rows below 9, gate runs from 12 up): the exchange of a slice then overlaps the local pass of the next.
        bool slice_pair = sr->slices > 1 && strided_gx.empty() && !qaoa && layer >= 0 && !strided.empty() && xb.front() >= 12 && strided.back().front() >= 12 &&
                          (int)strided.back().size() >= 3 && nl - 12 >= 4;
        for (size_t si = 0; si < n_strided && layer >= 0; ++si) {
            XOp op = new_op(layer, nv);
            if (!strided_gx.empty()) op.pp = strided_gx[si];
            else QR_TRY(xop_set_geometry(op, strided[si], false));
            QR_TRY(xop_set_map(c, op, lin_identity(nt), nl, g));
            op.src_buf[0] = psi; op.src_buf[1] = lam; op.dst_buf[0] = psi; op.dst_buf[1] = lam;
            if (slice_pair && si + 1 == n_strided) op.nslices = sr->slices;
            QR_TRY(finish_op(op, ly));
            sr->ops.push_back(op);
        }
        // ---- exchange pass: the rank bits trade places with the local bits [sigma, sigma+g) ----
        {
            int held[8];
            layout_held(ly, held);
            Layout nw = ly;
            for (int b = 0; b < g; ++b) {
                nw.phi.col[sigma + b] = (u64)1 << held[b];          // the formerly rank-held logical bits, in pure form
                nw.phi.col[nl + b] = ly.phi.col[sigma + b];         // the local bits that move onto the rank bits
            }
            Lin F = owed ? owed_f : lin_identity(nt);
            bool fold = owed;
            if (!owed && !qaoa && nv == 1 && next_stacking >= 0 && peel_geometry) {
                // forward: would the next ladder read other shards in the new layout?  Then apply its offending CNOT here.
                const Lin Gn = lin_ladder_gather(nt, next_stacking);
                int pc, pt;
                if (!gather_is_local(nw, Gn) && layout_offending_cnot(nw, &pc, &pt)) {
                    const Lin Fc = lin_cnot(nt, pc, pt);
                    if (gather_is_local(nw, lin_mul(Fc, Gn))) { F = Fc; fold = true; pre.push_back(Fc); }
                }
            }
            Lin inv;
            if (!lin_inverse(ly.phi, &inv)) return fail(QR_ESTATE, "internal: singular layout");
            const Lin Mx = fold ? lin_mul(inv, lin_mul(F, nw.phi)) : lin_mul(inv, nw.phi);   // new[j] = old[F j]
            XOp op = new_op(layer, nv);
            QR_TRY(xop_set_geometry(op, xb, false));
            QR_TRY(xop_set_map(c, op, Mx, nl, g));
            if (!op.xmap) return fail(QR_ESTATE, "internal: exchange pass without peers");
            op.src_buf[0] = psi; op.src_buf[1] = lam;
            const int d0 = pick_free(psi, lam, -1);
            const int d1 = nv == 2 ? pick_free(psi, lam, d0) : -1;
            op.dst_buf[0] = d0; op.dst_buf[1] = d1;
            if (slice_pair) {   // the slice bits must not select the source shard
                for (int b = QR_HOLE_POS; b < QR_HOLE_POS + 3; ++b)
                    if (lin_apply(Mx, (u64)1 << b) >> nl) slice_pair = false;
                if (!slice_pair) sr->ops.back().nslices = 1;   // (the local pass was already given result slots per slice: harmless)
            }
            if (slice_pair) op.nslices = sr->slices;
            if (qaoa && nv == 2 && layer >= 0) {   // 2 Im<lambda|H|psi> and the un-phase after the un-mixing (qaoa.py:65-68), H in the NEW layout
                op.post_phase = 1; op.angle_post = -circ.gammas[layer]; op.ham_layout = swapped ? 0 : 1; op.lut_index = L + layer;
            }
            QR_TRY(finish_op(op, nw));
            swapped = !swapped;
            if (sr->lay[1].nl == 0 && swapped) sr->lay[1] = nw;
            if (slice_pair) sr->ops.back().pair_gen = op.gen;
            psi = d0; if (nv == 2) lam = d1;
            sr->ops.push_back(op);
            ly = nw;
        }
        return 0;
    };
    // ---- initial product state prod_q Ry(pi/4)|0>: popcount of the logical index ----
    {
        XOp op;
        op.kind = 0;
        op.dst_buf[0] = psi;
        op.init_pop = __builtin_popcount((unsigned)layout_shard_value(ly, me));
        op.init_plus = qaoa ? 1 : 0;
        sr->ops.push_back(op);
    }
    for (int i = 0; i < L; ++i) QR_TRY(add_layer(i, 1, 0, i + 1 < L ? 0 : -1));
    pre.clear();
    if (qaoa && !want_grad && swapped) {
        // forward-only QAOA runs leave the state in the natural layout (rank r = amplitudes r * 2^nl ...), where the
        // sharded sampler can scan it in index order: one more exchange pass without gates
        QR_TRY(add_layer(-1, 1, -1, -1));
    }
    // ---- observable in the current layout (QAOA: the co-state H psi from the layout's H table) ----
    {
        XOp op;
        op.kind = qaoa ? 3 : 2;
        op.terms = o->terms;
        bool needs_peer = false;
        QR_TRY(layout_remap_terms(ly, op.terms, &needs_peer));
        if (qaoa) { needs_peer = false; op.ham_layout = swapped ? 1 : 0; }
        op.obs_base = (u64)layout_shard_value(ly, me) << nl;
        for (int rho = 0; rho < G; ++rho) op.obs_peer_rank[layout_shard_value(ly, rho)] = rho;
        op.src_buf[0] = psi;
        op.nv = 1;
        if (want_grad) { lam = pick_free(psi, -1, -1); op.dst_buf[0] = lam; }
        op.remote = needs_peer && G > 1;
        op.remote_frac = 0.0;
        if (op.remote) { op.new_step = true; op.gen = ++gen; }
        if (want_grad && readable[lam] > waited) { op.wait_done = readable[lam]; waited = op.wait_done; }
        if (op.remote) readable[psi] = op.gen;
        sr->ops.push_back(op);
    }
    if (want_grad)
        for (int i = L - 1; i >= 0; --i) QR_TRY(add_layer(i, 2, i < L - 1 ? 1 : -1, -1));
    // ---- steps: a remote pass starts a step, and so does whatever follows it ----
    sr->step_first.clear();
    for (size_t k = 0; k < sr->ops.size(); ++k) {
        const bool after_remote = k > 0 && sr->ops[k - 1].remote;
        if (k == 0 || sr->ops[k].new_step || after_remote) sr->step_first.push_back((int)k);
    }
    sr->sweeps = max_sweeps;
    sr->n_results = res;
    sr->gen_base = gen;
    sr->final_psi = want_grad ? lam : psi;   // what State.vec holds afterwards: psi_final, or the back-propagated co-state
    return 0;
}

// ------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------
static int swap_signal(qr_ctx* c, int kind, long long value, cudaStream_t stream) {
    FlagPeers fp;
    memset(&fp, 0, sizeof(fp));
    const int G = 1 << c->g;
    for (int r = 0; r < G; ++r) fp.p[r] = c->peer_flags[r];
    QR_LAUNCH(k_flag_signal, 1, 32, 0, stream, fp, G, c->rank, kind, (unsigned long long)value);
    KERNEL_CHECK();
    return 0;
}
static int swap_wait(qr_ctx* c, int kind, long long value, cudaStream_t stream) {
    QR_LAUNCH(k_flag_wait, 1, 32, 0, stream, c->d_flags, 1 << c->g, kind, (unsigned long long)value);
    KERNEL_CHECK();
    return 0;
}

// one tile pass (or one slice of it) on `stream`; lane 0 / 1 = the reduction scratch and arrival counter it may use (a
// local pass and an exchange pass can be in flight together)
static int swap_launch_tile(qr_ctx* c, SwapRun* sr, const XOp& op, int slice, cudaStream_t stream, int max_sms, int lane) {
    const int nl = c->n;
    int hole = 0;
    while ((1 << hole) < op.nslices) ++hole;
    const u64 tile_or = op.nslices > 1 ? (u64)slice << QR_HOLE_POS : 0;
    const size_t lane_doubles = (size_t)c->sm_count * 16 * QR_SLOTS;
    const GateP* d_tab = (const GateP*)((char*)c->d_small + sr->tab_off) + op.gate_off;
    double* final_out = op.nv == 2 ? c->d_result + op.res_off + (size_t)slice * QR_SLOTS : nullptr;
    PassExtra ex;
    ex.hole = hole; ex.tile_or = tile_or; ex.max_sms = max_sms;
    ex.partials = c->d_scratch + (size_t)lane * lane_doubles;
    ex.counter = c->d_counter + lane;
    ex.stream = stream;
    if (!op.xmap) {
        LayerPlan lp;
        lp.n = nl; lp.k = 12; lp.R = 3; lp.npasses = 1;
        lp.pass[0] = op.pp;
        PassIO io = {c->buf[op.src_buf[0]], op.nv == 2 ? c->buf[op.src_buf[1]] : nullptr, c->buf[op.dst_buf[0]],
                     op.nv == 2 ? c->buf[op.dst_buf[1]] : nullptr};
        int units = 0;
        const bool ph = op.pre_phase || op.post_phase;
        if (ph) ex.hidx = c->d_hidx_ly[op.ham_layout];
        const double2* lut = (ph && c->ham_integer && c->opt_ham_lut && sr->lut_off) ?
            (const double2*)((char*)c->d_small + sr->lut_off) + (size_t)op.lut_index * c->ham_range : nullptr;
        return launch_pass(c, lp, 0, op.nv, io, d_tab, 0, -1, 1, (i64)c->N, 0, ph ? c->d_ham_ly[op.ham_layout] : nullptr, op.pre_phase,
                           op.angle_pre, op.post_phase, op.angle_post, &units, op.gather ? &op.spec : nullptr, final_out, lut, nullptr, &ex);
    }
    const PassPlan& pp = op.pp;
    const u64 lmask = ((u64)1 << nl) - 1;
    TilePass tp;
    memset(&tp, 0, sizeof(tp));
    tp.k = 12; tp.c = pp.c; tp.h = pp.h; tp.m1 = pp.m1; tp.h2 = pp.h2; tp.nrounds = pp.nrounds;
    tp.hole = hole; tp.tile_or = tile_or;
    tp.tiles_log2 = nl - 12 - hole;
    tp.num_tiles = (i64)1 << tp.tiles_log2;
    tp.state_stride = (i64)c->N;
    tp.src0 = c->buf[op.src_buf[0]]; tp.src1 = op.nv == 2 ? c->buf[op.src_buf[1]] : nullptr;
    tp.dst0 = c->buf[op.dst_buf[0]]; tp.dst1 = op.nv == 2 ? c->buf[op.dst_buf[1]] : nullptr;
    tp.gates = d_tab; tp.gate_stride = 0;
    tp.final_out = final_out;
    tp.done_counter = ex.counter;
    tp.partials = ex.partials;
    const bool ph = op.pre_phase || op.post_phase;
    if (ph) {
        tp.ham = c->d_ham_ly[op.ham_layout];
        tp.pre_phase = op.pre_phase; tp.post_phase = op.post_phase;
        tp.angle_pre = op.angle_pre; tp.angle_post = op.angle_post;
        if (c->ham_integer && c->opt_ham_lut && sr->lut_off) {
            tp.hidx = c->d_hidx_ly[op.ham_layout];
            tp.lut = (const double2*)((char*)c->d_small + sr->lut_off) + (size_t)op.lut_index * c->ham_range;
            tp.lut_size = c->ham_range;
            tp.hmin = c->ham_min;
        }
    }
    Tile12X x;
    memset(&x, 0, sizeof(x));
    x.ngroups = pp.ngroups;
    x.last_group = pp.ngroups == 1 ? 9 : 6;
    const Geo12 geo0 = {pp.c, pp.h, pp.m1, pp.h2, 12, 0};
    for (int r = 0; r < 8; ++r) {
        x.droff_first[r] = geo12_local(geo0, (u64)r << 9);
        x.roff_first[r] = op.roff[r];
        x.roff_last[r] = geo12_local(geo0, (u64)r << x.last_group);
    }
    TileXMap xm;
    memset(&xm, 0, sizeof(xm));
    for (int v = 0; v < op.nv; ++v)
        for (int sl = 0; sl < 4; ++sl)
            for (int r = 0; r < 8; ++r) xm.src[v][sl][r] = c->peer[op.src_rank[sl][r]][op.src_buf[v]];
    memcpy(xm.lcol, op.lcol, sizeof(xm.lcol));
    xm.thr_bit[0] = op.thr_bit[0]; xm.thr_bit[1] = op.thr_bit[1];
    xm.sel_bit[0] = xm.sel_bit[1] = -1;
    // tile-index images for this launch's enumeration (the slice bits are fixed: their image joins the constant)
    const Geo12 geo = {pp.c, pp.h, pp.m1, pp.h2, 12, hole};
    for (int j = 0; j < tp.tiles_log2 && j < 24; ++j) {
        const u64 d = geo12_tile(geo, (u64)1 << j);
        xm.tcol[j] = lin_apply(op.M, d) & lmask;
        for (int k = 0; k < 2; ++k)
            if (op.sel_pos[k] >= 0 && d == ((u64)1 << op.sel_pos[k])) xm.sel_bit[k] = j;
    }
    for (int k = 0; k < 2; ++k)
        if (op.sel_pos[k] >= 0 && xm.sel_bit[k] < 0) return fail(QR_ESTATE, "internal: selector bit lost in a sliced launch");
    const u64 simg = lin_apply(op.M, tile_or);
    if (simg >> nl) return fail(QR_ESTATE, "internal: the slice bits select the source shard");
    xm.src_const = op.src_const ^ (simg & lmask);
    xm.local_only = op.remote ? 0 : 1;
    // L2 prefetch of the next tile: local sources only (a prefetch of peer memory would warm the PEER's L2)
    tp.prefetch = (!op.remote && op.nv == 2) ? (int)(c->opt_prefetch & 3) : 0;
    const int sms = max_sms > 0 ? std::min(max_sms, c->sm_count) : c->sm_count;
    const i64 grid = std::min<i64>(tp.num_tiles, (i64)sms * (op.nv == 1 ? 2 : 1));
    const size_t smem = (size_t)op.nv * (sizeof(double2) << 12);
    typedef void (*xfn)(const TilePass, const Tile12X, const TileXMap);
    const xfn fn = op.nv == 1 ? (ph ? k_tile12_x<1, true> : k_tile12_x<1, false>) : (ph ? k_tile12_x<2, true> : k_tile12_x<2, false>);
    QR_TRY(ensure_smem_attr(c, (const void*)fn, 27 + (op.nv - 1) * 2 + (ph ? 1 : 0)));
    QR_LAUNCH(fn, (unsigned)grid, 512, smem, stream, tp, x, xm);
    KERNEL_CHECK();
    c->tables_fresh = false;
    c->perf.kernel_launches++;
    return 0;
}

static int swap_launch_op(qr_ctx* c, SwapRun* sr, const XOp& op, int op_index) {
    const int nl = c->n, G = 1 << c->g;
    const bool async = !sr->lockstep;
    const bool sliced = op.kind == 1 && op.nslices > 1;
    // a sliced exchange pass runs on the second stream (asynchronous mode), behind everything issued so far
    cudaStream_t st = c->stream;
    if (async && sliced) {
        if (!c->stream2) CUDA_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) if (!c->ev_pair[i]) CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pair[i], cudaEventDisableTiming));
    }
    if (async && sliced && op.remote) {
        st = c->stream2;
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_pair[0], 0));   // recorded before the sliced local pass of this layer
    }
    if (async && op.wait_done > 0) QR_TRY(swap_wait(c, QR_FLAG_DONE, op.wait_done, st));
    if (async && op.remote && !sliced) {
        QR_TRY(swap_signal(c, QR_FLAG_READY, op.gen, st));
        QR_TRY(swap_wait(c, QR_FLAG_READY, op.gen, st));
    }
    if (sr->ev.size() < 2 * sr->ops.size()) {
        const size_t old = sr->ev.size();
        sr->ev.resize(2 * sr->ops.size(), nullptr);
        for (size_t i = old; i < sr->ev.size(); ++i) cudaEventCreate(&sr->ev[i]);
    }
    CUDA_TRY(cudaEventRecord(sr->ev[2 * op_index], st));
    if (op.kind == 0 && op.init_plus) {
        QR_LAUNCH(k_init_basis, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[op.dst_buf[0]], c->N, 1, std::pow(2.0, -0.5 * c->n_total));
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    } else if (op.kind == 3) {
        const int ogrid = grid_for(c, c->N);
        QR_LAUNCH(k_ham_costate, ogrid, QR_BLOCK, 0, c->stream, (const double2*)c->buf[op.src_buf[0]],
                  op.dst_buf[0] >= 0 ? c->buf[op.dst_buf[0]] : (double2*)nullptr, (const double*)c->d_ham_ly[op.ham_layout], c->N, c->d_scratch);
        KERNEL_CHECK();
        QR_LAUNCH(k_reduce_partials, 1, QR_BLOCK, 0, c->stream, (const double*)c->d_scratch, ogrid, 1, c->d_result);
        KERNEL_CHECK();
        c->perf.kernel_launches += 2;
    } else if (op.kind == 0) {
        const int nt = c->n_total;
        double table[48];
        const double cs = std::cos(M_PI / 8.0), sn = std::sin(M_PI / 8.0);
        for (int w = 0; w <= nt; ++w) { double v = 1.0; for (int q = 0; q < nt; ++q) v *= (q < nt - w) ? cs : sn; table[w] = v; }
        const size_t off = c->small_cap - 512;
        QR_TRY(upload_small(c, off, table, sizeof(double) * (nt + 1), c->pin_cap - 512));
        QR_LAUNCH(k_init_product, grid_for(c, c->N), QR_BLOCK, 0, c->stream, c->buf[op.dst_buf[0]], c->N,
                  (const double*)((char*)c->d_small + off) + op.init_pop, c->N - 1);
        KERNEL_CHECK();
        c->perf.kernel_launches++;
    } else if (op.kind == 2) {
        PeerTable peers;
        memset(&peers, 0, sizeof(peers));
        for (int u = 0; u < G; ++u) peers.p[u] = c->peer[op.obs_peer_rank[u]][op.src_buf[0]];
        const int ogrid = grid_for(c, c->N);
        QR_LAUNCH(k_apply_obs, ogrid, QR_BLOCK, 0, c->stream, (const double2*)c->buf[op.src_buf[0]],
                  op.dst_buf[0] >= 0 ? c->buf[op.dst_buf[0]] : (double2*)nullptr, c->N, (const ObsTerm*)c->d_small, (int)op.terms.size(),
                  c->d_scratch, op.obs_base, nl, peers);
        KERNEL_CHECK();
        QR_LAUNCH(k_reduce_partials, 1, QR_BLOCK, 0, c->stream, (const double*)c->d_scratch, ogrid, 1, c->d_result);
        KERNEL_CHECK();
        c->perf.kernel_launches += 2;
    } else if (!sliced) {
        QR_TRY(swap_launch_tile(c, sr, op, 0, c->stream, 0, 0));
    } else if (!async) {
        // lockstep: the slices one after another on the main stream (same kernels, same enumeration as the overlapped run)
        for (int sl = 0; sl < op.nslices; ++sl) QR_TRY(swap_launch_tile(c, sr, op, sl, c->stream, 0, 0));
    } else if (!op.remote) {
        // sliced local pass: raise READY of the exchange pass's generations slice by slice; leave SMs to the exchange pass
        CUDA_TRY(cudaEventRecord(c->ev_pair[0], c->stream));
        const int sms = std::max(8, c->sm_count - (int)std::min<long long>(c->opt_shard_xsms, c->sm_count - 8));
        for (int sl = 0; sl < op.nslices; ++sl) {
            QR_TRY(swap_launch_tile(c, sr, op, sl, c->stream, sl == 0 ? 0 : sms, 0));   // slice 0 has the GPU to itself
            QR_TRY(swap_signal(c, QR_FLAG_READY, op.pair_gen + sl, c->stream));
        }
    } else {
        // sliced exchange pass on the second stream: slice s waits for READY(s) of every rank (own local pass included)
        const int xs = (int)std::min<long long>(c->opt_shard_xsms, c->sm_count);
        for (int sl = 0; sl < op.nslices; ++sl) {
            QR_TRY(swap_wait(c, QR_FLAG_READY, op.gen + sl, st));
            QR_TRY(swap_launch_tile(c, sr, op, sl, st, sl + 1 == op.nslices ? 0 : xs, 1));   // the last slice runs alone
            QR_TRY(swap_signal(c, QR_FLAG_DONE, op.gen + sl, st));
        }
    }
    if (op.kind == 1) c->perf.link_bytes += op.remote_frac * (double)op.nv * 16.0 * (double)c->N;   // bytes this rank reads over NVLink
    CUDA_TRY(cudaEventRecord(sr->ev[2 * op_index + 1], st));
    if (async && sliced && op.remote) {   // the main stream continues behind the last slice
        CUDA_TRY(cudaEventRecord(c->ev_pair[1], st));
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_pair[1], 0));
    } else if (async && op.remote) QR_TRY(swap_signal(c, QR_FLAG_DONE, op.gen, st));
    return 0;
}

// per-kind device times of the finished run (the stream has been synchronised): local passes forward / backward, average
// exchange pass forward / backward, observable; total = first op to last op (the difference is waiting for peers)
static void swap_collect_times(qr_ctx* c, SwapRun* sr) {
    if (sr->ev.size() < 2 * sr->ops.size()) return;
    double loc[3] = {0, 0, 0}, xch[3] = {0, 0, 0}, obs = 0;
    int nx[3] = {0, 0, 0};
    for (size_t k = 0; k < sr->ops.size(); ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sr->ev[2 * k], sr->ev[2 * k + 1]) != cudaSuccess) { cudaGetLastError(); return; }
        const XOp& op = sr->ops[k];
        if (op.kind == 2) obs += ms;
        else if (op.kind == 1 && op.remote) { xch[op.nv] += ms; nx[op.nv]++; }
        else loc[op.kind == 0 ? 1 : op.nv] += ms;
    }
    float tot = 0.f;
    cudaEventElapsedTime(&tot, sr->ev[0], sr->ev[2 * sr->ops.size() - 1]);
    c->perf.ms_total = tot;
    c->perf.ms_forward = loc[1];
    c->perf.ms_backward = loc[2];
    c->perf.ms_observable = obs;
    c->perf.fwd_pass_ms_avg = nx[1] ? xch[1] / nx[1] : 0.0;
    c->perf.bwd_pass_ms_avg = nx[2] ? xch[2] / nx[2] : 0.0;
    c->perf.fwd_pass_bytes = (double)nx[1];   // number of exchange passes (forward / backward)
    c->perf.bwd_pass_bytes = (double)nx[2];
}
