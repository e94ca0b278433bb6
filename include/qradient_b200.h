/*
 * qradient_b200.h -- C ABI of libqradient_b200.so (hand-written sm_100a CUDA kernels).
 *
 * The reference (frederikwilde/qradient) is pure Python with no FFI: its drop-in boundary is
 * the Python class surface `qradient.physical_components.{State,Gates,Observable}` and
 * `qradient.circuit_logic.{McClean,Qaoa}`.  The Python package `qradient_b200` re-creates that
 * surface and binds these entry points through ctypes; each entry point
 * cites the reference code it replaces (paths relative to /root/reference/qradient).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; the message is available from
 *     qr_last_error() (thread-local).  The Python layer raises ValueError for QR_EINVAL and
 *     RuntimeError otherwise.
 *   - qr_ctx owns one CUDA stream and all device buffers of one state register (the state
 *     vector psi, the co-state lambda of the adjoint sweep, ping-pong scratch, the diagonal
 *     Hamiltonian table, reduction scratch).  A context must not be used from two threads at
 *     once (same rule as the reference objects).
 *   - host pointers are caller-owned and only read/written during the call.  Calls that take or return host data are
 *     synchronous at return; calls that only launch kernels on the context's stream (single gates, ladders, diagonal
 *     phases, snapshots) are stream ordered and return at once -- the next call that returns data synchronises
 *     (qr_sync forces it).
 *   - amplitudes are complex128, interleaved (re, im); amplitude index j = sum_q b_q 2^(n-1-q)
 *     (qubit 0 is the most significant bit; physical_components/state.py:84-88,163).
 *   - there is NO CPU fallback: every compute entry point launches CUDA kernels.
 */
#ifndef QRADIENT_B200_H
#define QRADIENT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qr_ctx qr_ctx;
typedef struct qr_obs qr_obs;

enum { QR_OK = 0, QR_EINVAL = 1, QR_ECUDA = 2, QR_ENOMEM = 3, QR_ESTATE = 4 };

/* observable term kinds (physical_components/observable.py:45-73) */
enum { QR_TERM_X = 0, QR_TERM_Y = 1, QR_TERM_Z = 2, QR_TERM_ZZ = 3 };

/* qr_set_option keys (numbering is stable; retired round-1 experiment keys 7-10, 14-18, 22-25, 27 are rejected) */
enum {
    QR_OPT_FUSION = 0,        /* 1 (default): fused tile passes; 0: one kernel per gate          */
    QR_OPT_TILE_BITS = 1,     /* log2 amplitudes per shared-memory tile; 0 (default) = auto: 11 or 12 */
    QR_OPT_PREFETCH = 2,      /* L2 prefetch distance in tiles: bits 0-1 backward (default 1), bits 2-3 forward (default 0), bit 4: contiguous passes only */
    QR_OPT_CTAS_PER_SM_FWD = 3,
    QR_OPT_CTAS_PER_SM_BWD = 4,
    QR_OPT_FINAL_LADDER = 5,  /* 1 (default): leave state.vec exactly as mc_clean.py:77 does     */
    QR_OPT_HAM_LUT = 6,       /* 1 (default): integer-valued diagonal Hamiltonians use a phase LUT */
    QR_OPT_TILE_BITS_STRIDED = 11, /* tile bits of the strided (non-first) passes; 0 = same as first */
    QR_OPT_MIN_ROW_BITS = 12, /* log2 of the minimum contiguous run (amplitudes) in strided passes */
    QR_OPT_BATCH_CHUNK_MB = 13, /* batched circuits: MiB of state per buffer processed per chunk (0 = 512) */
    QR_OPT_STAGED = 19,       /* k_tile12: next tile staged in shared memory by asynchronous copies: bit0 backward, bit1 forward, bit2 (default) auto */
    QR_OPT_STAGED_MIN_BIT = 20, /* auto mode: strided backward passes whose lowest gate bit is >= this (default 21) are staged */
    QR_OPT_PDL = 26,          /* k_tile12 passes use programmatic dependent launch (griddepcontrol): 0 off, 1 (default) auto: short passes (n <= 22), 2 always */
    QR_OPT_DEFER_REDUCE = 28, /* 1 (default): single circuits add the per-CTA gradient partials of all backward passes in ONE launch after the sweep; 0: last-CTA reduction fused into every pass */
    QR_OPT_SHARD_MODE = 30,   /* sharded registers: 0 (default) auto, 1 "peer" engine (peer loads + peer stores per global step; any size), 2 "swap" engine (one NVLink crossing per exchange fused into a tile pass; >= 12 + log2(ranks) local qubits) */
    QR_OPT_SHARD_LOCKSTEP = 31, /* swap engine: 1 (default) the caller runs qr_shard_step in lockstep over the ranks; 0: all steps are enqueued at once and device-side flags order the ranks */
    QR_OPT_SHARD_SLICES = 32, /* swap engine: the last local pass and the exchange pass of a layer are issued in 1 (default), 2, 4 or 8 slices of the index bits 9..11; asynchronous mode runs the exchange pass of a slice on a second stream while the local pass works on the next slice */
    QR_OPT_LOOP_GRAPH = 35,   /* 1 (default): the device optimiser loops (qr_mcclean_optimize, qr_qaoa_optimize) replay a CUDA graph captured from their second step; 0: every step is launched kernel by kernel */
    QR_OPT_AXIS_PLAN = 34,    /* single McClean circuits, bit mask (default 15, applied from 20 qubits on unless the option is set explicitly): bit 0 plan the strided passes of every layer from its axes -- an Rz needs no tile bit, so the tiles keep wider rows and fewer exchange rounds; bit 1 the contiguous pass trades Rz-only bits for high X / Y bits where that saves an exchange round; bit 2 passes without an exchange run the two-round program; bit 3 the even / odd warps of a two-round backward pass synchronise separately; 0: one static plan */
    QR_OPT_SHARD_XSMS = 33,   /* asynchronous mode: SMs given to the exchange pass of a slice (default 60); the concurrent local pass uses the others */
    QR_OPT_SHARD_ZSKIP = 29   /* 1 (default): sharded states apply an Rz on a global qubit as a per-subgroup phase without the NVLink exchange (only X / Y rotations are exchanged) */
};

typedef struct qr_perf {
    double ms_total;          /* device time of the last fused call (CUDA events on the ctx stream) */
    double ms_forward;
    double ms_observable;
    double ms_backward;
    double algorithmic_bytes; /* bytes the schedule moves by construction (B_sched)  */
    double bwd_pass_ms_avg;   /* average duration of one backward tile-pass launch                  */
    double bwd_pass_bytes;    /* algorithmic bytes of one backward tile-pass launch                 */
    double fwd_pass_ms_avg;
    double fwd_pass_bytes;
    long long kernel_launches;/* kernels launched by the last fused call                            */
    int passes_per_layer;
    int tile_bits;
    double link_bytes;        /* sharded states: bytes this rank moved over NVLink (peer loads + peer stores) since the run began */
} qr_perf;

const char* qr_last_error(void);
int qr_version(void);
int qr_device_count(int* out);

/* ---- context ------------------------------------------------------------------------- */
/* replaces State.__init__ (state.py:34-37): allocates psi on `device` and resets it to |0..0> */
int qr_ctx_create(int n_qubits, int device, qr_ctx** out);
int qr_ctx_destroy(qr_ctx* ctx);
int qr_set_option(qr_ctx* ctx, int key, long long value);
int qr_get_option(qr_ctx* ctx, int key, long long* value);
int qr_perf_last(qr_ctx* ctx, qr_perf* out);
int qr_sync(qr_ctx* ctx);

/* ---- state vector (physical_components/state.py) -------------------------------------- */
/* State.reset (state.py:61-71): which = 0 -> |0..0>, 1 -> |+..+> */
int qr_state_init(qr_ctx* ctx, int which);
/* State.vec setter / getter: n_amps must equal 2^n */
int qr_state_upload(qr_ctx* ctx, const double* re_im, size_t n_amps);
int qr_state_download(qr_ctx* ctx, double* re_im, size_t n_amps);
/* device-side copies of the state vector; replace the reference's `state_history[i] = state.vec`
 * snapshots (mc_clean.py:132, qaoa.py:118-120) without a host round trip */
int qr_state_save(qr_ctx* ctx, int slot);
int qr_state_load(qr_ctx* ctx, int slot);
int qr_state_free_snapshots(qr_ctx* ctx);
/* device address of the current state vector (for torch / NCCL plumbing; complex128[2^n]) */
int qr_state_device_ptr(qr_ctx* ctx, void** out);
/* xrot / yrot / zrot (state.py:90-92,142-144,168-170): axis 0,1,2 = X,Y,Z; exp(-i angle P/2) */
int qr_apply_rot(qr_ctx* ctx, int axis, double angle, int qubit);
/* dxrot / dyrot / dzrot (state.py:94-97,146-149,172-175): derivative of the rotation */
int qr_apply_drot(qr_ctx* ctx, int axis, double angle, int qubit);
/* cnot (state.py:198-199,336-356): control, target */
int qr_apply_cnot(qr_ctx* ctx, int control, int target);
/* cnot_ladder (state.py:209-251): stacking 0 or 1 (its inverse); periodic needs even n */
int qr_apply_cnot_ladder(qr_ctx* ctx, int stacking, int periodic);
/* xrot_all / x_summed (state.py:107-123): vec = sum_q 1/2 (-i X_q) vec */
int qr_apply_x_summed(qr_ctx* ctx);
/* norm_error (state.py:331-332) needs ||vec||^2 */
int qr_norm2(qr_ctx* ctx, double* out);

/* ---- observable (physical_components/observable.py) ------------------------------------ */
/* Observable.load_matrix (observable.py:34-79) in term-list form: kind[k], qubit i[k],
 * second qubit j[k] (ZZ only, i<j), weight w[k]; order = projector order (observable.py:90-101) */
int qr_obs_create(int n_qubits, int n_terms, const int32_t* kind, const int32_t* qi, const int32_t* qj,
                  const double* w, qr_obs** out);
int qr_obs_destroy(qr_obs* obs);
/* State.multiply_matrix(observable.matrix) (mc_clean.py:65): vec = O vec */
int qr_apply_observable(qr_ctx* ctx, const qr_obs* obs);
/* ParametrizedCircuit.expec_val (base.py:17-20): Re <vec|O|vec> */
int qr_expec_val(qr_ctx* ctx, const qr_obs* obs, double* out);
/* per-term <P_k> in projector order; prob_k = (1 + <P_k>)/2 is what base.py:28 computes */
int qr_term_expecs(qr_ctx* ctx, const qr_obs* obs, double* out_terms);

/* ---- diagonal ("classical") Hamiltonian (state.py:261-321) ------------------------------ */
/* load_classical_ham / Gates.add_classical_ham: builds H[j] on the device; z and zz terms only */
int qr_ham_load(qr_ctx* ctx, const qr_obs* obs);
int qr_ham_download(qr_ctx* ctx, double* out, size_t n_amps);
/* rot_classical_ham / exp_ham_classical (state.py:299-301): vec *= exp(-i angle H) */
int qr_apply_exp_ham(qr_ctx* ctx, double angle);
/* rot_classical_ham_component (state.py:309-311): same with the k-th term of obs only */
int qr_apply_exp_ham_component(qr_ctx* ctx, const qr_obs* obs, int k, double angle);
/* classical_ham() / ham_classical (state.py:319-321): mode 0: vec *= -i H ; mode 1: vec *= H
 * (the latter is `state.vec *= gates.classical_ham`, qaoa.py:56) */
int qr_apply_ham(qr_ctx* ctx, int mode);

/* ---- fused circuit paths (circuit_logic/mc_clean.py, qaoa.py) --------------------------- */
/* McClean.run_expec_val (mc_clean.py:27-45).  axes/angles: [L*n] row-major.  If
 * use_current_state != 0 the current vector is the ini_state (mc_clean.py:32), else reset. */
int qr_mcclean_expec(qr_ctx* ctx, int n_layers, const int32_t* axes, const double* angles,
                     const qr_obs* obs, int use_current_state, double* e_out);
/* McClean.grad_run (mc_clean.py:47-78): E and dE/d angles[L*n]; afterwards the state vector holds
 * the back-propagated co-state exactly as the reference leaves it. */
int qr_mcclean_grad(qr_ctx* ctx, int n_layers, const int32_t* axes, const double* angles,
                    const qr_obs* obs, int use_current_state, double* e_out, double* grad_out);

/* Layered circuit: |0..0> (or the current state), then n_layers x { CNOT ladder iff ladder_before[i]; one Pauli rotation
 * per qubit, axes/angles [n_layers * n] }; E = <O>, grad[i*n+q] = dE/d angles[i,q] (grad may be NULL: forward only,
 * state.vec = psi_final).  Engine of MeynardClassifier.run / grad_run (tutorials/meynard-classifier.ipynb cells 3, 11,
 * 14; the class has no source in the reference snapshot): a classifier layer [ladder] Rx Ry Rz is three
 * such sub-layers. */
int qr_layered_grad(qr_ctx* ctx, int n_layers, const int32_t* axes, const double* angles, const unsigned char* ladder_before,
                    int use_current_state, const qr_obs* obs, double* e_out, double* grad_out);
/* batched extension (timing-test.ipynb cell 6 loop): B independent parameter sets on one device.
 * axes/angles: [B*L*n]; e_out[B]; grad_out[B*L*n].  ctx must have been created with n qubits. */
int qr_mcclean_grad_batch(qr_ctx* ctx, int batch, int n_layers, const int32_t* axes, const double* angles,
                          const qr_obs* obs, double* e_out, double* grad_out);
/* Qaoa.run_expec_val (qaoa.py:23-38); requires qr_ham_load(ctx, obs) first */
int qr_qaoa_expec(qr_ctx* ctx, int n_layers, const double* betas, const double* gammas,
                  int use_current_state, double* e_out);
/* Qaoa.grad_run (qaoa.py:40-70): grad_out[p*2], column 0 = d/d beta, column 1 = d/d gamma */
int qr_qaoa_grad(qr_ctx* ctx, int n_layers, const double* betas, const double* gammas,
                 int use_current_state, double* e_out, double* grad_out);

/* ---- optimiser loop on the device (optimization.py:41-91 McCleanOpt.step, :131-194 update rules) ---------- */
/* `steps` x (grad_run, update) without host round trips.  rule: 0 Adam, 1 GradientDescent (constant step),
 * 2 RateDecayOnPlateau.  hyper[8] = {step_size, beta1, beta2, eps, plateau_length, decay_rate, cost, plateau_counter}
 * (step_size, cost and plateau_counter are updated); m / v: Adam moments [L*n] in/out (null otherwise);
 * angles [L*n] in/out; cost_history [steps]; param_history [steps][L*n] (parameters after each step) or null. */
int qr_mcclean_optimize(qr_ctx* ctx, int n_layers, const int32_t* axes, double* angles, const qr_obs* obs, int rule,
                        double* hyper, int* iter_inout, double* m_inout, double* v_inout, int steps,
                        double* cost_history, double* param_history);
/* QaoaOpt.step x steps (optimization.py:113-129): params [n_layers][2] = rows (beta_i, gamma_i), in/out; m / v [n_layers*2];
 * param_history [steps][n_layers*2].  Needs an integer-valued Hamiltonian (MaxCut: the phase look-up tables are rebuilt on
 * the device from the updated gammas); QR_EINVAL otherwise -- the caller keeps the host loop. */
int qr_qaoa_optimize(qr_ctx* ctx, int n_layers, double* params, int rule, double* hyper, int* iter_inout, double* m_inout,
                     double* v_inout, int steps, double* cost_history, double* param_history);

/* ---- finite-shot sampling (qaoa.py:196-198, mc_clean.py:259-261) ------------------------ */
/* index = first k with cumsum(|vec|^2)[k] >= u  (scipy rv_discrete inverse CDF); uniforms are
 * supplied by the caller from the numpy stream so the draw order matches the reference. */
int qr_sample_bitstrings(qr_ctx* ctx, int n_shots, const double* uniforms, int64_t* out_idx);
/* vec[k] <- vec[perm[k]] with a permutation kept on the device: measurement in the order of the sorted
 * eigenvalues of a diagonal observable (McClean.sample_grad_dense, mc_clean.py:207-268, matrix-free) */
int qr_perm_load(qr_ctx* ctx, const int64_t* perm, size_t n_amps);
int qr_state_permute(qr_ctx* ctx);
/* dense basis change psi <- M psi (M: 2^n x 2^n, row major, interleaved re/im; n <= 12): the eigenbasis measurement of
 * observables with x / y terms, replaces eigenvectors.transpose().conj() / lhs.dot(vec) of mc_clean.py:221-224, 255-256 */
int qr_dense_load(qr_ctx* ctx, const double* m_re_im, size_t dim);
int qr_state_apply_dense(qr_ctx* ctx);
/* mean of H[idx] over the sampled indices, computed on the device from the loaded H table */
int qr_ham_gather(qr_ctx* ctx, int n, const int64_t* idx, double* out_vals);

/* ---- sharded state vector (no reference counterpart) ------------------------ */
/* One context per rank holds 2^(n_total - log2_world) amplitudes: the top log2_world qubits are the
 * rank bits.  All four ping-pong buffers are allocated so their IPC handles can be exchanged once. */
int qr_shard_create(int n_total, int log2_world, int rank, int device, qr_ctx** out);
/* buf = 0..3: the ping-pong state buffers; buf = 4: the rank's array of device-side ordering flags (swap engine) */
int qr_shard_ipc_handle(qr_ctx* ctx, int buf, void* handle64);                    /* cudaIpcGetMemHandle  */
int qr_shard_ipc_open(qr_ctx* ctx, int peer_rank, int buf, const void* handle64); /* cudaIpcOpenMemHandle */
int qr_shard_set_peer_ptr(qr_ctx* ctx, int peer_rank, int buf, void* ptr, int peer_device); /* same process */
int qr_shard_buffer_ptr(qr_ctx* ctx, int buf, void** out);
int qr_shard_info(qr_ctx* ctx, int* n_total, int* log2_world, int* rank, int* logical_shard);
/* McClean forward (+ adjoint backward) on the sharded register, as a sequence of steps; the caller
 * must run a cross-rank barrier after every qr_shard_step (each step is stream-synchronised).
 * Local qubits: fused tile passes; CNOT ladder: shard relabelling, no data movement; global qubits:
 * one kernel per layer that exchanges and rotates over peer memory (NVLink P2P). */
int qr_shard_mcclean_begin(qr_ctx* ctx, int n_layers, const int32_t* axes, const double* angles, const qr_obs* obs,
                           int want_grad, int* n_steps);
int qr_shard_step(qr_ctx* ctx, int step);
/* Qaoa.run_expec_val / grad_run (qaoa.py:23-70) on the sharded register (swap engine; z / zz observable).  Same step /
 * finish protocol; qr_shard_mcclean_finish then returns this rank's partial E and grad[2 p] (column 0 = d/d beta, column 1 =
 * d/d gamma).  A forward-only run leaves the state in the natural layout (rank r holds amplitudes r * 2^nl ...), so that
 * qr_norm2 / qr_sample_bitstrings on every shard give the sharded sampler its local totals and indices (qaoa.py:196-198). */
int qr_shard_qaoa_begin(qr_ctx* ctx, int n_layers, const double* betas, const double* gammas, const qr_obs* obs,
                        int want_grad, int* n_steps);
/* this rank's partial sums of E and of dE/d angles[L * n_total]; sum over ranks (allreduce) */
int qr_shard_mcclean_finish(qr_ctx* ctx, double* e_partial, double* grad_partial);

#ifdef __cplusplus
}
#endif
#endif /* QRADIENT_B200_H */
