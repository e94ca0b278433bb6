"""CPU oracle for the qradient hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  Nothing under ``qradient_b200/``
imports it; the product path fails loudly when the CUDA library is missing.

What it is: an independent, matrix-free numpy restatement of the reference's
state-vector simulator and gradient algorithms.  Every function cites the reference
file:line it follows (paths relative to ``/root/reference/qradient``).  The reference
builds 2^n x 2^n scipy.sparse generators and multiplies them into the vector; here the
same linear maps are applied through reshaped views, so the oracle needs only O(2^n)
memory per vector and runs at n ~ 24 on a laptop.  The *algorithms* (history-based
gradient with derivative gates, per-term Bernoulli sampling, inverse-CDF bitstring
sampling) follow the reference step by step and deliberately do NOT use the two-vector
adjoint recurrence the CUDA path uses, so agreement between the two is a real check.

Parity pinning: the reference package at HEAD does not import (``Gates`` is missing,
SURVEY.md section 0.2) and its tests are stubs, so there are no reference-held golden
vectors.  The oracle is pinned against outputs of the reference's own
``circuit_logic/*.py`` executed verbatim in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle.py``
checks the oracle against those files.

Conventions (physical_components/state.py:84-88,163): qubit q is the q-th Kronecker
factor from the left, i.e. index bit n-1-q; amplitude index j = sum_q b_q 2^(n-1-q).
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "OracleState", "OracleObservable", "mcclean_run_expec_val", "mcclean_grad_run",
    "qaoa_run_expec_val", "qaoa_grad_run", "sample_expec_val", "sample_bitstrings",
    "classical_ham_vector", "ladder_scatter_map", "ladder_permutation", "maxcut_observable",
    "mcclean_sample_grad", "finite_difference_grad",
]


# --------------------------------------------------------------------------------------
# observable dictionary -> term lists     (physical_components/observable.py:34-79)
# --------------------------------------------------------------------------------------
class OracleObservable:
    """Term-list form of the reference's observable dictionary.

    observable.py:45-54 adds ``w * (1 x P_q x 1)`` for every non-None entry of the
    'x', 'y', 'z' arrays; observable.py:56-73 adds ``w * Z_i Z_j`` for every non-None
    entry of the strictly upper triangle of 'zz' and raises ValueError if the diagonal or
    lower triangle holds anything but None.  Projector order (observable.py:90-101):
    all 'x' by qubit, then 'y', then 'z', then 'zz' row-major.
    """

    def __init__(self, qubit_number: int, observable: dict):
        self.qnum = int(qubit_number)
        self.terms = []  # (kind, i, j, weight) in projector order
        for key in ("x", "y", "z"):
            if key in observable and observable[key] is not None:
                arr = observable[key]
                if len(arr) != self.qnum:
                    raise ValueError("Inconsistent shapes in observable dictionary.")
                for q, w in enumerate(arr):
                    if w is not None:
                        self.terms.append((key, q, -1, float(w)))
        if "zz" in observable and observable["zz"] is not None:
            zz = np.asarray(observable["zz"], dtype=object)
            if zz.shape != (self.qnum, self.qnum):
                raise ValueError("Inconsistent shapes in observable dictionary.")
            for i in range(self.qnum):
                for j in range(i + 1):
                    if zz[i, j] is not None:
                        raise ValueError("zz of observable should be an upper triangular matrix.")
                for j in range(i + 1, self.qnum):
                    if zz[i, j] is not None:
                        self.terms.append(("zz", i, j, float(zz[i, j])))

    @property
    def weights(self):
        return np.array([t[3] for t in self.terms], dtype=float)

    def scale(self) -> float:
        """Operator-norm bound sum_k |w_k| used by the parity criterion."""
        return float(np.abs(self.weights).sum()) if self.terms else 1.0


def maxcut_observable(vertex_num: int, edge_set) -> dict:
    """optimization_problems.py:33-37: +1 * Z_i Z_j for every edge, None elsewhere."""
    zz = np.full([vertex_num, vertex_num], None)
    for a, b in edge_set:
        zz[a, b] = 1.0
    return {"zz": zz}


# --------------------------------------------------------------------------------------
# state vector + gates                     (physical_components/state.py)
# --------------------------------------------------------------------------------------
class OracleState:
    def __init__(self, qubit_number: int, ini: str = "0"):
        self.qnum = int(qubit_number)
        self.ini = ini
        self.reset()

    # state.py:61-71
    def reset(self, ini=None):
        if ini is not None:
            self.ini = ini
        n = self.qnum
        if self.ini == "0":
            self.vec = np.zeros(2 ** n, dtype=complex)
            self.vec[0] = 1.0
        elif self.ini == "+":
            self.vec = 2.0 ** (-0.5 * n) * np.ones(2 ** n, dtype=complex)
        else:
            raise ValueError("Invalid initialization format {}.".format(self.ini))

    def _pair_view(self, q):
        n = self.qnum
        if not (0 <= q < n):
            raise ValueError("Invalid qubit index {} for {} qubits.".format(q, n))
        v = self.vec.reshape(2 ** q, 2, 2 ** (n - 1 - q))
        return v[:, 0, :], v[:, 1, :]

    def _gen(self, axis, q):
        """Return G v for the generator G = -i P_q (state.py:84-88,136-140)."""
        a, b = self._pair_view(q)
        out = np.empty_like(self.vec)
        o = out.reshape(2 ** q, 2, 2 ** (self.qnum - 1 - q))
        if axis == 0:      # -iX = [[0,-i],[-i,0]]
            o[:, 0, :] = -1j * b
            o[:, 1, :] = -1j * a
        elif axis == 1:    # -iY = [[0,-1],[1,0]]
            o[:, 0, :] = -b
            o[:, 1, :] = a
        else:
            raise ValueError("Invalid axis {}".format(axis))
        return out

    # state.py:90-92 / 142-144 : vec = sin(a/2) G vec + cos(a/2) vec
    def xrot(self, angle, q):
        self.vec = np.sin(0.5 * angle) * self._gen(0, q) + np.cos(0.5 * angle) * self.vec

    def yrot(self, angle, q):
        self.vec = np.sin(0.5 * angle) * self._gen(1, q) + np.cos(0.5 * angle) * self.vec

    # state.py:94-97 / 146-149 : derivative gates
    def dxrot(self, angle, q):
        self.vec = 0.5 * np.cos(0.5 * angle) * self._gen(0, q) - 0.5 * np.sin(0.5 * angle) * self.vec

    def dyrot(self, angle, q):
        self.vec = 0.5 * np.cos(0.5 * angle) * self._gen(1, q) - 0.5 * np.sin(0.5 * angle) * self.vec

    # state.py:168-170 : exp(-i a/2) on bit 0, exp(+i a/2) on bit 1
    def zrot(self, angle, q):
        out = np.empty_like(self.vec)
        a, b = self._pair_view(q)
        o = out.reshape(2 ** q, 2, 2 ** (self.qnum - 1 - q))
        o[:, 0, :] = np.exp(-0.5j * angle) * a
        o[:, 1, :] = np.exp(0.5j * angle) * b
        self.vec = out

    # state.py:172-175
    def dzrot(self, angle, q):
        out = np.empty_like(self.vec)
        a, b = self._pair_view(q)
        o = out.reshape(2 ** q, 2, 2 ** (self.qnum - 1 - q))
        o[:, 0, :] = -0.5j * np.exp(-0.5j * angle) * a
        o[:, 1, :] = 0.5j * np.exp(0.5j * angle) * b
        self.vec = out

    def rot(self, axis, angle, q):
        (self.xrot, self.yrot, self.zrot)[_axis_checked(axis)](angle, q)

    def drot(self, axis, angle, q):
        (self.dxrot, self.dyrot, self.dzrot)[_axis_checked(axis)](angle, q)

    # state.py:336-356 : P0_i x 1 + P1_i x X_j  (control i, target j)
    def cnot(self, i, j):
        n = self.qnum
        if i == j or not (0 <= i < n and 0 <= j < n):
            raise ValueError("Invalid CNOT indecies {} and {}, for {} qubits.".format(i, j, n))
        idx = np.arange(2 ** n)
        cbit, tbit = 1 << (n - 1 - i), 1 << (n - 1 - j)
        src = np.where(idx & cbit, idx ^ tbit, idx)
        self.vec = self.vec[src]

    # state.py:229-241 : ladder[0] = (C01 C23 ...)(C12 C34 ...) as a MATRIX, so acting on a
    # vector the odd-start CNOTs are applied first; ladder[1] = (C12 C34 ...)(C01 C23 ...)
    # is its inverse.
    def cnot_ladder(self, stacking, periodic=False):
        n = self.qnum
        if periodic and n % 2 != 0:
            raise ValueError("CNOT ladder with periodic boundaries is ambiguous for odd qubit number.")
        upper = n if periodic else n - 1
        even = [(i, (i + 1) % n) for i in range(0, upper, 2)]
        odd = [(i, (i + 1) % n) for i in range(1, upper, 2)]
        if stacking == 0:
            order = odd + even
        elif stacking == 1:
            order = even + odd
        else:
            raise ValueError("Invalid stacking {}".format(stacking))
        for c, t in order:
            self._cnot_view(c, t)

    def _cnot_view(self, c, t):
        n = self.qnum
        if t == c + 1:
            v = self.vec.reshape(2 ** c, 2, 2, 2 ** (n - c - 2))
            tmp = v[:, 1, 0, :].copy()
            v[:, 1, 0, :] = v[:, 1, 1, :]
            v[:, 1, 1, :] = tmp
        else:
            self.cnot(c, t)

    # mc_clean.py:65 (State.multiply_matrix is absent at HEAD; it is vec = M.dot(vec))
    def multiply_observable(self, obs: OracleObservable):
        self.vec = apply_observable(obs, self.vec)

    # state.py:299-301
    def exp_ham_classical(self, angle, ham):
        self.vec = self.vec * np.exp(-1.0j * angle * ham)

    # state.py:319-321
    def ham_classical(self, ham):
        self.vec = self.vec * (-1.0j * ham)

    # state.py:107-123 : sum_q 1/2 (-i X_q)
    def x_summed(self):
        out = np.zeros_like(self.vec)
        for q in range(self.qnum):
            out += 0.5 * self._gen(0, q)
        self.vec = out

    # state.py:331-332
    def norm_error(self):
        return 1.0 - np.linalg.norm(self.vec)


def _axis_checked(axis):
    axis = int(axis)
    if axis not in (0, 1, 2):
        raise ValueError("Invalid axis {}".format(axis))
    return axis


def apply_observable(obs: OracleObservable, vec: np.ndarray) -> np.ndarray:
    """O v with O = sum w X_q + sum w Y_q + sum w Z_q + sum w Z_i Z_j (observable.py:34-79)."""
    n = obs.qnum
    out = np.zeros_like(vec)
    for kind, i, j, w in obs.terms:
        v = vec.reshape(2 ** i, 2, 2 ** (n - 1 - i))
        o = out.reshape(2 ** i, 2, 2 ** (n - 1 - i))
        if kind == "x":
            o[:, 0, :] += w * v[:, 1, :]
            o[:, 1, :] += w * v[:, 0, :]
        elif kind == "y":      # Y = [[0,-i],[i,0]]
            o[:, 0, :] += (-1j * w) * v[:, 1, :]
            o[:, 1, :] += (1j * w) * v[:, 0, :]
        elif kind == "z":
            o[:, 0, :] += w * v[:, 0, :]
            o[:, 1, :] -= w * v[:, 1, :]
        else:  # zz
            out += w * _zz_sign(n, i, j) * vec
    return out


def _z_sign(n, q):
    idx = np.arange(2 ** n)
    return 1.0 - 2.0 * ((idx >> (n - 1 - q)) & 1)


def _zz_sign(n, i, j):
    return _z_sign(n, i) * _z_sign(n, j)


def classical_ham_vector(obs: OracleObservable) -> np.ndarray:
    """state.py:273-292: H[j] = sum w z_q(j) + sum w z_i(j) z_j(j), z = +1 for bit 0."""
    n = obs.qnum
    ham = np.zeros(2 ** n, dtype=float)
    for kind, i, j, w in obs.terms:
        if kind == "z":
            ham += w * _z_sign(n, i)
        elif kind == "zz":
            ham += w * _zz_sign(n, i, j)
        else:
            raise ValueError("Non-classical observable component found. Only 'z' and 'zz' are accepted.")
    return ham


def expec_val(obs: OracleObservable, vec: np.ndarray) -> float:
    """circuit_logic/base.py:17-20."""
    return float(np.vdot(vec, apply_observable(obs, vec)).real)


# --------------------------------------------------------------------------------------
# CNOT-ladder index maps (for testing the CUDA index arithmetic)
# --------------------------------------------------------------------------------------
def ladder_scatter_map(n: int, stacking: int) -> np.ndarray:
    """dest[j] such that (ladder(stacking) v)[dest[j]] = v[j], from the bit formula.

    Acting order for stacking 0 (state.py:235-238): CNOT(1,2), CNOT(3,4), ... then
    CNOT(0,1), CNOT(2,3), ...  With b_q the bit of qubit q this gives
    b'_0=b_0, b'_1=b_1^b_0, b'_t=b_t^b_{t-1} (t even>=2), b'_t=b_t^b_{t-1}^b_{t-2} (t odd>=3).
    Stacking 1 (the inverse) swaps the parities.  In index-bit terms (bit p = n-1-q) the map
    is j' = j ^ ((j>>1)&M1) ^ ((j>>2)&M2).
    """
    m1, m2 = ladder_masks(n, stacking)
    j = np.arange(2 ** n, dtype=np.int64)
    return j ^ ((j >> 1) & m1) ^ ((j >> 2) & m2)


def ladder_masks(n: int, stacking: int):
    m1 = m2 = 0
    for t in range(1, n):
        p = n - 1 - t
        m1 |= 1 << p
        three = (t % 2 == 1) if stacking == 0 else (t % 2 == 0)
        if t >= 2 and three:
            m2 |= 1 << p
    return m1, m2


def ladder_permutation(n: int, stacking: int) -> np.ndarray:
    """Same map obtained by pushing basis states through the CNOT sequence (independent check)."""
    st = OracleState(n)
    st.vec = np.arange(2 ** n).astype(complex)
    st.cnot_ladder(stacking)
    src = st.vec.real.astype(np.int64)           # out[i] = in[src[i]]
    dest = np.empty_like(src)
    dest[src] = np.arange(2 ** n)
    return dest


# --------------------------------------------------------------------------------------
# McClean circuit                          (circuit_logic/mc_clean.py)
# --------------------------------------------------------------------------------------
def _mcclean_forward(state, axes, angles, history=None):
    L, n = angles.shape
    for q in range(n):                      # mc_clean.py:35-36
        state.yrot(np.pi / 4.0, q)
    for i in range(L):                      # mc_clean.py:38-41 / 58-62
        state.cnot_ladder(0)
        if history is not None:
            history[i] = state.vec
        for q in range(n):
            state.rot(axes[i, q], angles[i, q], q)


def mcclean_run_expec_val(n, observable, axes, angles, ini_state=None, return_state=False):
    """mc_clean.py:27-45 (exact expectation value)."""
    obs = observable if isinstance(observable, OracleObservable) else OracleObservable(n, observable)
    axes = np.asarray(axes)
    angles = np.asarray(angles, dtype=float)
    st = OracleState(n)
    if ini_state is not None:
        st.vec = np.array(ini_state, dtype=complex)
    _mcclean_forward(st, axes, angles)
    e = expec_val(obs, st.vec)
    return (e, st.vec) if return_state else e


def mcclean_grad_run(n, observable, axes, angles, ini_state=None, return_state=False):
    """mc_clean.py:47-78, step by step (history + derivative gates)."""
    obs = observable if isinstance(observable, OracleObservable) else OracleObservable(n, observable)
    axes = np.asarray(axes)
    angles = np.asarray(angles, dtype=float)
    L = angles.shape[0]
    st = OracleState(n)
    if ini_state is not None:
        st.vec = np.array(ini_state, dtype=complex)
    history = np.empty([L + 1, 2 ** n], dtype=complex)
    grad = np.empty([L, n], dtype=float)
    _mcclean_forward(st, axes, angles, history)
    history[L] = st.vec                                   # :63
    st.multiply_observable(obs)                           # :65
    e = float(np.vdot(history[L], st.vec).real)           # :66
    for i in range(L - 1, -1, -1):                        # :68
        for q in range(n):
            st.rot(axes[i, q], -angles[i, q], q)          # :69-70
        tmp = st.vec.copy()                               # :71
        for q in range(n):
            st.rot(axes[i, q], angles[i, q], q)           # :73
            st.drot(axes[i, q], -angles[i, q], q)         # :74
            grad[i, q] = -2.0 * np.vdot(history[i], st.vec).real   # :75
            st.vec = tmp.copy()                           # :76
        st.cnot_ladder(1)                                 # :77
    return (e, grad, st.vec) if return_state else (e, grad)


# ------------------------------------------------------------------------------------------
# MeynardClassifier (tutorials/meynard-classifier.ipynb cells 3, 7-8, 11, 14).  PARITY UNPINNED: the class has no
# source in the reference snapshot; the circuit below is this repo's documented definition:
#   |0..0>;  encoding layer l:    [ladder(0) unless l == 0]  Rx(data[l,q])  Ry(enc[l,q,0])  Rz(enc[l,q,1])   on every q
#            classifying layer l: [ladder(0) unless it is the very first layer]  Rx(cls[l,q,0]) Ry(cls[l,q,1]) Rz(cls[l,q,2])
#   observable: Z on qubit 0 unless given.  Gradients by the exact parameter-shift rule
#   dE/dtheta = (E(theta + pi/2) - E(theta - pi/2)) / 2  -- independent of the adjoint sweep the CUDA path uses.
# ------------------------------------------------------------------------------------------
def classifier_default_observable(n):
    return {"z": np.array([1.0] + [None] * (n - 1), dtype=object)}


def classifier_run(n, data, enc_angles, cls_angles, observable=None, return_state=False):
    obs_dict = classifier_default_observable(n) if observable is None else observable
    obs = obs_dict if isinstance(obs_dict, OracleObservable) else OracleObservable(n, obs_dict)
    data, enc, cls = np.asarray(data, dtype=float), np.asarray(enc_angles, dtype=float), np.asarray(cls_angles, dtype=float)
    st = OracleState(n)
    first = True
    for l in range(data.shape[0]):
        if not first:
            st.cnot_ladder(0)
        first = False
        for q in range(n):
            st.rot(0, data[l, q], q)
            st.rot(1, enc[l, q, 0], q)
            st.rot(2, enc[l, q, 1], q)
    for l in range(cls.shape[0]):
        if not first:
            st.cnot_ladder(0)
        first = False
        for q in range(n):
            for a in range(3):
                st.rot(a, cls[l, q, a], q)
    e = expec_val(obs, st.vec)
    return (e, st.vec) if return_state else e


def classifier_grad_run(n, data, enc_angles, cls_angles, observable=None):
    enc, cls = np.array(enc_angles, dtype=float), np.array(cls_angles, dtype=float)
    e = classifier_run(n, data, enc, cls, observable)
    enc_grad, cls_grad = np.empty_like(enc), np.empty_like(cls)
    for arr, grad, which in ((enc, enc_grad, 0), (cls, cls_grad, 1)):
        for idx in np.ndindex(arr.shape):
            keep = arr[idx]
            arr[idx] = keep + np.pi / 2
            ep = classifier_run(n, data, enc, cls, observable)
            arr[idx] = keep - np.pi / 2
            em = classifier_run(n, data, enc, cls, observable)
            arr[idx] = keep
            grad[idx] = 0.5 * (ep - em)
    return e, enc_grad, cls_grad


def mcclean_sample_grad(n, observable, axes, angles, shot_num, rng_uniform=None):
    """mc_clean.py:117-156 parameter-shift gradient with finite shots (exact E returned)."""
    obs = observable if isinstance(observable, OracleObservable) else OracleObservable(n, observable)
    axes = np.asarray(axes)
    angles = np.asarray(angles, dtype=float)
    L = angles.shape[0]
    st = OracleState(n)
    for q in range(n):
        st.yrot(np.pi / 4.0, q)
    hist = np.empty([L, 2 ** n], dtype=complex)
    for i in range(L):
        st.cnot_ladder(0)
        for q in range(n):
            st.rot(axes[i, q], angles[i, q], q)
        hist[i] = st.vec                                  # :132 (after the rotations)
    e = expec_val(obs, st.vec)
    grad = np.empty([L, n], dtype=float)
    for i in range(L):
        for dq in range(n):
            vals = []
            for shift in (np.pi / 2, -np.pi / 2):
                st.vec = hist[i].copy()
                st.rot(axes[i, dq], shift, dq)            # __manual_rot
                for j in range(i + 1, L):
                    st.cnot_ladder(0)
                    for q in range(n):
                        st.rot(axes[j, q], angles[j, q], q)
                vals.append(sample_expec_val(obs, st.vec, shot_num, rng_uniform))
            grad[i, dq] = 0.5 * (vals[0] - vals[1])
    return e, grad


# --------------------------------------------------------------------------------------
# QAOA circuit                             (circuit_logic/qaoa.py)
# --------------------------------------------------------------------------------------
def _check_parameters(betas, gammas, p):
    if betas.size != p or gammas.size != p:       # qaoa.py:186-191
        raise ValueError("Wrong amount of parameters. Expected {} and {}.".format(p, p))


def qaoa_run_expec_val(n, observable, betas, gammas, ini_state=None, return_state=False):
    """qaoa.py:23-38."""
    obs = observable if isinstance(observable, OracleObservable) else OracleObservable(n, observable)
    betas = np.asarray(betas, dtype=float)
    gammas = np.asarray(gammas, dtype=float)
    _check_parameters(betas, gammas, betas.size)
    ham = classical_ham_vector(obs)
    st = OracleState(n, "+")
    if ini_state is not None:
        st.vec = np.array(ini_state, dtype=complex)
    for i in range(betas.size):
        st.exp_ham_classical(gammas[i], ham)
        for q in range(n):
            st.xrot(betas[i], q)
    e = expec_val(obs, st.vec)
    return (e, st.vec) if return_state else e


def qaoa_grad_run(n, observable, betas, gammas, ini_state=None, return_state=False):
    """qaoa.py:40-70, step by step."""
    obs = observable if isinstance(observable, OracleObservable) else OracleObservable(n, observable)
    betas = np.asarray(betas, dtype=float)
    gammas = np.asarray(gammas, dtype=float)
    p = betas.size
    _check_parameters(betas, gammas, p)
    ham = classical_ham_vector(obs)
    st = OracleState(n, "+")
    if ini_state is not None:
        st.vec = np.array(ini_state, dtype=complex)
    hist = np.empty([2 * p + 1, 2 ** n], dtype=complex)
    grad = np.empty([p, 2], dtype=float)
    for i in range(p):                                    # :49-53
        hist[2 * i] = st.vec
        st.exp_ham_classical(gammas[i], ham)
        hist[2 * i + 1] = st.vec
        for q in range(n):
            st.xrot(betas[i], q)
    hist[2 * p] = st.vec                                  # :54
    psi_final = st.vec.copy()
    st.vec = st.vec * ham                                 # :56
    e = float(np.vdot(hist[2 * p], st.vec).real)          # :57
    for i in range(p - 1, -1, -1):                        # :59-69
        for q in range(n):
            st.xrot(-betas[i], q)
        tmp = st.vec.copy()
        st.x_summed()
        grad[i, 0] = -2.0 * np.vdot(hist[2 * i + 1], st.vec).real
        st.vec = tmp
        st.exp_ham_classical(-gammas[i], ham)
        tmp = st.vec.copy()
        st.ham_classical(ham)
        grad[i, 1] = -2.0 * np.vdot(hist[2 * i], st.vec).real
        st.vec = tmp
    return (e, grad, psi_final) if return_state else (e, grad)


# --------------------------------------------------------------------------------------
# finite-shot sampling                     (base.py:22-33, qaoa.py:196-198)
# --------------------------------------------------------------------------------------
def term_expectations(obs: OracleObservable, vec: np.ndarray) -> np.ndarray:
    """<P_k> for every Pauli term, projector order.  prob_k = (1+<P_k>)/2 (observable.py:126-177)."""
    n = obs.qnum
    out = np.empty(len(obs.terms))
    for k, (kind, i, j, _w) in enumerate(obs.terms):
        one = OracleObservable.__new__(OracleObservable)
        one.qnum, one.terms = n, [(kind, i, j, 1.0)]
        out[k] = expec_val(one, vec)
    return out


def _rv_discrete_rvs(pk: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """scipy.stats.rv_discrete(values=(xk, pk)).rvs: index = first k with cumsum(pk)[k] >= u.

    scipy's rv_sample._ppf computes ``argmax(cumsum(pk) >= u)``, which returns 0 when no
    entry qualifies (u above the last cdf value through rounding).
    """
    cdf = np.cumsum(pk)
    idx = np.searchsorted(cdf, uniforms, side="left")
    idx[idx >= pk.size] = 0
    return idx


def sample_expec_val(obs: OracleObservable, vec: np.ndarray, shot_num: int, rng_uniform=None) -> float:
    """base.py:22-33: per-term Bernoulli(+-w) estimate; one uniform(size=shot_num) draw per term
    from the global numpy stream (scipy ``rvs`` with random_state=None), in projector order."""
    draw = rng_uniform if rng_uniform is not None else (lambda size: np.random.uniform(size=size))
    total = 0.0
    for k, ek in enumerate(term_expectations(obs, vec)):
        w = obs.terms[k][3]
        prob = 0.5 * (1.0 + ek)
        idx = _rv_discrete_rvs(np.array([prob, 1.0 - prob]), draw(shot_num))
        total += np.array([w, -w])[idx].mean()
    return float(total)


def sample_bitstrings(vec: np.ndarray, uniforms: np.ndarray) -> np.ndarray:
    """qaoa.py:196-198 / mc_clean.py:259-261: indices drawn from |psi|^2 by inverse CDF."""
    return _rv_discrete_rvs(np.abs(vec) ** 2, np.asarray(uniforms, dtype=float))


# --------------------------------------------------------------------------------------
# finite differences (tutorials/qaoa-max-cut.ipynb cell 6): an independent gradient check
# --------------------------------------------------------------------------------------
def finite_difference_grad(fun, params: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    params = np.array(params, dtype=float)
    g = np.empty_like(params)
    it = np.nditer(params, flags=["multi_index"])
    for _ in it:
        k = it.multi_index
        p1, p2 = params.copy(), params.copy()
        p1[k] += eps
        p2[k] -= eps
        g[k] = (fun(p1) - fun(p2)) / (2 * eps)
    return g
