"""State: complex128 state vector resident in GPU memory.

Mirrors qradient/physical_components/state.py (dialect used by circuit_logic/*: a separate
``Gates`` object assigned to ``state.gates``; method names per SURVEY.md appendix A).  Every
method launches CUDA kernels through the C ABI; ``vec`` downloads / uploads the full vector and
is therefore expensive -- the circuits' ``run_expec_val`` / ``grad_run`` never touch it.
"""
import ctypes
import warnings

import numpy as np

from .. import _lib
from .gates import Gates


class _VecView(np.ndarray):
    """Host copy of the device vector that writes item assignments back to the device, so that
    reference idioms like ``state.vec[:] = tmp_vec`` (mc_clean.py:76) keep working."""

    def __new__(cls, arr, owner):
        obj = np.asarray(arr).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = getattr(obj, '_owner', None) if obj is not None and obj.shape == self.shape else None

    def __setitem__(self, key, value):
        np.ndarray.__setitem__(self, key, value)
        owner = getattr(self, '_owner', None)
        if owner is not None and self.shape == (owner._N,):
            owner._upload(np.asarray(self))


class State:
    def __init__(self, qubit_number, ini='0', device=0):
        self._lib = _lib.lib()
        self._qnum = int(qubit_number)
        self._N = 2**self._qnum
        self._ini = ini
        self._gates = None
        self._ctx = None
        self._check_ini(ini)
        h = ctypes.c_void_p()
        self._lib.call('qr_ctx_create', self._qnum, int(device), ctypes.byref(h))
        self._ctx = h
        self.device = int(device)
        self.reset()

    # ---------------------------------------------------------------------------------------
    @staticmethod
    def _check_ini(ini):
        if ini not in ('0', '+'):
            raise ValueError('Invalid initialization format {}.'.format(ini))   # state.py:71

    def reset(self, ini=None):
        """state.py:61-71; ``reset('+')`` also changes the reference point of later resets
        (qaoa.py:17,26)."""
        if ini is not None:
            self._check_ini(ini)
            self._ini = ini
        self._lib.call('qr_state_init', self._ctx, 0 if self._ini == '0' else 1)
        # state.py:72-77: the dense-operator tracking attributes return to their initial values as well
        if 'lhs' in self.__dict__:
            import scipy.sparse as sp
            self.lhs = sp.identity(2**self.qnum, dtype='complex', format='csr')
        if 'center_matrix' in self.__dict__:
            self.center_matrix = self._center_matrix_ini.copy()

    @property
    def qnum(self):
        return self._qnum

    # -- vec ---------------------------------------------------------------------------------
    @property
    def vec(self):
        host = np.empty(self._N, dtype=np.complex128)
        self._lib.call('qr_state_download', self._ctx, _lib.ptr(host), self._N)
        return _VecView(host, self)

    @vec.setter
    def vec(self, value):
        self._upload(value)

    def _upload(self, value):
        arr = np.ascontiguousarray(np.asarray(value), dtype=np.complex128)
        if arr.shape != (self._N,):
            raise ValueError('state vector must have shape ({},), got {}'.format(self._N, arr.shape))
        self._lib.call('qr_state_upload', self._ctx, _lib.ptr(arr), self._N)

    def save(self, slot):
        """Device-side snapshot of the vector (stands in for `state_history[i] = state.vec`)."""
        self._lib.call('qr_state_save', self._ctx, int(slot))

    def load(self, slot):
        self._lib.call('qr_state_load', self._ctx, int(slot))

    def free_snapshots(self):
        self._lib.call('qr_state_free_snapshots', self._ctx)

    def device_ptr(self):
        p = ctypes.c_void_p()
        self._lib.call('qr_state_device_ptr', self._ctx, ctypes.byref(p))
        return p.value

    # -- gates container ---------------------------------------------------------------------
    @property
    def gates(self):
        return self._gates

    @gates.setter
    def gates(self, g):
        if not isinstance(g, Gates):
            raise TypeError('state.gates must be a qradient_b200 Gates object')
        if g.qnum != self._qnum:
            raise ValueError('Gates built for {} qubits attached to a {}-qubit state'.format(g.qnum, self._qnum))
        self._gates = g
        g._state = self
        if g.ham_observable is not None:
            self._load_ham(g.ham_observable)

    def _load_ham(self, observable):
        self._lib.call('qr_ham_load', self._ctx, observable._handle)

    def _download_ham(self):
        out = np.empty(self._N, dtype=np.float64)
        self._lib.call('qr_ham_download', self._ctx, _lib.ptr(out), self._N)
        return out

    # -- rotations and derivatives (state.py:90-97,142-149,168-175) ----------------------------
    def xrot(self, angle, i): self._lib.call('qr_apply_rot', self._ctx, 0, float(angle), int(i))
    def yrot(self, angle, i): self._lib.call('qr_apply_rot', self._ctx, 1, float(angle), int(i))
    def zrot(self, angle, i): self._lib.call('qr_apply_rot', self._ctx, 2, float(angle), int(i))
    def dxrot(self, angle, i): self._lib.call('qr_apply_drot', self._ctx, 0, float(angle), int(i))
    def dyrot(self, angle, i): self._lib.call('qr_apply_drot', self._ctx, 1, float(angle), int(i))
    def dzrot(self, angle, i): self._lib.call('qr_apply_drot', self._ctx, 2, float(angle), int(i))

    # -- entanglers (state.py:198-199, 243-251) ------------------------------------------------
    def cnot(self, i, j):
        self._lib.call('qr_apply_cnot', self._ctx, int(i), int(j))

    def cnot_ladder(self, stacking):
        periodic = bool(self._gates.ladder_periodic) if self._gates is not None else False
        self._lib.call('qr_apply_cnot_ladder', self._ctx, int(stacking), int(periodic))

    # -- observable / Hamiltonian multiplications ---------------------------------------------
    def multiply_matrix(self, matrix):
        """mc_clean.py:65 ``state.multiply_matrix(observable.matrix)``: vec = O vec on the device.
        Accepts an Observable or the host matrix obtained from ``Observable.matrix``."""
        obs = getattr(matrix, '_qr_observable', matrix)
        if hasattr(obs, '_handle'):
            self._lib.call('qr_apply_observable', self._ctx, obs._handle)
            return
        # any other 2^n x 2^n host matrix (dense or scipy.sparse): vec = M.dot(vec) as one dense matrix-vector product on
        # the device -- an inspection aid for small registers, like the reference's own 2^n x 2^n matrices
        if self.qnum > 12:
            raise TypeError('multiply_matrix with an arbitrary host matrix is limited to 12 qubits (dense 2^n x 2^n product); '
                            'pass an Observable (or Observable.matrix) for larger registers')
        dense = matrix.toarray() if hasattr(matrix, 'toarray') else np.asarray(matrix)
        if dense.shape != (self._N, self._N):
            raise ValueError('matrix must have shape ({0}, {0})'.format(self._N))
        self.load_dense(dense)
        self.apply_dense()

    def exp_ham_classical(self, angle):            # state.py:299-301
        self._lib.call('qr_apply_exp_ham', self._ctx, float(angle))

    def exp_ham_classical_component(self, angle, i):   # state.py:309-311
        self._lib.call('qr_apply_exp_ham_component', self._ctx, self._gates.ham_observable._handle, int(i), float(angle))

    def ham_classical(self):                       # state.py:319-321
        self._lib.call('qr_apply_ham', self._ctx, 0)

    def mul_ham_classical(self):                   # qaoa.py:56  vec *= gates.classical_ham
        self._lib.call('qr_apply_ham', self._ctx, 1)

    def x_summed(self):                            # state.py:121-123
        self._lib.call('qr_apply_x_summed', self._ctx)

    def load_permutation(self, perm):
        """Keep an index permutation on the device for `permute()` (measurement orderings)."""
        perm = np.ascontiguousarray(perm, dtype=np.int64)
        self._lib.call('qr_perm_load', self._ctx, _lib.ptr(perm), int(perm.size))

    def permute(self):
        """vec[k] <- vec[perm[k]] on the device."""
        self._lib.call('qr_state_permute', self._ctx)

    def load_dense(self, matrix):
        """Keep a dense 2^n x 2^n matrix on the device for `apply_dense()` (eigenbasis of an observable, n <= 12)."""
        m = np.ascontiguousarray(matrix, dtype=np.complex128)
        if m.ndim != 2 or m.shape[0] != m.shape[1]:
            raise ValueError('matrix must be square')
        self._lib.call('qr_dense_load', self._ctx, _lib.ptr(m), int(m.shape[0]))

    def apply_dense(self):
        """vec <- M vec on the device."""
        self._lib.call('qr_state_apply_dense', self._ctx)

    def norm_error(self):                          # state.py:331-332
        out = ctypes.c_double()
        self._lib.call('qr_norm2', self._ctx, ctypes.byref(out))
        return 1. - np.sqrt(out.value)

    # -- dialect-A names of the HEAD state.py ---------------------------
    rot_classical_ham = exp_ham_classical
    rot_classical_ham_component = exp_ham_classical_component
    xrot_all = x_summed

    def load_xrots(self): return None
    def load_yrots(self): return None
    def load_zrots(self): return None
    def load_xrot_all(self): return None
    def load_cnots(self, which): return None

    def load_cnot_ladder(self, periodic=False):
        if self._gates is None:
            self.gates = Gates(self._qnum)
        self._gates.add_cnot_ladder(periodic)

    def load_classical_ham(self, observable, include_individual_components=False):
        if self._gates is None:
            self.gates = Gates(self._qnum)
        self._gates.add_classical_ham(observable, include_individual_components)
        return self

    # -- options / perf ------------------------------------------------------------------------
    def set_option(self, name, value):
        self._lib.call('qr_set_option', self._ctx, _lib.OPT[name], int(value))

    def perf(self):
        p = _lib.QrPerf()
        self._lib.call('qr_perf_last', self._ctx, ctypes.byref(p))
        return p.as_dict()

    # -- state.py:39-59: dense-operator tracking attributes (host-side scipy objects, as in the reference; every gate
    #    variant that would use them is a 'Not implemented.' stub there and here) ---------------------------------
    def activate_lefthandside(self):
        if 'lhs' not in self.__dict__:
            import scipy.sparse as sp
            self.lhs = sp.identity(2**self.qnum, dtype='complex', format='csr')

    def activate_center_matrix(self):
        if 'center_matrix' not in self.__dict__:
            import scipy.sparse as sp
            self.center_matrix = sp.csr_matrix((2**self.qnum, 2**self.qnum), dtype='complex')
            self._center_matrix_ini = sp.csr_matrix((2**self.qnum, 2**self.qnum), dtype='complex')

    def set_center_matrix(self, matrix):
        if 'center_matrix' not in self.__dict__:
            raise AttributeError('center_matrix is not initialized yet.')     # state.py:57
        self.center_matrix = matrix.copy()
        self._center_matrix_ini = matrix.copy()

    # the reference's *_lhs / *_center_matrix variants are 'Not implemented.' stubs (state.py:99-103 ...)
    _STUB_GATES = ('xrot', 'xrot_all', 'x_summed', 'yrot', 'zrot', 'cnot', 'cnot_ladder', 'rot_classical_ham',
                   'rot_classical_ham_component', 'classical_ham', 'exp_ham_classical', 'exp_ham_classical_component', 'ham_classical')

    def __getattr__(self, name):
        for suffix in ('_lhs', '_center_matrix'):
            if not (name.endswith(suffix) and name[:-len(suffix)] in self._STUB_GATES):
                continue

            def stub(*args, **kwargs):
                warnings.warn('Not implemented.')
            return stub
        raise AttributeError(name)

    def close(self):
        if getattr(self, '_ctx', None) is not None:
            try:
                self._lib.cdll.qr_ctx_destroy(self._ctx)
            except Exception:
                pass
            self._ctx = None

    def __del__(self):
        self.close()
