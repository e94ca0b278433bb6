"""Observable: dictionary of Pauli terms -> device term list.

Mirrors qradient/physical_components/observable.py (constructor as called from
circuit_logic/base.py:12: ``Observable(qubit_number, observable, store_components=False)``).
The reference materialises a 2^n x 2^n CSR matrix (observable.py:34-79); here the observable
stays a list of (kind, qubit, qubit, weight) terms that the CUDA kernels evaluate on the fly.
``matrix`` is still available as a lazily built host CSR matrix for small registers because
users inspect it (tutorials/optimize-test.ipynb: ``eigsh(crct.observable.matrix)``); it is
never used by the simulation path.
"""
import ctypes
import warnings

import numpy as np

from .. import _lib

_KNOWN = ('x', 'y', 'z', 'zz')


class Projector:
    """Descriptor of the +1 eigenspace projector of one Pauli term (observable.py:126-177).

    The reference stores a dense mask or CSR matrix per term; the device path only needs the
    term's expectation value (prob = (1 + <P>)/2), so this is a plain record.
    """

    def __init__(self, key, *args):
        if key not in _KNOWN:
            raise ValueError('Unknown key for projector {}.'.format(key))
        self.key = key
        self.qubits = tuple(int(a) for a in args)
        self.is_classical = key in ('z', 'zz')

    def __repr__(self):
        return 'Projector({!r}, {})'.format(self.key, ', '.join(map(str, self.qubits)))

    qnum = None                 # class attribute like the reference's (observable.py:170-171)

    @staticmethod
    def set_qnum(qnum):
        Projector.qnum = qnum

    def dot(self, vec):
        """Host-side inspection aid with the semantics of observable.py:173-177: the +1-eigenspace projector applied to a
        host vector, matrix-free (x / y: (1 + P)/2 on the strided pairs; z / zz: the 0/1 mask of observable.py:142-166).
        The device path never calls it -- it takes the term's expectation value instead."""
        n = Projector.qnum
        v = np.asarray(vec, dtype=complex)
        if n is None or v.shape != (2 ** n,):
            raise ValueError('Projector.dot needs Projector.set_qnum(n) and a vector with 2^n entries')
        if self.key in ('x', 'y'):
            q = self.qubits[0]
            w = v.reshape(2 ** q, 2, 2 ** (n - 1 - q))
            out = np.empty_like(w)
            off = 1. if self.key == 'x' else -1.j          # [[.5, .5], [.5, .5]] resp. [[.5, -.5j], [.5j, .5]]
            out[:, 0, :] = .5 * w[:, 0, :] + .5 * off * w[:, 1, :]
            out[:, 1, :] = .5 * np.conj(off) * w[:, 0, :] + .5 * w[:, 1, :]
            return out.reshape(-1)
        idx = np.arange(2 ** n)
        bit = lambda q: (idx >> (n - 1 - q)) & 1
        if self.key == 'z':
            mask = bit(self.qubits[0]) == 0
        else:
            mask = bit(self.qubits[0]) == bit(self.qubits[1])
        return mask.astype(complex) * v


class Observable:
    def __init__(self, qubit_number, observable, store_components=False):
        self.qnum = int(qubit_number)
        self.info = observable
        self.dict = observable
        self.store_components = store_components
        self.has_loaded_projectors = False
        self._lib = _lib.lib()
        self._handle = None
        self._matrix = None
        self.check_observable(known_keys=list(_KNOWN))
        self.load_matrix(observable)

    # -- validation (observable.py:19-26, 57-64, 106-124) ------------------------------------
    def check_observable(self, known_keys, warning=None):
        for key in list(self.info.keys()):
            if key not in known_keys:
                warnings.warn(warning if warning is not None
                              else 'Unknown element of observable {} will be ignored.'.format(key))

    @staticmethod
    def _weight_check(weight, component):
        if abs(weight) < 10.**-15:
            warnings.warn('Weight in observable {} is zero or almost zero. '
                          'If you dont\'t want to include it, set it to None.'.format(component))

    def load_matrix(self, observable, store_components=None):
        """Parse the dictionary into the term list (projector order) and create the device handle."""
        n = self.qnum
        kinds, qi, qj, ws = [], [], [], []
        for key in ('x', 'y', 'z'):
            if key in observable and observable[key] is not None:
                arr = observable[key]
                if len(arr) != n:
                    raise ValueError('Inconsistent shapes in observable dictionary. Cannot infer qubit_number.')
                for i, w in enumerate(arr):
                    if w is not None:
                        self._weight_check(w, key)
                        kinds.append(_lib.TERM_KIND[key]); qi.append(i); qj.append(0); ws.append(float(w))
        if 'zz' in observable and observable['zz'] is not None:
            zz = np.asarray(observable['zz'], dtype=object)
            if zz.shape != (n, n):
                raise ValueError('Inconsistent shapes in observable dictionary. Cannot infer qubit_number.')
            for i in range(n):
                for j in range(i + 1):
                    if zz[i, j] is not None:
                        raise ValueError(
                            'zz of observable should be a upper triangular {0} by {0} matrix. Diagonal and lower '
                            'triangle should contain None\'s, not {1}.'.format(n, zz[i, j]))
                for j in range(i + 1, n):
                    if zz[i, j] is not None:
                        self._weight_check(zz[i, j], 'zz')
                        kinds.append(_lib.TERM_KIND['zz']); qi.append(i); qj.append(j); ws.append(float(zz[i, j]))
        self.term_kinds = np.array(kinds, dtype=np.int32)
        self.term_qi = np.array(qi, dtype=np.int32)
        self.term_qj = np.array(qj, dtype=np.int32)
        self.term_weights = np.array(ws, dtype=np.float64)
        self.num_components = len(ws)
        if self.store_components:
            # observable.py:75-79 (weights are not abs-normalised there either)
            self.weight_distribution = self.term_weights / np.sum(self.term_weights)
        self._destroy()
        h = ctypes.c_void_p()
        self._lib.call('qr_obs_create', n, len(ws), _lib.ptr(self.term_kinds), _lib.ptr(self.term_qi),
                       _lib.ptr(self.term_qj), _lib.ptr(self.term_weights), ctypes.byref(h))
        self._handle = h
        self._matrix = None

    # -- term helpers ----------------------------------------------------------------------
    @property
    def is_classical(self):
        return bool(np.all(self.term_kinds >= 2))

    def scale(self):
        """sum_k |w_k|: operator-norm bound used by the parity criterion."""
        return float(np.abs(self.term_weights).sum()) if self.num_components else 1.0

    def component(self, k):
        """Single-term Observable (weight included), e.g. for component sampling (mc_clean.py:103)."""
        kind = ('x', 'y', 'z', 'zz')[int(self.term_kinds[k])]
        n = self.qnum
        if kind == 'zz':
            m = np.full((n, n), None)
            m[int(self.term_qi[k]), int(self.term_qj[k])] = float(self.term_weights[k])
            return Observable(n, {'zz': m})
        arr = np.full(n, None)
        arr[int(self.term_qi[k])] = float(self.term_weights[k])
        return Observable(n, {kind: arr})

    # -- projectors (observable.py:82-104) ---------------------------------------------------
    def load_projectors(self):
        if self.has_loaded_projectors:
            return None
        projs = []
        Projector.set_qnum(self.qnum)                  # observable.py:89
        for k in range(self.num_components):
            kind = ('x', 'y', 'z', 'zz')[int(self.term_kinds[k])]
            if kind == 'zz':
                projs.append(Projector('zz', self.term_qi[k], self.term_qj[k]))
            else:
                projs.append(Projector(kind, self.term_qi[k]))
        self.projectors = np.array(projs, dtype=object)
        self.projector_weights = self.term_weights.copy()
        self.has_loaded_projectors = True

    # -- host matrix for inspection (small registers only) -----------------------------------
    @property
    def matrix(self):
        if self._matrix is None:
            self._matrix = self._host_matrix(range(self.num_components))
            self._matrix._qr_observable = self
        return self._matrix

    @property
    def component_array(self):
        out = np.empty(self.num_components, dtype=object)
        for k in range(self.num_components):
            m = self._host_matrix([k])
            m._qr_observable = self.component(k)
            out[k] = m
        return out

    def _host_matrix(self, ks):
        import scipy.sparse as sp
        n = self.qnum
        if n > 16:
            raise MemoryError('Observable.matrix is a host-side inspection aid for <= 16 qubits; '
                              'the device path never builds it.')
        pauli = {0: sp.csr_matrix([[0., 1.], [1., 0.]], dtype=complex),
                 1: sp.csr_matrix([[0., -1.j], [1.j, 0.]], dtype=complex),
                 2: sp.csr_matrix([[1., 0.], [0., -1.]], dtype=complex)}
        mat = sp.csr_matrix((2**n, 2**n), dtype=complex)
        for k in ks:
            kind, i, j, w = int(self.term_kinds[k]), int(self.term_qi[k]), int(self.term_qj[k]), self.term_weights[k]
            ops = [sp.identity(2, dtype=complex, format='csr')] * n
            if kind == 3:
                ops[i] = pauli[2]; ops[j] = pauli[2]
            else:
                ops[i] = pauli[kind]
            full = ops[0]
            for o in ops[1:]:
                full = sp.kron(full, o, format='csr')
            mat = mat + w * full
        return mat.asformat('csr')

    def _destroy(self):
        if getattr(self, '_handle', None) is not None:
            try:
                self._lib.cdll.qr_obs_destroy(self._handle)
            except Exception:
                pass
            self._handle = None

    def __del__(self):
        self._destroy()
