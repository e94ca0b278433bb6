"""McClean random-Pauli-rotation ansatz (reference: circuit_logic/mc_clean.py).

Circuit (mc_clean.py:35-41): Ry(pi/4) on every qubit, then L x [CNOT ladder, one Pauli rotation
per qubit with axis axes[i,q] in {0,1,2} = {X,Y,Z} and angle angles[i,q]].

grad_run replaces the reference's (L+1)-copy history algorithm (mc_clean.py:47-78) by the
two-vector adjoint recurrence  grad[i,q] = Im<lambda_i| P_q |psi_i>  executed in fused CUDA
tile passes (qradient_b200/csrc/qr_tile.cuh); results agree to ~1e-15.
"""
import ctypes
import warnings

import numpy as np

from .. import _lib
from ..physical_components import Gates
from .base import ParametrizedCircuit


class McClean(ParametrizedCircuit):
    def __init__(self, qubit_number, observable, layer_number, use_observable_components=False, **kwargs):
        ParametrizedCircuit.init(self, qubit_number, observable, use_observable_components,
                                 device=kwargs.get('device', 0))
        self.lnum = layer_number
        # same draws, same order as mc_clean.py:13-14 (axes first, then angles, global numpy stream)
        self.axes = kwargs.get('axes', (3 * np.random.rand(self.lnum, self.qnum)).astype('int'))
        self.angles = kwargs.get('angles', 2 * np.pi * np.random.rand(self.lnum, self.qnum))
        self.state.gates = Gates(self.qnum) \
            .add_xrots() \
            .add_yrots() \
            .add_zrots() \
            .add_cnot_ladder()

    # -- helpers ------------------------------------------------------------------------------
    def _params(self):
        axes = np.asarray(self.axes)
        angles = np.asarray(self.angles, dtype=np.float64)
        if axes.shape != (self.lnum, self.qnum) or angles.shape != (self.lnum, self.qnum):
            raise ValueError('axes and angles must have shape ({}, {})'.format(self.lnum, self.qnum))
        bad = (axes < 0) | (axes > 2)
        if np.any(bad):
            raise ValueError('Invalid axis {}'.format(axes[bad].ravel()[0]))   # mc_clean.py:392
        return _lib.as_i32(axes), np.ascontiguousarray(angles)

    def _adopt(self, ini_state):
        if ini_state is None:
            return 0
        self.state.vec = ini_state      # mc_clean.py:32
        return 1

    # -- mc_clean.py:27-45 --------------------------------------------------------------------
    def run_expec_val(self, hide_progbar=True, exact_expec_val=True, shot_num=1, ini_state=None):
        '''Runs the circuit and returns the expectation value under observable'''
        axes, angles = self._params()
        use_current = self._adopt(ini_state)
        e = ctypes.c_double()
        self._lib.call('qr_mcclean_expec', self.state._ctx, self.lnum, _lib.ptr(axes), _lib.ptr(angles),
                       self.observable._handle, use_current, ctypes.byref(e))
        if exact_expec_val:
            return e.value
        return self.sample_expec_val(shot_num)

    # -- mc_clean.py:47-78 --------------------------------------------------------------------
    def grad_run(self, hide_progbar=True, ini_state=None):
        axes, angles = self._params()
        use_current = self._adopt(ini_state)
        e = ctypes.c_double()
        grad = np.empty([self.lnum, self.qnum], dtype='double')
        self._lib.call('qr_mcclean_grad', self.state._ctx, self.lnum, _lib.ptr(axes), _lib.ptr(angles),
                       self.observable._handle, use_current, ctypes.byref(e), _lib.ptr(grad))
        return e.value, grad

    # -- mc_clean.py:80-115 ----------------------------------------------------------------------
    def grad_run_with_component_sampling(self, hide_progbar=True, ini_state=None):
        """Exact E under the full observable, gradient under ONE observable component drawn with
        np.random.choice(p=weight_distribution) (mc_clean.py:101-103).  Needs
        use_observable_components=True at construction, like the reference."""
        if not getattr(self.observable, 'store_components', False):
            raise AttributeError('construct the circuit with use_observable_components=True')
        ini = None if ini_state is None else np.array(ini_state, dtype=complex)
        expec_val = self.run_expec_val(ini_state=None if ini is None else ini.copy())
        component = np.random.choice(np.arange(self.observable.num_components), p=self.observable.weight_distribution)
        axes, angles = self._params()
        use_current = self._adopt(None if ini is None else ini.copy())
        e = ctypes.c_double()
        grad = np.empty([self.lnum, self.qnum], dtype='double')
        comp_obs = self.observable.component(component)
        self._lib.call('qr_mcclean_grad', self.state._ctx, self.lnum, _lib.ptr(axes), _lib.ptr(angles),
                       comp_obs._handle, use_current, ctypes.byref(e), _lib.ptr(grad))
        return expec_val, grad

    # -- extension: optimiser loop on the device (optimization.py:69-91 repeated, no host round trips) --
    def optimize_on_device(self, rule, hyper, iteration, steps, m=None, v=None, keep_param_history=True):
        """`steps` x (grad_run, parameter update) enqueued back to back on the device.

        rule: 0 Adam, 1 GradientDescent (constant step), 2 RateDecayOnPlateau; hyper: float64[8] =
        (step_size, beta1, beta2, eps, plateau_length, decay_rate, cost, plateau_counter), updated in place;
        m, v: Adam moments float64[L, n], updated in place.  self.angles is updated in place.
        Returns (cost_history float64[steps], param_history float64[steps, L, n] or None, iteration)."""
        axes, angles = self._params()
        hyper = np.ascontiguousarray(hyper, dtype=np.float64)
        cost = np.zeros(steps, dtype=np.float64)
        hist = np.zeros((steps, self.lnum, self.qnum), dtype=np.float64) if keep_param_history else None
        it = ctypes.c_int(int(iteration))
        mm = np.ascontiguousarray(m, dtype=np.float64) if m is not None else None
        vv = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
        self._lib.call('qr_mcclean_optimize', self.state._ctx, self.lnum, _lib.ptr(axes), _lib.ptr(angles), self.observable._handle,
                       int(rule), _lib.ptr(hyper), ctypes.byref(it), _lib.ptr(mm) if mm is not None else None,
                       _lib.ptr(vv) if vv is not None else None, int(steps), _lib.ptr(cost), _lib.ptr(hist) if hist is not None else None)
        self.angles = angles
        if m is not None:
            m[...] = mm
            v[...] = vv
        return cost, hist, it.value, hyper

    # -- extension: many parameter sets at once (timing-test.ipynb cell 6 host loop) -------------
    def grad_run_batch(self, angles, axes=None):
        """angles: [B, L, n]; axes: [B, L, n] or [L, n] (default: self.axes).
        Returns (float64[B], float64[B, L, n])."""
        angles = np.ascontiguousarray(angles, dtype=np.float64)
        if angles.ndim != 3 or angles.shape[1:] != (self.lnum, self.qnum):
            raise ValueError('angles must have shape (B, {}, {})'.format(self.lnum, self.qnum))
        B = angles.shape[0]
        axes = np.asarray(self.axes if axes is None else axes)
        if axes.ndim == 2:
            axes = np.broadcast_to(axes, angles.shape)
        if axes.shape != angles.shape:
            raise ValueError('axes must have shape (B, L, n) or (L, n)')
        if axes.size and (axes.min() < 0 or axes.max() > 2):     # two reductions, no B*L*n temporaries
            raise ValueError('Invalid axis')
        axes = _lib.as_i32(axes)
        e = np.empty(B, dtype=np.float64)
        grad = np.empty(angles.shape, dtype=np.float64)
        self._lib.call('qr_mcclean_grad_batch', self.state._ctx, B, self.lnum, _lib.ptr(axes), _lib.ptr(angles),
                       self.observable._handle, _lib.ptr(e), _lib.ptr(grad))
        return e, grad

    # -- mc_clean.py:117-156: parameter-shift gradient with finite shots ------------------------
    def sample_grad(self, hide_progbar=True, shot_num=1, exact_expec_val=True, ini_state=None):
        axes, angles = self._params()
        n, L = self.qnum, self.lnum
        st = self.state
        if ini_state is None:
            st.reset()
        else:
            st.vec = ini_state
        grad = np.ndarray([L, n], dtype='double')
        for q in range(n):
            st.yrot(np.pi / 4., q)
        for i in range(L):
            st.cnot_ladder(0)
            for q in range(n):
                self._rot(i, q)
            st.save(i)                                  # mc_clean.py:132 state_history[i] (device snapshot, no host round trip)
        expec_val = self.expec_val() if exact_expec_val else self.sample_expec_val(shot_num)
        for i in range(L):
            for dq in range(n):
                shifted = []
                for shift in (np.pi / 2, -np.pi / 2):
                    st.load(i)
                    self._manual_rot(i, dq, shift)
                    self._forward_tail(i + 1)
                    shifted.append(self.sample_expec_val(shot_num))
                grad[i, dq] = .5 * (shifted[0] - shifted[1])
        st.free_snapshots()
        return expec_val, grad

    # -- mc_clean.py:158-198: the same with one observable component drawn per parameter --------
    def sample_grad_with_component_sampling(self, hide_progbar=True, shot_num=1, exact_expec_val=True, ini_state=None):
        if not getattr(self.observable, 'store_components', False):
            raise AttributeError('construct the circuit with use_observable_components=True')
        axes, angles = self._params()
        n, L = self.qnum, self.lnum
        st = self.state
        if ini_state is None:
            st.reset()
        else:
            st.vec = ini_state
        grad = np.ndarray([L, n], dtype='double')
        for q in range(n):
            st.yrot(np.pi / 4., q)
        for i in range(L):
            st.cnot_ladder(0)
            for q in range(n):
                self._rot(i, q)
            st.save(i)                                  # mc_clean.py:173 state_history[i] (device snapshot)
        expec_val = self.expec_val() if exact_expec_val else self.sample_expec_val(shot_num)
        for i in range(L):
            for dq in range(n):
                component = np.random.choice(np.arange(self.observable.num_components), p=self.observable.weight_distribution)
                shifted = []
                for shift in (np.pi / 2, -np.pi / 2):
                    st.load(i)
                    self._manual_rot(i, dq, shift)
                    self._forward_tail(i + 1)
                    shifted.append(self.sample_component_expec_val(shot_num, component))
                grad[i, dq] = .5 * (shifted[0] - shifted[1])
        st.free_snapshots()
        return expec_val, grad

    # -- mc_clean.py:200-205, 270-275: renamed methods keep their deprecation stubs -------------
    def sample_grad_observable(self, *args):
        warnings.warn('Method sample_grad_observable is now called sample_grad_dense.', DeprecationWarning, stacklevel=2)

    def sample_grad_observable_with_component_sampling(self, *args):
        warnings.warn('Method sample_grad_observable_with_component_sampling is now called '
                      'sample_grad_dense_with_component_sampling.', DeprecationWarning, stacklevel=2)

    # -- mc_clean.py:207-268: finite-shot gradient, observable measured in its eigenbasis ---------
    def sample_grad_dense(self, shot_num=1, hide_progbar=True, exact_expec_val=True, ini_state=None):
        """Parameter-shift gradient where every shifted circuit is measured `shot_num` times in the eigenbasis
        of the observable (returns the exact expectation value, like the reference).

        The reference diagonalises the dense 2^n x 2^n observable and propagates dense left-hand-side
        matrices (mc_clean.py:221-244).  Here each shifted state is pushed through the remaining layers on the
        device and then measured: observables made of z / zz terms matrix-free (their eigenbasis is the
        computational basis and the eigenvalues are the diagonal H: the state is re-ordered by the ascending
        eigenvalues, `numpy.linalg.eigh` order; ties do not change which eigenvalue a draw selects);
        observables with x / y terms through the dense eigensystem like the reference (numpy.linalg.eigh on
        the host, one dense matrix-vector product per measurement on the device; n <= 12).  The draws are those
        scipy's `rvs` would make on the global numpy stream."""
        if getattr(self, '_eig_basis', None) is None:          # the reference's has_loaded_eigensystem
            self._eig_basis = self._measurement_basis(self.observable, lambda: self.observable.matrix)
            self.eigenvalues = self._eig_basis[1]
        return self._sample_grad_measured(self._eig_basis, shot_num, exact_expec_val, ini_state)

    # -- mc_clean.py:277-350: the same with ONE observable component, drawn per call ----------------
    def sample_grad_dense_with_component_sampling(self, shot_num=1, hide_progbar=True, exact_expec_val=True, ini_state=None):
        """Like sample_grad_dense, but every shifted circuit is measured in the eigenbasis of ONE component of the
        observable, drawn with np.random.choice(p=weight_distribution) before the circuit runs (mc_clean.py:301).
        Needs use_observable_components=True at construction, like the reference.  Returns the exact expectation
        value under the FULL observable."""
        obs = self.observable
        if not getattr(obs, 'store_components', False):
            raise AttributeError('construct the circuit with use_observable_components=True')
        if getattr(self, '_component_bases', None) is None:    # the reference's has_loaded_component_eigensystems
            self._component_bases = [None] * obs.num_components
        component = np.random.choice(np.arange(obs.num_components), p=obs.weight_distribution)
        if self._component_bases[component] is None:
            self._component_bases[component] = self._measurement_basis(obs.component(component), lambda: obs._host_matrix([component]))
        return self._sample_grad_measured(self._component_bases[component], shot_num, exact_expec_val, ini_state)

    def _measurement_basis(self, obs, host_matrix):
        """('perm', eigenvalues, order) for z / zz observables, ('dense', eigenvalues, V^dagger) otherwise.
        `host_matrix` is a callable: the 2^n x 2^n host matrix is only built on the dense branch (z / zz observables stay
        matrix-free at any register size)."""
        if np.any(obs.term_kinds < 2):
            if self.qnum > 12:
                raise NotImplementedError('observables with x / y terms are measured in the eigenbasis of the dense 2^n x 2^n '
                                          'observable (mc_clean.py:221), which is limited to 12 qubits here')
            eigenvalues, eigenvectors = np.linalg.eigh(host_matrix().toarray())         # mc_clean.py:221
            return ('dense', eigenvalues, np.ascontiguousarray(eigenvectors.transpose().conj()))   # mc_clean.py:222
        st = self.state
        st._load_ham(obs)
        ham = st._download_ham()
        order = np.argsort(ham, kind='stable')
        return ('perm', ham[order], order)

    def _sample_grad_measured(self, basis, shot_num, exact_expec_val, ini_state):
        kind, eigenvalues, table = basis
        axes, angles = self._params()
        n, L = self.qnum, self.lnum
        st = self.state
        if self.__dict__.get('_loaded_basis') is not table:    # the device holds one permutation / one dense matrix
            if kind == 'dense':
                st.load_dense(table)
            else:
                st.load_permutation(table)
            self._loaded_basis = table
        if ini_state is None:
            st.reset()
        else:
            st.vec = ini_state
        grad = np.ndarray([L, n], dtype='double')
        for q in range(n):
            st.yrot(np.pi / 4., q)
        for i in range(L):
            st.cnot_ladder(0)
            for q in range(n):
                self._rot(i, q)
            st.save(i)                                          # mc_clean.py:238 state_history[i]
        expec_val = self.expec_val() if exact_expec_val else self.sample_expec_val(shot_num)

        def measure(i):
            """layers i+1 .. L-1 on the current state, then shot_num draws of the eigenvalue"""
            self._forward_tail(i + 1)
            if kind == 'dense':
                st.apply_dense()                                # amplitudes in the eigenbasis (mc_clean.py:255)
            else:
                st.permute()
            return eigenvalues[self.sample_bitstrings(shot_num)].mean()

        for i in range(L):
            for q in range(n):
                st.load(i)
                self._manual_rot(i, q, np.pi / 2)
                sample1 = measure(i)
                st.load(i)                                      # the reference shifts the same vector on by -pi (mc_clean.py:262)
                self._manual_rot(i, q, np.pi / 2)
                self._manual_rot(i, q, -np.pi)
                sample2 = measure(i)
                grad[i, q] = (sample1 - sample2) / 2.
        st.free_snapshots()
        return expec_val, grad

    def _forward_tail(self, first_layer):
        """Layers first_layer .. L-1 applied to the CURRENT state in fused tile passes (one call, P launches per layer,
        instead of n + 1 gate-at-a-time launches per layer): the tail every shifted circuit of the sampled-gradient
        estimators runs (mc_clean.py:142-145, 148-151, 257-261)."""
        L = self.lnum
        if first_layer >= L:
            return
        axes, angles = self._params()
        k = L - first_layer
        ladder = np.ones(k, dtype=np.uint8)
        e = ctypes.c_double()
        self._lib.call('qr_layered_grad', self.state._ctx, int(k), _lib.ptr(np.ascontiguousarray(axes[first_layer:])),
                       _lib.ptr(np.ascontiguousarray(angles[first_layer:])), _lib.ptr(ladder), 1, self.observable._handle,
                       ctypes.byref(e), None)

    def _rot(self, i, q, angle_sign=1.):
        self._manual_rot(i, q, angle_sign * self.angles[i, q])

    def _manual_rot(self, i, q, angle):                 # mc_clean.py:394-403
        ax = self.axes[i, q]
        if ax == 0:
            self.state.xrot(angle, q)
        elif ax == 1:
            self.state.yrot(angle, q)
        elif ax == 2:
            self.state.zrot(angle, q)
        else:
            raise ValueError('Invalid axis {}'.format(ax))
