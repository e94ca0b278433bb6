"""QAOA ansatz (reference: circuit_logic/qaoa.py).

Circuit (qaoa.py:32-34): |+>^n, then p x [exp(-i gamma_i H), Rx(beta_i) on every qubit] with H
the diagonal Hamiltonian of the observable ('z' / 'zz' terms).  grad_run returns
(E, grad[p, 2]) with grad[:, 0] = dE/dbeta, grad[:, 1] = dE/dgamma (qaoa.py:62-68), computed by
the two-vector adjoint recurrence in fused CUDA tile passes.
"""
import ctypes
import warnings

import numpy as np

from .. import _lib
from ..physical_components import Gates
from .base import ParametrizedCircuit


class Qaoa(ParametrizedCircuit):
    def __init__(self, qubit_number, observable, layer_number, **kwargs):
        ParametrizedCircuit.init(self, qubit_number, observable, device=kwargs.get('device', 0))
        self.state.gates = Gates(qubit_number) \
            .add_xrots() \
            .add_x_summed() \
            .add_classical_ham(self.observable, include_individual_components=True)
        self.lnum = layer_number
        self.state.reset('+')   # initialize in uniform-superposition state (qaoa.py:17)

    def _check_parameters(self, betas, gammas):
        betas, gammas = np.asarray(betas), np.asarray(gammas)
        if (betas.size != self.lnum) or (gammas.size != self.lnum):   # qaoa.py:186-191
            raise ValueError('Wrong amount of parameters. Expected {0} and {0}, found {1} and {2}.'.format(
                self.lnum, betas.size, gammas.size))
        return (np.ascontiguousarray(betas, dtype=np.float64).ravel(),
                np.ascontiguousarray(gammas, dtype=np.float64).ravel())

    def _adopt(self, ini_state):
        if ini_state is None:
            return 0
        self.state.vec = ini_state      # qaoa.py:28
        return 1

    # -- qaoa.py:23-38 ------------------------------------------------------------------------
    def run_expec_val(self, betas, gammas, hide_progbar=True, exact_expec_val=True, shot_num=1, ini_state=None):
        '''Runs the circuit and returns the expectation value under observable'''
        use_current = self._adopt(ini_state)
        betas, gammas = self._check_parameters(betas, gammas)
        e = ctypes.c_double()
        self._lib.call('qr_qaoa_expec', self.state._ctx, self.lnum, _lib.ptr(betas), _lib.ptr(gammas), use_current,
                       ctypes.byref(e))
        if exact_expec_val:
            # qaoa.py:35-36 returns expec_val() under the FULL observable; qr_qaoa_expec reduces <psi|H|psi> over the z / zz
            # terms (all there is for MaxCut), so observables that also carry x / y terms take the general reduction
            if np.any(self.observable.term_kinds < 2):
                return self.expec_val()
            return e.value
        return self.sample_expec_val(shot_num)

    # -- qaoa.py:40-70 ------------------------------------------------------------------------
    def grad_run(self, betas, gammas, hide_progbar=True, ini_state=None):
        use_current = self._adopt(ini_state)
        betas, gammas = self._check_parameters(betas, gammas)
        e = ctypes.c_double()
        grad = np.empty([self.lnum, 2], dtype='double')
        self._lib.call('qr_qaoa_grad', self.state._ctx, self.lnum, _lib.ptr(betas), _lib.ptr(gammas), use_current,
                       ctypes.byref(e), _lib.ptr(grad))
        return e.value, grad

    # -- optimiser loop on the device (optimization.py:113-129 QaoaOpt.step, repeated) ----------
    def optimize_on_device(self, rule, hyper, iteration, steps, params, m=None, v=None, keep_param_history=True):
        """`steps` x (grad_run, parameter update) enqueued back to back on the device.

        params: float64[p, 2] = rows (beta_i, gamma_i), updated in place; rule / hyper / m / v as in
        `McClean.optimize_on_device`.  Needs an integer-valued Hamiltonian (MaxCut); raises ValueError otherwise.
        Returns (cost_history float64[steps], param_history float64[steps, p, 2] or None, iteration, hyper)."""
        par = np.ascontiguousarray(params, dtype=np.float64)
        if par.shape != (self.lnum, 2):
            raise ValueError('params must have shape ({}, 2)'.format(self.lnum))
        hyper = np.ascontiguousarray(hyper, dtype=np.float64)
        cost = np.zeros(steps, dtype=np.float64)
        hist = np.zeros((steps, self.lnum, 2), dtype=np.float64) if keep_param_history else None
        it = ctypes.c_int(int(iteration))
        mm = np.ascontiguousarray(m, dtype=np.float64) if m is not None else None
        vv = np.ascontiguousarray(v, dtype=np.float64) if v is not None else None
        self._lib.call('qr_qaoa_optimize', self.state._ctx, self.lnum, _lib.ptr(par), int(rule), _lib.ptr(hyper), ctypes.byref(it),
                       _lib.ptr(mm) if mm is not None else None, _lib.ptr(vv) if vv is not None else None, int(steps), _lib.ptr(cost),
                       _lib.ptr(hist) if hist is not None else None)
        params[...] = par
        if m is not None:
            m[...] = mm
            v[...] = vv
        return cost, hist, it.value, hyper

    # -- qaoa.py:72-81 ------------------------------------------------------------------------
    def sample_grad(self, betas, gammas, shot_num=1, hide_progbar=True, exact_expec_val=True, ini_state=None):
        warnings.warn('Not implemented yet.')

    # -- qaoa.py:83-158, matrix free ------------------------------------------------------------
    def sample_grad_dense(self, betas, gammas, shot_num=1, hide_progbar=True, exact_expec_val=True, ini_state=None):
        """Finite-shot parameter-shift gradient where every shifted circuit is measured in the
        computational basis (the eigenbasis of H).  The reference propagates dense 2^n x 2^n
        left-hand-side matrices (qaoa.py:95-134); here each shifted state is pushed through the
        remaining layers on the device and `shot_num` bitstrings are drawn from |psi|^2 by the
        prefix-sum sampler.  Same RNG consumption as the reference: one uniform(size=shot_num) per
        `__sample` call, in the same order (component +, component -, ..., qubit +, qubit -)."""
        st, n, p = self.state, self.qnum, self.lnum
        use_current = self._adopt(ini_state)
        betas, gammas = self._check_parameters(betas, gammas)
        if not use_current:
            st.reset()
        for i in range(p):                                   # qaoa.py:116-120
            st.exp_ham_classical(gammas[i])
            st.save(2 * i)
            for q in range(n):
                st.xrot(betas[i], q)
            st.save(2 * i + 1)
        expec_val = self.expec_val() if exact_expec_val else self.sample_expec_val(shot_num)   # :122-126
        grad = np.ndarray([p, 2], dtype='double')
        ncomp = st.gates.num_ham_components()

        def finish_and_sample(first_layer):
            """run layers first_layer..p-1 on the current state, then sample the cost"""
            if first_layer < p:
                e = ctypes.c_double()
                b, g = np.ascontiguousarray(betas[first_layer:]), np.ascontiguousarray(gammas[first_layer:])
                self._lib.call('qr_qaoa_expec', st._ctx, p - first_layer, _lib.ptr(b), _lib.ptr(g), 1, ctypes.byref(e))
            return self.sample_cost(shot_num)

        for i in range(p):
            plus, minus = np.empty(ncomp), np.empty(ncomp)
            for j in range(ncomp):                           # gamma_i : qaoa.py:139-147
                for sign, out in ((1., plus), (-1., minus)):
                    st.load(2 * i)
                    st.exp_ham_classical_component(sign * np.pi / 4, j)
                    for q in range(n):
                        st.xrot(betas[i], q)
                    out[j] = finish_and_sample(i + 1)
            grad[i, 1] = (plus - minus).sum()
            plus, minus = np.empty(n), np.empty(n)
            for q in range(n):                               # beta_i : qaoa.py:149-157
                for sign, out in ((1., plus), (-1., minus)):
                    st.load(2 * i + 1)
                    st.xrot(sign * np.pi / 2, q)
                    out[q] = finish_and_sample(i + 1)
            grad[i, 0] = .5 * (plus - minus).sum()
        st.free_snapshots()
        return expec_val, grad

    # -- qaoa.py:196-198 applied to the current state --------------------------------------------
    def sample_cost(self, shot_num, uniforms=None):
        """Mean of H over `shot_num` bitstrings drawn from |psi|^2 (the reference's private
        ``__sample(dist, shot_num)`` with dist = |state.vec|^2), entirely on the device."""
        idx = self.sample_bitstrings(shot_num, uniforms)
        vals = np.empty(idx.size, dtype=np.float64)
        self._lib.call('qr_ham_gather', self.state._ctx, int(idx.size), _lib.ptr(idx), _lib.ptr(vals))
        return vals.mean()
