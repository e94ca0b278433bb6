"""ParametrizedCircuit: shared plumbing of the circuits (reference: circuit_logic/base.py:9-58)."""
import ctypes

import numpy as np

from .. import _lib
from ..physical_components import State, Observable


def _inverse_cdf_draw(pk, uniforms):
    """Index rule of scipy.stats.rv_discrete(values=(xk, pk)).rvs: first k with cumsum(pk)[k] >= u,
    and 0 if no entry qualifies (scipy's argmax over an all-False row).  base.py:30-31."""
    cdf = np.cumsum(pk)
    idx = np.searchsorted(cdf, uniforms, side='left')
    idx[idx >= len(pk)] = 0
    return idx


class ParametrizedCircuit:
    '''Parent class for VQE circuits. Not meant for instantiation.'''

    def init(self, qubit_number, observable, use_observable_components=False, device=0):
        self.qnum = qubit_number
        self.state = State(qubit_number, device=device)   # gates are attached by the child class
        self.observable = Observable(qubit_number, observable, store_components=use_observable_components)
        self.has_loaded_projectors = False
        self._lib = self.state._lib

    # base.py:17-20: Re <psi|O|psi>, one fused reduction kernel
    def expec_val(self):
        out = ctypes.c_double()
        self._lib.call('qr_expec_val', self.state._ctx, self.observable._handle, ctypes.byref(out))
        return out.value

    def term_expectations(self):
        """<P_k> of every Pauli term of the observable (projector order), computed on the device."""
        out = np.empty(max(self.observable.num_components, 1), dtype=np.float64)
        self._lib.call('qr_term_expecs', self.state._ctx, self.observable._handle, _lib.ptr(out))
        return out[:self.observable.num_components]

    # base.py:22-33: per-term Bernoulli(+-w) estimate.  The probabilities come from the device;
    # the draws replay scipy's rvs on the global numpy stream (one uniform(size=shot_num) per term).
    def sample_expec_val(self, shot_num):
        if not self.has_loaded_projectors:
            self.observable.load_projectors()
            self.has_loaded_projectors = True
        expec = self.term_expectations()
        expec_val = 0.
        for k in range(self.observable.num_components):
            expec_val += self._bernoulli_mean(expec[k], self.observable.projector_weights[k], shot_num)
        return expec_val

    # base.py:35-46
    def sample_component_expec_val(self, shot_num, component):
        if not self.has_loaded_projectors:
            self.observable.load_projectors()
            self.has_loaded_projectors = True
        expec = self.term_expectations()
        return 0. + self._bernoulli_mean(expec[component], self.observable.projector_weights[component], shot_num)

    @staticmethod
    def _bernoulli_mean(term_expec, weight, shot_num):
        prob = .5 * (1. + term_expec)          # = sum |P_k psi|^2  (observable.py:126-177)
        draws = _inverse_cdf_draw(np.array([prob, 1. - prob]), np.random.uniform(size=shot_num))
        return np.array([weight, -weight])[draws].mean()

    def sample_bitstrings(self, shot_num, uniforms=None):
        """Indices drawn from |psi|^2 by inverse CDF on the device (qaoa.py:196-198).  `uniforms`
        defaults to np.random.uniform(size=shot_num), the draw scipy's rvs would make."""
        u = np.random.uniform(size=shot_num) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        out = np.empty(u.size, dtype=np.int64)
        self._lib.call('qr_sample_bitstrings', self.state._ctx, int(u.size), _lib.ptr(u), _lib.ptr(out))
        return out

    def perf(self):
        return self.state.perf()


def progbar_range(hide_progbar):
    '''base.py:51-58.  The device path runs a whole circuit in one call, so there is no per-layer
    host loop to decorate; the kwarg is accepted for compatibility.'''
    return np.arange
