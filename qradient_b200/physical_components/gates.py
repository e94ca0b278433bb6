"""Gates: the builder object the circuits attach to ``state.gates``.

The reference's ``Gates`` class is missing from its HEAD snapshot (imported at
circuit_logic/mc_clean.py:2,16-20 and qaoa.py:2,11-14, defined nowhere); its surface is
reconstructed from those call sites.  In the reference every ``add_*``
call builds 2^n x 2^n sparse generators.  On the device no matrices exist -- rotations,
ladders and diagonal phases are kernels -- so ``add_*`` only records which gate families are
enabled; ``add_classical_ham`` remembers the observable so that ``State`` can build the
diagonal table H[j] on the GPU when the gates are attached.
"""
import warnings

import numpy as np


class Gates:
    def __init__(self, qubit_number):
        self.qnum = int(qubit_number)
        self.has_xrots = self.has_yrots = self.has_zrots = False
        self.has_cnot_ladder = False
        self.ladder_periodic = False
        self.cnots_which = None
        self.has_x_summed = False
        self.ham_observable = None
        self.include_individual_components = False
        self._state = None
        self._ham_host = None

    # chainable builders (mc_clean.py:16-20, qaoa.py:11-14)
    def add_xrots(self):
        self.has_xrots = True
        return self

    def add_yrots(self):
        self.has_yrots = True
        return self

    def add_zrots(self):
        self.has_zrots = True
        return self

    def add_cnots(self, which):
        which = np.asarray(which, dtype=bool)
        if which.shape != (self.qnum, self.qnum) or np.any(np.diag(which)):
            raise ValueError('which must be a {0} by {0} boolean matrix with a False diagonal.'.format(self.qnum))
        self.cnots_which = which
        return self

    def add_cnot_ladder(self, periodic=False):
        if periodic and (self.qnum % 2 != 0):   # state.py:211-215
            raise ValueError('CNOT gates in a ladder structure with periodic boundaries '
                             'are ambiguous for uneven number of qubits.')
        self.has_cnot_ladder = True
        self.ladder_periodic = bool(periodic)
        return self

    def add_x_summed(self):
        self.has_x_summed = True
        return self

    def add_classical_ham(self, observable, include_individual_components=False):
        # state.py:268-271: only 'z' and 'zz' enter the classical Hamiltonian
        observable.check_observable(
            known_keys=['z', 'zz'],
            warning='Non-classical observable component found. Only \'z\' and \'zz\' are accepted in this method.')
        self.ham_observable = observable
        self.include_individual_components = bool(include_individual_components)
        self._ham_host = None
        if self._state is not None:
            self._state._load_ham(observable)
        return self

    # host views (users read circuit.state.gates.classical_ham, tutorials/qaoa-max-cut.ipynb cell 10)
    @property
    def classical_ham(self):
        if self.ham_observable is None:
            raise AttributeError('classical_ham: call add_classical_ham first')
        if self._ham_host is None:
            if self._state is None:
                raise AttributeError('classical_ham lives on the device: attach the gates to a State first')
            self._ham_host = self._state._download_ham()
        return self._ham_host

    @property
    def classical_ham_components(self):
        """float64[K, 2^n] host array of the individual diagonal terms (small registers only)."""
        if self.ham_observable is None or not self.include_individual_components:
            raise AttributeError('classical_ham_components: call add_classical_ham(obs, True) first')
        obs, n = self.ham_observable, self.qnum
        if n > 20:
            raise MemoryError('classical_ham_components is a host-side inspection aid for <= 20 qubits')
        idx = np.arange(2**n)
        rows = []
        for k in range(obs.num_components):
            kind = int(obs.term_kinds[k])
            if kind < 2:
                continue
            sign = 1.0 - 2.0 * ((idx >> (n - 1 - int(obs.term_qi[k]))) & 1)
            if kind == 3:
                sign = sign * (1.0 - 2.0 * ((idx >> (n - 1 - int(obs.term_qj[k]))) & 1))
            rows.append(obs.term_weights[k] * sign)
        return np.array(rows)

    def num_ham_components(self):
        return int(np.sum(self.ham_observable.term_kinds >= 2)) if self.ham_observable is not None else 0

    def __getattr__(self, name):
        if name in ('xrot', 'yrot', 'zrot_pos', 'zrot_neg', 'cnot_ladder', 'x_summed', 'cnots'):
            raise AttributeError(
                'Gates.{}: the device path holds no 2^n x 2^n generator matrices (the reference builds them at '
                'state.py:81-88,133-140,159-166,209-241); apply gates through State methods instead.'.format(name))
        raise AttributeError(name)
