"""MeynardClassifier: data re-uploading classifier circuit (reference: tutorials/meynard-classifier.ipynb).

The reference snapshot ships only the tutorial (cells 0, 3, 7-8, 11, 14), not the class, so the API below follows the
notebook and the circuit is this repo's documented definition -- PARITY UNPINNED against the reference:

    |0..0>
    encoding layer l    (l < encoding_layer_number):    [CNOT ladder(0) unless l == 0]
                        Rx(data[l, q])  Ry(encoding_angles[l, q, 0])  Rz(encoding_angles[l, q, 1])      on every qubit q
    classifying layer l (l < classifying_layer_number): [CNOT ladder(0) unless it is the very first layer of the circuit]
                        Rx(classifying_angles[l, q, 0])  Ry(..[l, q, 1])  Rz(..[l, q, 2])               on every qubit q
    observable: Z on qubit 0 (override with the `observable` kwarg)

("encoding_angles[0, 1, 0] is the angle for the y-rotation on the second qubit in the first layer; the x-rotations in
the encoding part are done with the data; classifying_angles[2, 0, 2] is the angle for the z-rotation", cell 8.)

Every layer is three "sub-layers" of one Pauli rotation per qubit; the whole circuit and its adjoint gradient sweep run
in the fused tile passes of the McClean engine through qr_layered_grad (ladder optional per sub-layer, no Ry(pi/4) layer).
"""
import ctypes

import numpy as np

from .. import _lib
from ..physical_components import Gates
from .base import ParametrizedCircuit


class MeynardClassifier(ParametrizedCircuit):
    def __init__(self, qubit_number, encoding_layer_number, classifying_layer_number, observable=None, **kwargs):
        if observable is None:
            observable = {'z': np.array([1.] + [None] * (qubit_number - 1), dtype=object)}
        ParametrizedCircuit.init(self, qubit_number, observable, False, device=kwargs.get('device', 0))
        self.elnum = encoding_layer_number
        self.clnum = classifying_layer_number
        self.state.gates = Gates(self.qnum).add_xrots().add_yrots().add_zrots().add_cnot_ladder()

    # -- the circuit as sub-layers: axes / angles [3 (Le + Lc), n] and the ladder flag of each sub-layer -----------
    def _sublayers(self, data, encoding_angles, classifying_angles):
        n, Le, Lc = self.qnum, self.elnum, self.clnum
        data = np.asarray(data, dtype=np.float64)
        enc = np.asarray(encoding_angles, dtype=np.float64)
        cls = np.asarray(classifying_angles, dtype=np.float64)
        if data.shape != (Le, n):
            raise ValueError('data must have shape ({}, {})'.format(Le, n))
        if enc.shape != (Le, n, 2):
            raise ValueError('encoding_angles must have shape ({}, {}, 2)'.format(Le, n))
        if cls.shape != (Lc, n, 3):
            raise ValueError('classifying_angles must have shape ({}, {}, 3)'.format(Lc, n))
        S = 3 * (Le + Lc)
        angles = np.empty((S, n), dtype=np.float64)
        angles[0:3 * Le:3] = data
        angles[1:3 * Le:3] = enc[:, :, 0]
        angles[2:3 * Le:3] = enc[:, :, 1]
        for a in range(3):
            angles[3 * Le + a::3] = cls[:, :, a]
        axes = np.ascontiguousarray(np.broadcast_to(np.tile(np.arange(3, dtype=np.int32), Le + Lc)[:, None], (S, n)))
        ladder = np.zeros(S, dtype=np.uint8)
        ladder[3::3] = 1                      # every layer but the first starts with the ladder
        return axes, np.ascontiguousarray(angles), ladder

    def run(self, data, encoding_angles, classifying_angles):
        '''Runs the circuit; state.vec holds the final state afterwards (notebook cell 11).'''
        axes, angles, ladder = self._sublayers(data, encoding_angles, classifying_angles)
        e = ctypes.c_double()
        self._lib.call('qr_layered_grad', self.state._ctx, int(axes.shape[0]), _lib.ptr(axes), _lib.ptr(angles), _lib.ptr(ladder),
                       0, self.observable._handle, ctypes.byref(e), None)
        return e.value

    def run_expec_val(self, data, encoding_angles, classifying_angles):
        return self.run(data, encoding_angles, classifying_angles)

    def grad_run(self, data, encoding_angles, classifying_angles):
        '''Returns (expectation value, encoding_angles gradient [Le, n, 2], classifying_angles gradient [Lc, n, 3])
        (notebook cell 14).  No run() is required before calling it.'''
        axes, angles, ladder = self._sublayers(data, encoding_angles, classifying_angles)
        Le, Lc, n = self.elnum, self.clnum, self.qnum
        e = ctypes.c_double()
        grad = np.empty(angles.shape, dtype=np.float64)
        self._lib.call('qr_layered_grad', self.state._ctx, int(axes.shape[0]), _lib.ptr(axes), _lib.ptr(angles), _lib.ptr(ladder),
                       0, self.observable._handle, ctypes.byref(e), _lib.ptr(grad))
        enc_grad = np.empty((Le, n, 2), dtype=np.float64)
        enc_grad[:, :, 0] = grad[1:3 * Le:3]
        enc_grad[:, :, 1] = grad[2:3 * Le:3]
        cls_grad = np.empty((Lc, n, 3), dtype=np.float64)
        for a in range(3):
            cls_grad[:, :, a] = grad[3 * Le + a::3]
        return e.value, enc_grad, cls_grad
