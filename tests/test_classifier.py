"""MeynardClassifier (tutorials/meynard-classifier.ipynb): API of the notebook, circuit as documented in
qradient_b200/circuit_logic/meynard_classifier.py (parity unpinned: the reference snapshot has no source for it).
Checked against the oracle's gate-by-gate forward pass and exact parameter-shift gradients."""
import numpy as np
import pytest

from backends import _NO_GPU, backend, activate  # noqa: F401
from oracle import qr_oracle as orc
from qradient_b200.circuit_logic import MeynardClassifier


def _inputs(n, Le, Lc, seed):
    rng = np.random.default_rng(seed)
    return rng.random((Le, n)), rng.random((Le, n, 2)), rng.random((Lc, n, 3))


@pytest.mark.parametrize("n,Le,Lc", [(3, 2, 2), (4, 2, 1), (5, 1, 2), (6, 2, 2), (5, 0, 2), (5, 2, 0), (13, 1, 1)])
def test_classifier_vs_oracle(backend, n, Le, Lc):
    data, enc, cls = _inputs(n, Le, Lc, 100 * n + 10 * Le + Lc)
    c = MeynardClassifier(n, Le, Lc)
    e_ref, psi_ref = orc.classifier_run(n, data, enc, cls, return_state=True)
    c.run(data, enc, cls)
    assert abs(c.expec_val() - e_ref) < 1e-12
    np.testing.assert_allclose(c.state.vec, psi_ref, atol=1e-13)
    e, ge, gc = c.grad_run(data, enc, cls)
    assert ge.shape == (Le, n, 2) and gc.shape == (Lc, n, 3)
    assert abs(e - e_ref) < 1e-12
    if n <= 6:
        _, ge_ref, gc_ref = orc.classifier_grad_run(n, data, enc, cls)
        np.testing.assert_allclose(ge, ge_ref, rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(gc, gc_ref, rtol=1e-10, atol=1e-10)
    else:                                   # exact parameter shift on a few entries
        for idx in ((0, 0, 1), (0, n - 1, 0), (0, 5, 1)):
            for arr, grad in ((enc, ge), (cls, gc)):
                keep = arr[idx]
                arr[idx] = keep + np.pi / 2
                ep = orc.classifier_run(n, data, enc, cls)
                arr[idx] = keep - np.pi / 2
                em = orc.classifier_run(n, data, enc, cls)
                arr[idx] = keep
                assert abs(0.5 * (ep - em) - grad[idx]) < 1e-10


def test_classifier_custom_observable_and_errors(backend):
    n, Le, Lc = 5, 2, 2
    data, enc, cls = _inputs(n, Le, Lc, 7)
    zz = np.full((n, n), None)
    zz[1, 3] = 0.5
    obs = {"zz": zz, "x": np.array([None, 0.25] + [None] * (n - 2), dtype=object)}
    c = MeynardClassifier(n, Le, Lc, observable=obs)
    e, ge, gc = c.grad_run(data, enc, cls)
    e_ref, ge_ref, gc_ref = orc.classifier_grad_run(n, data, enc, cls, observable=obs)
    assert abs(e - e_ref) < 1e-12
    np.testing.assert_allclose(ge, ge_ref, rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(gc, gc_ref, rtol=1e-10, atol=1e-10)
    with pytest.raises(ValueError):
        c.grad_run(data[:1], enc, cls)
    with pytest.raises(ValueError):
        c.run(data, enc[:, :, :1], cls)
    with pytest.raises(ValueError):
        c.grad_run(data, enc, cls[:, :, :2])


@pytest.mark.gpu
@_NO_GPU
def test_classifier_notebook_sizes_gpu():
    """Notebook cell 18 sweeps 2..14 qubits at 5 + 5 layers; here 14 and 20 qubits with finite differences."""
    activate("cuda")
    for n in (14, 20):
        data, enc, cls = _inputs(n, 5, 5, n)
        c = MeynardClassifier(n, 5, 5)
        e, ge, gc = c.grad_run(data, enc, cls)
        assert abs(c.run(data, enc, cls) - e) < 1e-12
        assert abs(c.state.norm_error()) < 1e-12
        eps = 1e-5
        for arr, grad, idx in ((enc, ge, (2, 3, 1)), (cls, gc, (4, 0, 0)), (cls, gc, (0, n - 1, 2)), (enc, ge, (0, 0, 0))):
            keep = arr[idx]
            arr[idx] = keep + eps
            ep = c.run(data, enc, cls)
            arr[idx] = keep - eps
            em = c.run(data, enc, cls)
            arr[idx] = keep
            assert abs((ep - em) / (2 * eps) - grad[idx]) < 1e-7
