"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"` tests run in the build container (no GPU): oracle vs golden vectors, host logic,
C-ABI library loading.  `-m gpu` tests are the parity tests proper and call the CUDA path.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def obs_from_golden(d):
    """Rebuild the reference-style observable dict (None = absent) from a golden file (NaN = None)."""
    n = int(d["n"])
    obs = {}
    for k in ("x", "y", "z"):
        a = d["obs_" + k]
        if not np.all(np.isnan(a)):
            obs[k] = np.array([None if np.isnan(v) else float(v) for v in a], dtype=object)
    zz = d["obs_zz"]
    if not np.all(np.isnan(zz)):
        m = np.full((n, n), None)
        for i in range(n):
            for j in range(n):
                if not np.isnan(zz[i, j]):
                    m[i, j] = float(zz[i, j])
        obs["zz"] = m
    return obs


def obs_scale(obs):
    s = 0.0
    for v in obs.values():
        for w in np.asarray(v, dtype=object).ravel():
            if w is not None:
                s += abs(w)
    return s if s > 0 else 1.0


def assert_parity(e, g, e_ref, g_ref, scale, tol=1e-10):
    """SURVEY.md 8(c) criterion: |dE| <= tol*scale and allclose(g, g_ref, rtol=tol, atol=tol*scale)."""
    assert abs(e - e_ref) <= tol * scale, (e, e_ref)
    np.testing.assert_allclose(g, g_ref, rtol=tol, atol=tol * scale)


@pytest.fixture(scope="session")
def golden():
    return load_golden
