"""The oracle (oracle/qr_oracle.py) against the golden vectors produced by the reference's own
circuit_logic files (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import load_golden, obs_from_golden, obs_scale, assert_parity
from oracle import qr_oracle as orc

MCCLEAN = ["gv1_mcclean_3x3", "gv2_mcclean_4x2_mixed", "gv2b_mcclean_4x2_ini", "gv5_mcclean_12x6",
           "gv6_mcclean_7x5_mixed"]
QAOA = ["gv3_qaoa_4x2", "gv4_qaoa_12x3"]


@pytest.mark.parametrize("name", MCCLEAN)
def test_mcclean_grad_matches_reference(name):
    d = load_golden(name)
    n, obs = int(d["n"]), obs_from_golden(d)
    ini = d["ini_state"] if "ini_state" in d else None
    e, g, co = orc.mcclean_grad_run(n, obs, d["axes"], d["angles"], ini_state=ini, return_state=True)
    assert_parity(e, g, float(d["E"]), d["grad"], obs_scale(obs), tol=1e-13)
    np.testing.assert_allclose(co, d["costate"], atol=1e-13 * obs_scale(obs))
    e2, psi = orc.mcclean_run_expec_val(n, obs, d["axes"], d["angles"], ini_state=ini, return_state=True)
    assert abs(e2 - float(d["E_run"])) < 1e-13 * obs_scale(obs)
    np.testing.assert_allclose(psi, d["psi_final"], atol=1e-13)


@pytest.mark.parametrize("name", QAOA)
def test_qaoa_grad_matches_reference(name):
    d = load_golden(name)
    n = int(d["n"])
    obs = orc.maxcut_observable(n, d["edges"])
    e, g, psi = orc.qaoa_grad_run(n, obs, d["betas"], d["gammas"], return_state=True)
    assert_parity(e, g, float(d["E"]), d["grad"], len(d["edges"]), tol=1e-13)
    np.testing.assert_allclose(psi, d["psi_final"], atol=1e-13)
    np.testing.assert_array_equal(orc.classical_ham_vector(orc.OracleObservable(n, obs)), d["ham"])
    assert abs(orc.qaoa_run_expec_val(n, obs, d["betas"], d["gammas"]) - float(d["E_run"])) < 1e-12


@pytest.mark.parametrize("name", QAOA)
def test_sampling_matches_reference(name):
    d = load_golden(name)
    n = int(d["n"])
    idx = orc.sample_bitstrings(d["psi_final"], d["uniforms"])
    np.testing.assert_array_equal(idx, d["bitstrings"])
    assert abs(d["ham"][idx].mean() - float(d["bit_mean"])) < 1e-12
    # per-term Bernoulli estimate, global numpy stream (base.py:22-33)
    obs = orc.OracleObservable(n, orc.maxcut_observable(n, d["edges"]))
    np.random.seed(0)
    est = orc.sample_expec_val(obs, d["psi_final"], 100)
    assert abs(est - float(d["E_shots100_seed"])) < 1e-12


def test_mcclean_sample_expec_val_matches_reference():
    d = load_golden("gv9_mcclean_10x4_shots")
    n, obs = int(d["n"]), obs_from_golden(d)
    e, psi = orc.mcclean_run_expec_val(n, obs, d["axes"], d["angles"], return_state=True)
    assert abs(e - float(d["E"])) < 1e-13
    np.random.seed(3)
    est = orc.sample_expec_val(orc.OracleObservable(n, obs), psi, 50)
    assert abs(est - float(d["E_shots50_seed3"])) < 1e-12


def test_single_gates_match_reference():
    d = load_golden("gv7_gates_5")
    n = 5
    for name in ("xrot", "yrot", "zrot", "dxrot", "dyrot", "dzrot"):
        for q in range(n):
            st = orc.OracleState(n)
            st.vec = d["v0"].copy()
            getattr(st, name)(0.37 + 0.11 * q, q)
            np.testing.assert_allclose(st.vec, d["%s_q%d" % (name, q)], atol=1e-15)
    for s in (0, 1):
        st = orc.OracleState(n); st.vec = d["v0"].copy(); st.cnot_ladder(s)
        np.testing.assert_array_equal(st.vec, d["ladder%d" % s])
    st = orc.OracleState(n); st.vec = d["v0"].copy(); st.x_summed()
    np.testing.assert_allclose(st.vec, d["x_summed"], atol=1e-15)
    for key in d.files:
        if key.startswith("cnot_"):
            _, c, t = key.split("_")
            st = orc.OracleState(n); st.vec = d["v0"].copy(); st.cnot(int(c), int(t))
            np.testing.assert_array_equal(st.vec, d[key])


def test_ladder_maps_match_reference_matrices():
    d = load_golden("gv8_ladder_maps")
    for n in range(2, 11):
        for s in (0, 1):
            ref = d["n%d_s%d" % (n, s)]
            np.testing.assert_array_equal(orc.ladder_scatter_map(n, s), ref)
            np.testing.assert_array_equal(orc.ladder_permutation(n, s), ref)
        # stacking 1 is the inverse of stacking 0 (state.py:235-241)
        a, b = orc.ladder_scatter_map(n, 0), orc.ladder_scatter_map(n, 1)
        np.testing.assert_array_equal(b[a], np.arange(2 ** n))


def test_finite_difference_agrees_with_oracle_gradient():
    d = load_golden("gv6_mcclean_7x5_mixed")
    n, obs = int(d["n"]), obs_from_golden(d)
    f = lambda a: orc.mcclean_run_expec_val(n, obs, d["axes"], a)
    fd = orc.finite_difference_grad(f, d["angles"], eps=1e-6)
    np.testing.assert_allclose(fd, d["grad"], atol=5e-9)


def test_observable_validation():
    zz = np.full((3, 3), None); zz[1, 0] = 1.0
    with pytest.raises(ValueError):
        orc.OracleObservable(3, {"zz": zz})
    with pytest.raises(ValueError):
        orc.OracleState(3, "x")
    st = orc.OracleState(3)
    with pytest.raises(ValueError):
        st.rot(3, 0.1, 0)
    with pytest.raises(ValueError):
        orc.qaoa_grad_run(3, orc.maxcut_observable(3, [(0, 1)]), np.zeros(2), np.zeros(3))
