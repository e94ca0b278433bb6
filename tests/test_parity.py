"""Parity of the CUDA path (through the Python drop-in classes -> ctypes -> C ABI) with the golden
vectors produced by the reference and with the CPU oracle.

Tolerance (the reference's own cross-check is np.allclose(..., atol=1e-10),
qradient-vs-pennylane-benchmark.py:129-130):
    |E - E_ref| <= 1e-10 * scale,   allclose(grad, grad_ref, rtol=1e-10, atol=1e-10 * scale),
    scale = sum_k |w_k| of the observable.  Sampling: integer equality of the drawn indices.
"""
import numpy as np
import pytest

from backends import backend  # noqa: F401
from conftest import load_golden, obs_from_golden, obs_scale, assert_parity
from oracle import qr_oracle as orc
from qradient_b200.circuit_logic import McClean, Qaoa
from qradient_b200.optimization_problems import MaxCut
from qradient_b200.physical_components import State, Gates, Observable

TOL = 1e-10
MCCLEAN = ["gv1_mcclean_3x3", "gv2_mcclean_4x2_mixed", "gv2b_mcclean_4x2_ini", "gv5_mcclean_12x6",
           "gv6_mcclean_7x5_mixed"]


def zz01(n):
    m = np.full((n, n), None)
    m[0, 1] = 1.0
    return {"zz": m}


def mixed_obs(n):
    obs = zz01(n)
    obs["x"] = np.array([0.3] + [None] * (n - 1), dtype=object)
    obs["y"] = np.array([None] * (n - 1) + [0.7], dtype=object)
    obs["z"] = np.array([None, -0.4] + [None] * (n - 2), dtype=object)
    obs["zz"][1, n - 1] = 0.5
    return obs


@pytest.mark.parametrize("name", MCCLEAN)
def test_mcclean_golden(backend, name):
    d = load_golden(name)
    n, L, obs = int(d["n"]), int(d["L"]), obs_from_golden(d)
    ini = d["ini_state"] if "ini_state" in d else None
    c = McClean(n, obs, L, axes=d["axes"], angles=d["angles"])
    e_run = c.run_expec_val(ini_state=None if ini is None else ini.copy())
    assert abs(e_run - float(d["E_run"])) <= TOL * obs_scale(obs)
    np.testing.assert_allclose(c.state.vec, d["psi_final"], atol=1e-12)          # state.vec = psi_final
    e, g = c.grad_run(ini_state=None if ini is None else ini.copy())
    assert g.shape == (L, n) and g.dtype == np.float64
    assert_parity(e, g, float(d["E"]), d["grad"], obs_scale(obs), TOL)
    np.testing.assert_allclose(c.state.vec, d["costate"], atol=1e-12 * obs_scale(obs))   # mc_clean.py:77
    assert abs(c.expec_val() - np.vdot(d["costate"], orc.apply_observable(orc.OracleObservable(n, obs), d["costate"])).real) < 1e-10


def test_mcclean_readme_example(backend):
    """README.md:30-37 with the global RNG seeded: constructor draws axes then angles."""
    d = load_golden("gv1_mcclean_3x3")
    np.random.seed(0)
    interactions = np.full((3, 3), None)
    interactions[0, 1] = 1.
    circuit = McClean(3, {'zz': interactions}, 3)
    np.testing.assert_array_equal(circuit.axes, d["axes"])
    np.testing.assert_array_equal(circuit.angles, d["angles"])
    e, g = circuit.grad_run()
    assert_parity(e, g, float(d["E"]), d["grad"], 1.0, TOL)


@pytest.mark.parametrize("name", ["gv3_qaoa_4x2", "gv4_qaoa_12x3"])
def test_qaoa_golden(backend, name):
    d = load_golden(name)
    n, p = int(d["n"]), int(d["p"])
    q = Qaoa(n, MaxCut(n, edge_set=d["edges"]).to_observable(), p)
    scale = float(len(d["edges"]))
    e, g = q.grad_run(d["betas"], d["gammas"])
    assert g.shape == (p, 2)
    assert_parity(e, g, float(d["E"]), d["grad"], scale, TOL)
    e_run = q.run_expec_val(d["betas"], d["gammas"])
    assert abs(e_run - float(d["E_run"])) <= TOL * scale
    np.testing.assert_allclose(q.state.vec, d["psi_final"], atol=1e-12)
    np.testing.assert_array_equal(q.state.gates.classical_ham, d["ham"])
    # bitstring sampling: integer equality on the same uniform stream (qaoa.py:196-198)
    idx = q.sample_bitstrings(len(d["uniforms"]), d["uniforms"])
    np.testing.assert_array_equal(idx, d["bitstrings"])
    assert abs(q.sample_cost(len(d["uniforms"]), d["uniforms"]) - float(d["bit_mean"])) < 1e-12
    # global-stream variant and the per-term Bernoulli estimate (base.py:22-33)
    np.random.seed(0)
    np.testing.assert_array_equal(q.sample_bitstrings(len(d["uniforms"])), d["bitstrings"])
    np.random.seed(0)
    assert abs(q.run_expec_val(d["betas"], d["gammas"], exact_expec_val=False, shot_num=100) - float(d["E_shots100_seed"])) < 1e-12


def test_mcclean_sample_expec_val_golden(backend):
    d = load_golden("gv9_mcclean_10x4_shots")
    n, obs = int(d["n"]), obs_from_golden(d)
    c = McClean(n, obs, int(d["L"]), axes=d["axes"], angles=d["angles"])
    assert abs(c.run_expec_val() - float(d["E"])) < TOL * obs_scale(obs)
    np.random.seed(3)
    assert abs(c.run_expec_val(exact_expec_val=False, shot_num=50) - float(d["E_shots50_seed3"])) < 1e-12
    te = c.term_expectations()
    psi = c.state.vec
    np.testing.assert_allclose(te, orc.term_expectations(orc.OracleObservable(n, obs), np.array(psi)), atol=1e-12)


def test_state_single_gates_golden(backend):
    d = load_golden("gv7_gates_5")
    n = 5
    st = State(n)
    st.gates = Gates(n).add_xrots().add_yrots().add_zrots().add_cnot_ladder().add_x_summed()
    for name in ("xrot", "yrot", "zrot", "dxrot", "dyrot", "dzrot"):
        for q in range(n):
            st.vec = d["v0"]
            getattr(st, name)(0.37 + 0.11 * q, q)
            np.testing.assert_allclose(st.vec, d["%s_q%d" % (name, q)], atol=1e-15)
    for s in (0, 1):
        st.vec = d["v0"]
        st.cnot_ladder(s)
        np.testing.assert_array_equal(st.vec, d["ladder%d" % s])
    st.vec = d["v0"]
    st.x_summed()
    np.testing.assert_allclose(st.vec, d["x_summed"], atol=1e-15)
    for key in d.files:
        if key.startswith("cnot_"):
            _, c, t = key.split("_")
            st.vec = d["v0"]
            st.cnot(int(c), int(t))
            np.testing.assert_array_equal(st.vec, d[key])


def test_ladder_maps_golden(backend):
    """Ladder permutation for n = 2..10 against the reference's own ladder matrices."""
    d = load_golden("gv8_ladder_maps")
    for n in range(2, 11):
        st = State(n)
        st.gates = Gates(n).add_cnot_ladder()
        for s in (0, 1):
            st.vec = np.arange(2 ** n).astype(complex)
            st.cnot_ladder(s)
            dest = np.empty(2 ** n, dtype=np.int64)
            dest[np.asarray(st.vec).real.astype(np.int64)] = np.arange(2 ** n)
            np.testing.assert_array_equal(dest, d["n%d_s%d" % (n, s)])


@pytest.mark.parametrize("n,L,tile_bits", [(4, 2, 12), (5, 3, 4), (9, 3, 12), (10, 2, 5), (11, 2, 4), (12, 2, 6),
                                            (12, 2, 7), (10, 2, 8), (13, 2, 12), (14, 1, 9)])
def test_mcclean_tile_geometries_vs_oracle(backend, n, L, tile_bits):
    """Every pass geometry the planner can produce (1..8 passes per layer, 1..3 rounds per pass)."""
    rng = np.random.default_rng(100 * n + tile_bits)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    e_ref, g_ref, co = orc.mcclean_grad_run(n, obs, axes, angles, return_state=True)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    c.state.set_option("tile_bits", tile_bits)
    e, g = c.grad_run()
    assert_parity(e, g, e_ref, g_ref, obs_scale(obs), TOL)
    np.testing.assert_allclose(c.state.vec, co, atol=1e-12 * obs_scale(obs))
    assert abs(c.run_expec_val() - e_ref) <= TOL * obs_scale(obs)
    assert abs(c.state.norm_error()) < 1e-12


@pytest.mark.parametrize("opts", [dict(prefetch=1), dict(prefetch=5), dict(staged=3), dict(staged=1), dict(staged=0, prefetch=0),
                                  dict(tile_bits_strided=5, min_row_bits=2), dict(tile_bits_strided=4, min_row_bits=1),
                                  dict(defer_reduce=0), dict(pdl=2)])
@pytest.mark.parametrize("n,L,tile_bits", [(7, 2, 5), (10, 2, 12), (13, 1, 12), (9, 2, 4), (12, 2, 12), (14, 2, 11), (15, 1, 12)])
def test_kernel_variants_vs_oracle(backend, opts, n, L, tile_bits):
    """Loads (direct, L2 prefetch, staged by asynchronous copies), strided tile sizes, reduction and launch modes."""
    rng = np.random.default_rng(7 * n + tile_bits)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    c.state.set_option("tile_bits", tile_bits)
    for k, v in opts.items():
        c.state.set_option(k, v)
    e, g = c.grad_run()
    assert_parity(e, g, e_ref, g_ref, obs_scale(obs), TOL)
    q = Qaoa(n, MaxCut(n, edge_set=[(i, i + 1) for i in range(n - 1)]).to_observable(), 2)
    q.state.set_option("tile_bits", tile_bits)
    for k, v in opts.items():
        q.state.set_option(k, v)
    b, gm = rng.random(2), rng.random(2)
    e_ref, g_ref = orc.qaoa_grad_run(n, orc.maxcut_observable(n, [(i, i + 1) for i in range(n - 1)]), b, gm)
    e, g = q.grad_run(b, gm)
    assert_parity(e, g, e_ref, g_ref, float(n - 1), TOL)


@pytest.mark.parametrize("n,L", [(12, 3), (14, 2), (16, 2), (17, 1), (19, 1)])
def test_lean_tile_kernel_group_counts(backend, n, L):
    """k_tile12 (qr_tile12.cuh): strided passes with 1, 2 and 3 register groups next to the 4-group first
    pass, against the generic tile kernel (10-bit tiles) and, where it finishes in seconds, the oracle."""
    rng = np.random.default_rng(1200 + n)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    c.state.set_option("tile_bits", 12)
    c.state.set_option("axis_plan", 0)   # the static plan (axis-aware plans: test_axis_aware_plans)
    e1, g1 = c.grad_run()
    v1 = np.array(c.state.vec)
    c.state.set_option("tile_bits", 10)
    e0, g0 = c.grad_run()
    assert_parity(e1, g1, e0, g0, obs_scale(obs), 1e-12)
    np.testing.assert_allclose(v1, c.state.vec, atol=1e-13)
    if n <= 14:
        e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
        assert_parity(e1, g1, e_ref, g_ref, obs_scale(obs), TOL)
    c.state.set_option("tile_bits", 12)
    np.testing.assert_allclose(c.run_expec_val(), e0, atol=1e-12 * obs_scale(obs))


@pytest.mark.parametrize("n,L,min_row_bits", [(11, 3, 3), (12, 2, 3), (14, 2, 3), (17, 2, 3), (19, 1, 3), (13, 2, 2), (20, 1, 2)])
def test_lean_tile_kernel_half_size_tiles(backend, n, L, min_row_bits):
    """k_tile12 with K = 11 (2048-amplitude tiles, 256 threads): every chain of register groups, incl. the
    64 B-row chain 8 | 2 | 5, against the 12-bit tiles (and the oracle where it finishes in seconds)."""
    rng = np.random.default_rng(1100 + n)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    c.state.set_option("tile_bits", 12)
    e0, g0 = c.grad_run()
    v0 = np.array(c.state.vec)
    c.state.set_option("tile_bits", 11)
    c.state.set_option("min_row_bits", min_row_bits)
    e1, g1 = c.grad_run()
    assert_parity(e1, g1, e0, g0, obs_scale(obs), 1e-12)
    np.testing.assert_allclose(v0, c.state.vec, atol=1e-13)
    np.testing.assert_allclose(c.run_expec_val(), e0, atol=1e-12 * obs_scale(obs))
    if n <= 14:
        e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
        assert_parity(e1, g1, e_ref, g_ref, obs_scale(obs), TOL)
    if n <= 17:
        q = Qaoa(n, MaxCut(n, edge_set=[(i, i + 1) for i in range(n - 1)]).to_observable(), 2)
        b, gm = rng.random(2), rng.random(2)
        q.state.set_option("tile_bits", 12)
        e0, g0 = q.grad_run(b, gm)
        q.state.set_option("tile_bits", 11)
        q.state.set_option("min_row_bits", min_row_bits)
        e1, g1 = q.grad_run(b, gm)
        assert_parity(e1, g1, e0, g0, float(n - 1), 1e-12)


def _axis_case(case, n, L, rng):
    """Axes (L, n) of the named pattern; index bit b <-> qubit n-1-b."""
    axes = rng.integers(0, 3, (L, n))
    if case == "few_xy":            # one or two X / Y rotations above the contiguous tile: tiles padded with fillers
        axes[:, :n - 12] = 2
        axes[:, 0] = 0
        axes[1:, 3] = 1
    elif case == "all_z_high":      # nothing but Rz above the contiguous tile
        axes[:, :n - 12] = 2
    elif case == "all_xy":          # no Rz at all
        axes = rng.integers(0, 2, (L, n))
    elif case == "three_rounds":    # at most 4 X / Y gates on bits 5..11: the contiguous pass runs the chain L | 0 | 3
        for b in range(5, 12):
            axes[:, n - 1 - b] = 2
        axes[0, n - 1 - 5], axes[0, n - 1 - 9] = 0, 1
        axes[1:, n - 1 - 6], axes[1:, n - 1 - 7], axes[1:, n - 1 - 10], axes[1:, n - 1 - 11] = 1, 0, 0, 1
    elif case == "absorb":          # 7 X / Y bits above the tile, Rz on bits 7..11: the contiguous pass trades bits for them
        for b in range(n):
            axes[:, n - 1 - b] = (b % 2) if b >= 12 else (2 if b >= 7 else b % 3)
        axes[1:, n - 1 - 9] = 1
    return axes


@pytest.mark.parametrize("n,L,tile_bits,case", [(15, 2, 12, "random"), (17, 2, 11, "random"), (17, 1, 12, "few_xy"),
                                                 (16, 2, 12, "all_z_high"), (16, 1, 12, "all_xy"), (19, 1, 12, "absorb"), (18, 1, 11, "absorb"), (16, 2, 12, "three_rounds")])
def test_axis_aware_plans(backend, n, L, tile_bits, case):
    """QR_OPT_AXIS_PLAN: per-layer plans from the axes (general tile geometry, Rz gates of index bits outside the tile
    applied through the tile's own index bits, split barriers, the contiguous pass on a general tile with the ladder
    gather) against the static plan and the oracle."""
    rng = np.random.default_rng(31 * n + L)
    axes = _axis_case(case, n, L, rng)
    angles = rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    c.state.set_option("tile_bits", tile_bits)
    c.state.set_option("axis_plan", 0)
    e0, g0 = c.grad_run()
    v0 = np.array(c.state.vec)
    if n <= 17:
        e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
        assert_parity(e0, g0, e_ref, g_ref, obs_scale(obs), TOL)
    for mode in ((1, 15) if case in ("random", "absorb", "three_rounds") else (15,)):   # 1: block barriers, no trade with the contiguous pass
        c.state.set_option("axis_plan", mode)
        e1, g1 = c.grad_run()
        assert_parity(e1, g1, e0, g0, obs_scale(obs), 1e-12)
        np.testing.assert_allclose(c.state.vec, v0, atol=1e-13)
        np.testing.assert_allclose(c.run_expec_val(), e0, atol=1e-12 * obs_scale(obs))


@pytest.mark.parametrize("case", ["all_x", "all_y", "all_z", "special_angles"])
def test_lean_tile_kernel_special_gates(backend, case):
    """tan-form edge cases: cos = 0 or sin = 0 exactly, |cos| = |sin|, negative factors, passes without / with only Rz."""
    n, L = 13, 2
    rng = np.random.default_rng(77)
    angles = rng.uniform(0, 2 * np.pi, (L, n))
    axes = rng.integers(0, 3, (L, n))
    if case == "all_x":
        axes[:] = 0
    elif case == "all_y":
        axes[:] = 1
    elif case == "all_z":
        axes[:] = 2
    else:
        special = np.array([0.0, np.pi, 2 * np.pi, np.pi / 2, 3 * np.pi / 2, -np.pi / 2, 4 * np.pi, 1e-9, np.pi - 1e-9])
        angles = special[rng.integers(0, len(special), (L, n))]
    obs = mixed_obs(n)
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    e, g = c.grad_run()
    assert_parity(e, g, e_ref, g_ref, obs_scale(obs), TOL)
    np.testing.assert_allclose(c.run_expec_val(), e_ref, atol=TOL * obs_scale(obs))


def test_fused_equals_unfused(backend):
    n, L = 11, 3
    rng = np.random.default_rng(5)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = McClean(n, mixed_obs(n), L, axes=axes, angles=angles)
    e1, g1 = c.grad_run()
    v1 = np.array(c.state.vec)
    c.state.set_option("fusion", 0)
    e0, g0 = c.grad_run()
    assert_parity(e1, g1, e0, g0, obs_scale(mixed_obs(n)), 1e-12)
    np.testing.assert_allclose(v1, c.state.vec, atol=1e-13)
    q = Qaoa(n, MaxCut(n, edge_set=[(min(i, (i + 1) % n), max(i, (i + 1) % n)) for i in range(n)]).to_observable(), 2)
    b, gm = rng.random(2), rng.random(2)
    e1, g1 = q.grad_run(b, gm)
    q.state.set_option("fusion", 0)
    e0, g0 = q.grad_run(b, gm)
    assert_parity(e1, g1, e0, g0, float(n), 1e-12)


def test_qaoa_weighted_and_z_terms_vs_oracle(backend):
    n, p = 9, 3
    rng = np.random.default_rng(42)
    zz = np.full((n, n), None)
    for (a, b) in [(0, 1), (1, 5), (2, 8), (3, 4), (0, 8), (6, 7)]:
        zz[a, b] = float(rng.normal())
    obs = {"z": np.array([0.25, None, -0.5] + [None] * (n - 3), dtype=object), "zz": zz}
    betas, gammas = rng.random(p), rng.random(p)
    e_ref, g_ref = orc.qaoa_grad_run(n, obs, betas, gammas)
    for tb in (12, 5):
        q = Qaoa(n, obs, p)
        q.state.set_option("tile_bits", tb)
        e, g = q.grad_run(betas, gammas)
        assert_parity(e, g, e_ref, g_ref, obs_scale(obs), TOL)
    # ini_state is adopted (qaoa.py:28)
    ini = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    ini /= np.linalg.norm(ini)
    e_ref, g_ref = orc.qaoa_grad_run(n, obs, betas, gammas, ini_state=ini)
    e, g = q.grad_run(betas, gammas, ini_state=ini)
    assert_parity(e, g, e_ref, g_ref, obs_scale(obs), TOL)
    # and a later call without ini_state returns to |+> (qaoa.py:17,26)
    e2, _ = q.grad_run(betas, gammas)
    assert abs(e2 - orc.qaoa_grad_run(n, obs, betas, gammas)[0]) <= TOL * obs_scale(obs)


def test_state_api_vs_oracle(backend):
    n = 6
    rng = np.random.default_rng(3)
    v0 = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    v0 /= np.linalg.norm(v0)
    obs_d = mixed_obs(n)
    obs = Observable(n, obs_d)
    oobs = orc.OracleObservable(n, obs_d)
    st = State(n)
    assert st.vec[0] == 1.0 and np.count_nonzero(st.vec) == 1
    st.reset('+')
    np.testing.assert_allclose(st.vec, 2.0 ** (-0.5 * n) * np.ones(2 ** n), atol=1e-16)
    st.reset()
    np.testing.assert_allclose(st.vec, 2.0 ** (-0.5 * n) * np.ones(2 ** n), atol=1e-16)   # reset() keeps '+'
    st.vec = v0
    assert abs(st.norm_error()) < 1e-14
    st.multiply_matrix(obs.matrix)
    np.testing.assert_allclose(st.vec, orc.apply_observable(oobs, v0), atol=1e-14)
    st.vec = v0
    st.multiply_matrix(obs)
    np.testing.assert_allclose(st.vec, obs.matrix.dot(v0), atol=1e-14)
    # any other host matrix (the method's contract is vec = M.dot(vec)): dense and scipy.sparse, small registers
    import scipy.sparse as sp
    rng_m = np.random.default_rng(5)
    dense = rng_m.normal(size=(2 ** n, 2 ** n)) + 1j * rng_m.normal(size=(2 ** n, 2 ** n))
    for m in (dense, sp.csr_matrix(dense * (np.abs(dense) > 1.0))):
        st.vec = v0
        st.multiply_matrix(m)
        np.testing.assert_allclose(st.vec, m.dot(v0), atol=1e-11)
    with pytest.raises(ValueError):
        st.multiply_matrix(np.eye(3))
    # classical Hamiltonian family (z / zz terms only; x and y are ignored with a warning)
    with pytest.warns(UserWarning):
        st.gates = Gates(n).add_classical_ham(obs, include_individual_components=True)
    cl = orc.OracleObservable(n, {k: v for k, v in obs_d.items() if k in ("z", "zz")})
    ham = orc.classical_ham_vector(cl)
    np.testing.assert_allclose(st.gates.classical_ham, ham, atol=1e-15)
    comps = st.gates.classical_ham_components
    np.testing.assert_allclose(comps.sum(axis=0), ham, atol=1e-14)
    st.vec = v0
    st.exp_ham_classical(0.37)
    np.testing.assert_allclose(st.vec, v0 * np.exp(-0.37j * ham), atol=1e-14)
    st.vec = v0
    st.ham_classical()
    np.testing.assert_allclose(st.vec, v0 * (-1j * ham), atol=1e-14)
    for k in range(comps.shape[0]):
        st.vec = v0
        st.exp_ham_classical_component(0.81, k)
        np.testing.assert_allclose(st.vec, v0 * np.exp(-0.81j * comps[k]), atol=1e-14)
    # item assignment on the host view writes back (mc_clean.py:76 idiom)
    st.vec[:] = v0[::-1]
    np.testing.assert_array_equal(st.vec, v0[::-1])
    st.vec *= 2.0
    np.testing.assert_array_equal(st.vec, 2.0 * v0[::-1])
    # periodic ladder (state.py:209-241) for even n
    st2 = State(n)
    st2.gates = Gates(n).add_cnot_ladder(periodic=True)
    for s in (0, 1):
        st2.vec = v0
        st2.cnot_ladder(s)
        o = orc.OracleState(n); o.vec = v0.copy(); o.cnot_ladder(s, periodic=True)
        np.testing.assert_array_equal(st2.vec, o.vec)
    with pytest.warns(UserWarning, match="Not implemented"):
        st.xrot_lhs(0.1, 0)


def test_error_behaviour(backend):
    with pytest.raises(ValueError):
        State(3, ini='x')                                   # state.py:71
    st = State(3)
    with pytest.raises(ValueError):
        st.cnot(1, 1)                                       # state.py:355
    with pytest.raises(ValueError):
        st.cnot(0, 3)
    with pytest.raises(ValueError):
        st.xrot(0.1, 3)
    with pytest.raises(ValueError):
        st.vec = np.zeros(4)
    with pytest.raises(ValueError):
        Gates(3).add_cnot_ladder(periodic=True)             # state.py:211-215
    zz = np.full((3, 3), None)
    zz[1, 0] = 1.0
    with pytest.raises(ValueError):
        Observable(3, {"zz": zz})                           # observable.py:57-64
    with pytest.raises(ValueError):
        Observable(3, {"x": np.array([1.0, None], dtype=object)})
    with pytest.warns(UserWarning):
        Observable(2, {"x": np.array([1.0, None], dtype=object), "foo": np.zeros(2)})   # observable.py:115
    c = McClean(3, zz01(3), 2, axes=np.array([[0, 1, 3], [0, 0, 0]]), angles=np.zeros((2, 3)))
    with pytest.raises(ValueError):
        c.grad_run()                                        # mc_clean.py:392
    with pytest.raises(ValueError):
        c.run_expec_val()
    q = Qaoa(3, zz01(3), 2)
    with pytest.raises(ValueError):
        q.grad_run(np.zeros(3), np.zeros(2))                # qaoa.py:186-191
    with pytest.raises(ValueError):
        q.run_expec_val(np.zeros(2), np.zeros(1))


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_empty_and_tiny_inputs(backend, n):
    """Edge cases on both backends: zero layers, one- and two-qubit registers (below the fused path, which starts at
    4 qubits), observables without terms."""
    _empty_and_tiny_inputs(n)


def _empty_and_tiny_inputs(n):
    rng = np.random.default_rng(40 + n)
    obs = {"z": np.array([0.7] + [None] * (n - 1), dtype=object), "x": np.array([None] * (n - 1) + [0.4], dtype=object)}
    # zero layers: E of Ry(pi/4)^n |0>, empty gradient
    c0 = McClean(n, obs, 0, axes=np.zeros((0, n), dtype=int), angles=np.zeros((0, n)))
    e, g = c0.grad_run()
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, np.zeros((0, n), dtype=int), np.zeros((0, n)))
    assert g.shape == (0, n) and abs(e - e_ref) < 1e-12 and abs(c0.run_expec_val() - e_ref) < 1e-12
    # a few layers on tiny registers
    L = 3
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = McClean(n, obs, L, axes=axes, angles=angles)
    e, g = c.grad_run()
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
    assert_parity(e, g, e_ref, g_ref, 1.1, TOL)
    # an observable without any term: E = 0, zero gradient
    empty = {"z": np.array([None] * n, dtype=object)}
    ce = McClean(n, empty, L, axes=axes, angles=angles)
    e, g = ce.grad_run()
    assert e == 0.0 and not np.any(g)
    if n >= 2:
        zz = np.full((n, n), None)
        zz[0, n - 1] = 1.0
        q = Qaoa(n, {"zz": zz}, 0)                      # zero QAOA layers: |+>^n, <ZZ> = 0
        e, g = q.grad_run(np.zeros(0), np.zeros(0))
        assert abs(e) < 1e-15 and g.shape == (0, 2)
        q1 = Qaoa(n, {"zz": zz}, 1)
        b, gm = rng.random(1), rng.random(1)
        e, g = q1.grad_run(b, gm)
        e_ref, g_ref = orc.qaoa_grad_run(n, {"zz": zz}, b, gm)
        assert_parity(e, g, e_ref, g_ref, 1.0, TOL)


def test_optimizer_style_parameter_updates(backend):
    """optimization.py:61,91 assign circuit.angles between grad_run calls."""
    n, L = 6, 2
    rng = np.random.default_rng(8)
    axes = rng.integers(0, 3, (L, n))
    c = McClean(n, zz01(n), L, axes=axes, angles=rng.uniform(0, 2 * np.pi, (L, n)))
    for _ in range(3):
        e, g = c.grad_run()
        e_ref, g_ref = orc.mcclean_grad_run(n, zz01(n), axes, c.angles)
        assert_parity(e, g, e_ref, g_ref, 1.0, TOL)
        c.angles = c.angles - 0.1 * g


def test_grad_run_batch(backend):
    n, L, B = 8, 3, 5
    rng = np.random.default_rng(4)
    axes, angles = rng.integers(0, 3, (B, L, n)), rng.uniform(0, 2 * np.pi, (B, L, n))
    obs = mixed_obs(n)
    c = McClean(n, obs, L, axes=axes[0], angles=angles[0])
    for tb in (12, 5):
        c.state.set_option("tile_bits", tb)
        e, g = c.grad_run_batch(angles, axes)
        assert e.shape == (B,) and g.shape == (B, L, n)
        for b in range(B):
            e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes[b], angles[b])
            assert_parity(e[b], g[b], e_ref, g_ref, obs_scale(obs), TOL)
    # shared axes
    e2, g2 = c.grad_run_batch(angles, axes[0])
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes[0], angles[1])
    assert_parity(e2[1], g2[1], e_ref, g_ref, obs_scale(obs), TOL)
    # the single-circuit path still works on the same object afterwards
    e, g = c.grad_run()
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes[0], angles[0])
    assert_parity(e, g, e_ref, g_ref, obs_scale(obs), TOL)


def test_sample_grad_parameter_shift(backend):
    """mc_clean.py:117-156 on the device State API, same global-RNG draw order as the oracle."""
    n, L = 4, 2
    rng = np.random.default_rng(6)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    np.random.seed(11)
    e, g = c.sample_grad(shot_num=20)
    np.random.seed(11)
    e_ref, g_ref = orc.mcclean_sample_grad(n, obs, axes, angles, 20)
    assert abs(e - e_ref) < 1e-12
    np.testing.assert_allclose(g, g_ref, atol=1e-12)


def test_sampling_edge_cases(backend):
    n = 13   # two scan chunks
    rng = np.random.default_rng(2)
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    v[:100] = 0.0        # leading zeros: ties at cdf = 0
    v /= np.linalg.norm(v)
    q = Qaoa(n, zz01(n), 1)
    q.state.vec = v
    u = np.concatenate([rng.uniform(size=200), [0.0, 1e-300, 0.5, 1.0 - 1e-16]])
    idx = q.sample_bitstrings(u.size, u)
    ref = orc.sample_bitstrings(v, u)
    cdf = np.cumsum(np.abs(v) ** 2)
    for a, b, uu in zip(idx, ref, u):
        if a != b:   # only allowed when u sits within rounding of a cdf step
            lo, hi = min(a, b), max(a, b)
            assert abs(cdf[lo] - uu) < 1e-13 and abs(cdf[hi - 1] - uu) < 1e-13, (a, b, uu)
    assert np.mean(idx == ref) > 0.99


def test_qaoa_sample_grad_dense_golden(backend):
    """qaoa.py:83-158 through the matrix-free device path; same global-RNG draw order."""
    d = load_golden("gv11_qaoa_sample_grad_dense")
    n, p = int(d["n"]), int(d["p"])
    q = Qaoa(n, MaxCut(n, edge_set=d["edges"]).to_observable(), p)
    np.random.seed(int(d["seed"]))
    e, g = q.sample_grad_dense(d["betas"], d["gammas"], shot_num=int(d["shot_num"]))
    assert abs(e - float(d["E"])) < 1e-10 * len(d["edges"])
    np.testing.assert_allclose(g, d["grad"], atol=1e-12)
    with pytest.warns(UserWarning, match="Not implemented"):
        q.sample_grad(d["betas"], d["gammas"])


def test_mcclean_component_sampling_golden(backend):
    """mc_clean.py:80-115: np.random.choice over observable components, then the adjoint sweep."""
    d = load_golden("gv12_mcclean_component_sampling")
    n, L, obs = int(d["n"]), int(d["L"]), obs_from_golden(d)
    c = McClean(n, obs, L, use_observable_components=True, axes=d["axes"], angles=d["angles"])
    np.random.seed(int(d["seed"]))
    e, g = c.grad_run_with_component_sampling()
    assert_parity(e, g, float(d["E"]), d["grad"], obs_scale(obs), TOL)
    c2 = McClean(n, obs, L, axes=d["axes"], angles=d["angles"])
    with pytest.raises(AttributeError):
        c2.grad_run_with_component_sampling()


def test_state_snapshots(backend):
    n = 6
    rng = np.random.default_rng(1)
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    st = State(n)
    st.vec = v
    st.save(3)
    st.xrot(0.3, 2)
    st.save(0)
    w = np.array(st.vec)
    st.load(3)
    np.testing.assert_array_equal(st.vec, v)
    st.load(0)
    np.testing.assert_array_equal(st.vec, w)
    with pytest.raises(ValueError):
        st.load(2)
    st.free_snapshots()
    with pytest.raises(ValueError):
        st.load(0)


@pytest.mark.parametrize("n,L", [(12, 3), (14, 2)])
def test_deferred_reduction_matches_per_pass_reduction(backend, n, L):
    """QR_OPT_DEFER_REDUCE: one reduction launch per gradient (k_reduce_slots_strided) against the last-CTA reduction
    fused into every backward pass -- same partials, same order of additions."""
    rng = np.random.default_rng(900 + n)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    e1, g1 = c.grad_run()
    c.state.set_option("defer_reduce", 0)
    e0, g0 = c.grad_run()
    assert e0 == e1
    np.testing.assert_allclose(g1, g0, rtol=0, atol=1e-15 * obs_scale(obs))
    edges = [(i, i + 1) for i in range(n - 1)]
    q = Qaoa(n, MaxCut(n, edge_set=edges).to_observable(), 2)
    b, gm = rng.random(2), rng.random(2)
    e1, g1 = q.grad_run(b, gm)
    q.state.set_option("defer_reduce", 0)
    e0, g0 = q.grad_run(b, gm)
    assert e0 == e1
    np.testing.assert_allclose(g1, g0, rtol=0, atol=1e-15 * (n - 1))


def test_projector_dot_and_dense_tracking_attributes(backend):
    """API parity rows outside the hot path: Projector.dot (observable.py:126-177) against the Kronecker constructions of
    the reference, and State.activate_lefthandside / activate_center_matrix / set_center_matrix (state.py:39-59)."""
    import scipy.sparse as sp
    n = 4
    x = np.array([[.5, .5], [.5, .5]], dtype=complex)
    y = np.array([[.5, -.5j], [.5j, .5]], dtype=complex)
    up, dn = np.diag([1., 0.]).astype(complex), np.diag([0., 1.]).astype(complex)

    def kron_at(ops):
        m = np.eye(1, dtype=complex)
        for q in range(n):
            m = np.kron(m, ops.get(q, np.eye(2, dtype=complex)))
        return m

    zz = np.full((n, n), None)
    zz[1, 3] = 0.6
    obs = Observable(n, {"x": np.array([0.5, None, None, None], dtype=object), "y": np.array([None, None, 0.7, None], dtype=object),
                         "z": np.array([None, 0.8, None, None], dtype=object), "zz": zz})
    obs.load_projectors()
    rng = np.random.default_rng(5)
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    expected = [kron_at({0: x}), kron_at({2: y}), kron_at({1: up}), kron_at({1: up, 3: up}) + kron_at({1: dn, 3: dn})]
    assert len(obs.projectors) == 4
    for proj, m in zip(obs.projectors, expected):
        np.testing.assert_allclose(proj.dot(v), m.dot(v), atol=1e-14)
    st = State(n)
    with pytest.raises(AttributeError):
        st.set_center_matrix(sp.identity(2 ** n, format='csr'))
    st.activate_lefthandside()
    st.activate_center_matrix()
    assert st.lhs.shape == (2 ** n, 2 ** n) and st.lhs.nnz == 2 ** n and st.center_matrix.nnz == 0
    st.set_center_matrix(sp.identity(2 ** n, dtype=complex, format='csr'))
    assert st.center_matrix.nnz == 2 ** n
    with pytest.warns(UserWarning):
        st.xrot_lhs(0.1, 0)                      # still the reference's 'Not implemented.' stub


def test_options_and_permutation_api_validation(backend):
    """Argument checks of the kernel-selection options and of qr_perm_load / qr_state_permute."""
    n = 6
    st = State(n)
    for name, bad in (("tile_bits", 3), ("tile_bits", 13), ("prefetch", 32), ("staged", 16), ("pdl", 3), ("min_row_bits", 12), ("axis_plan", 16), ("axis_plan", -1)):
        with pytest.raises(ValueError):
            st.set_option(name, bad)
    with pytest.raises(ValueError):                   # retired round-1 experiment key (decoupled exchange)
        st._lib.call("qr_set_option", st._ctx, 14, 1)
    st.set_option("tile_bits", 0)                     # auto
    st.set_option("loop_graph", 0)
    st.set_option("axis_plan", 15)
    with pytest.raises(ValueError):
        st.load_permutation(np.arange(2 ** n - 1))    # wrong length
    with pytest.raises(ValueError):
        st.load_permutation(np.zeros(2 ** n, dtype=np.int64))   # not a permutation
    fresh = State(n)
    with pytest.raises(RuntimeError):
        fresh.permute()                                # nothing loaded
    rng = np.random.default_rng(3)
    v = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    perm = rng.permutation(2 ** n)
    st.vec = v
    st.load_permutation(perm)
    st.permute()
    np.testing.assert_array_equal(np.array(st.vec), v[perm])
