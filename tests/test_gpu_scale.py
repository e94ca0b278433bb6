"""GPU-only parity at BASELINE.json sizes: golden config 2, and size-independent properties
(fused == gate-at-a-time, norm preservation, finite differences, E consistency) where the
oracle would take too long."""
import numpy as np
import pytest

from backends import activate, _NO_GPU
from conftest import load_golden, obs_from_golden, assert_parity
from oracle import qr_oracle as orc

pytestmark = [pytest.mark.gpu, _NO_GPU]


@pytest.fixture(autouse=True)
def _cuda():
    activate("cuda")
    yield


def zz01(n):
    m = np.full((n, n), None)
    m[0, 1] = 1.0
    return {"zz": m}


def test_config2_mcclean_20x20_golden():
    """BASELINE config 2 against the reference's own output (tests/golden/gv10)."""
    from qradient_b200.circuit_logic import McClean
    d = load_golden("gv10_mcclean_20x20")
    c = McClean(20, obs_from_golden(d), 20, axes=d["axes"], angles=d["angles"])
    e, g = c.grad_run()
    assert_parity(e, g, float(d["E"]), d["grad"], 1.0, 1e-10)
    assert abs(c.run_expec_val() - float(d["E"])) < 1e-10
    assert abs(c.state.norm_error()) < 1e-12
    for opt in (("prefetch", 1), ("tile_bits", 10), ("ctas_per_sm_fwd", 1), ("tile_bits", 12), ("staged", 3), ("staged", 0),
                ("prefetch", 5), ("defer_reduce", 0), ("pdl", 2), ("tile_bits", 11)):
        c.state.set_option(*opt)       # options accumulate: every kernel variant is exercised
        e2, g2 = c.grad_run()
        assert_parity(e2, g2, float(d["E"]), d["grad"], 1.0, 1e-10)


def test_mcclean_16x8_mixed_vs_oracle():
    from qradient_b200.circuit_logic import McClean
    n, L = 16, 8
    rng = np.random.default_rng(16)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = zz01(n)
    obs["x"] = np.array([0.3] + [None] * (n - 1), dtype=object)
    obs["y"] = np.array([None] * (n - 1) + [0.7], dtype=object)
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
    c = McClean(n, obs, L, axes=axes, angles=angles)
    e, g = c.grad_run()
    assert_parity(e, g, e_ref, g_ref, 2.0, 1e-10)


def test_fused_equals_gate_at_a_time_24_qubits():
    from qradient_b200.circuit_logic import McClean
    n, L = 24, 3
    rng = np.random.default_rng(24)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = McClean(n, zz01(n), L, axes=axes, angles=angles)
    e1, g1 = c.grad_run()
    assert c.perf()["passes_per_layer"] == 3
    c.state.set_option("fusion", 0)
    e0, g0 = c.grad_run()
    assert_parity(e1, g1, e0, g0, 1.0, 1e-11)


def test_qaoa_20_qubits_vs_oracle_and_sampling():
    from qradient_b200.circuit_logic import Qaoa
    from qradient_b200.optimization_problems import MaxCut
    n, p = 20, 2
    edges = MaxCut.random_regular(n, 3, seed=20)
    rng = np.random.default_rng(10)
    gammas, betas = rng.random(p), rng.random(p)
    e_ref, g_ref, psi = orc.qaoa_grad_run(n, orc.maxcut_observable(n, edges), betas, gammas, return_state=True)
    q = Qaoa(n, MaxCut(n, edge_set=edges).to_observable(), p)
    e, g = q.grad_run(betas, gammas)
    assert_parity(e, g, e_ref, g_ref, float(len(edges)), 1e-10)
    q.run_expec_val(betas, gammas)
    u = np.random.RandomState(0).uniform(size=100)
    idx = q.sample_bitstrings(100, u)
    ref = orc.sample_bitstrings(psi, u)
    assert np.mean(idx == ref) >= 0.99     # ties within rounding of a cdf step are the only allowed difference
    cdf = np.cumsum(np.abs(psi) ** 2)
    for a, b, uu in zip(idx, ref, u):
        if a != b:
            assert abs(cdf[min(a, b)] - uu) < 1e-12


def test_qaoa_config3_26_qubits_vs_full_size_oracle():
    """BASELINE config 3 at its stated size (QAOA MaxCut, 26 qubits, p = 10, the committed 39-edge graph, default_rng(10))
    against the full-size oracle fixture tests/golden/gv18 (tests/golden/make_golden_big.py: 21 history vectors of 1 GiB,
    24 minutes on one host core): E and grad[10, 2] at 1e-10 * sum|w| = 3.9e-9, and INTEGER EQUALITY of the 100 bitstring
    indices drawn with RandomState(0).uniform(size=100) (the fixture records that no uniform lies within 3e-11 of a cdf
    step, six orders of magnitude above the rounding of a parallel scan)."""
    from qradient_b200.circuit_logic import Qaoa
    from qradient_b200.optimization_problems import MaxCut
    import bench
    d = load_golden("gv18_qaoa_config3_26x10")
    n, p = int(d["n"]), int(d["p"])
    rng = np.random.default_rng(10)
    gammas, betas = rng.random(p), rng.random(p)
    assert np.array_equal(gammas, d["gammas"]) and np.array_equal(betas, d["betas"])
    assert np.array_equal(np.array(bench.CONFIG3_EDGES), d["edges"])
    q = Qaoa(n, MaxCut(n, edge_set=bench.CONFIG3_EDGES).to_observable(), p)
    scale = float(len(bench.CONFIG3_EDGES))
    e, g = q.grad_run(betas, gammas)
    assert_parity(e, g, float(d["e"]), d["grad"], scale, 1e-10)
    q.state.set_option("fusion", 0)                      # gate-at-a-time kernels on the same device
    e0, g0 = q.grad_run(betas, gammas)
    q.state.set_option("fusion", 1)
    assert_parity(e0, g0, float(d["e"]), d["grad"], scale, 1e-10)
    assert abs(q.run_expec_val(betas, gammas) - float(d["e"])) <= 1e-10 * scale
    assert abs(q.state.norm_error()) < 1e-12
    assert float(d["cdf_margin"].min()) > 1e-12
    idx = q.sample_bitstrings(100, d["uniforms"])
    assert np.array_equal(idx, d["idx"])
    assert abs(q.sample_cost(100, d["uniforms"]) - float(d["mean_cost"])) < 1e-12


def test_batched_config4_14x14_vs_oracle_fixture():
    """BASELINE config 4 at its stated size: all 8192 parameter sets of McClean 14 x 14 (default_rng(4)) in one
    grad_run_batch call; 8 of them (first, last, both sides of the chunk boundary at 4096) against the oracle fixture gv19."""
    from qradient_b200.circuit_logic import McClean
    d = load_golden("gv19_mcclean_config4_14x14")
    n, L, B = int(d["n"]), int(d["L"]), int(d["B"])
    rng = np.random.default_rng(int(d["seed"]))
    axes, angles = rng.integers(0, 3, (B, L, n)), rng.uniform(0, 2 * np.pi, (B, L, n))
    c = McClean(n, zz01(n), L, axes=axes[0], angles=angles[0])
    e, g = c.grad_run_batch(angles, axes)
    assert e.shape == (B,) and g.shape == (B, L, n)
    for k, b in enumerate(d["indices"]):
        assert_parity(e[b], g[b], float(d["e"][k]), d["grad"][k], 1.0, 1e-10)
    # every parameter set of the batch against its own single-circuit run would take too long: spot-check 4 more
    for b in (777, 2048, 5000, 6001):
        c.axes, c.angles = axes[b], angles[b]
        e1, g1 = c.grad_run()
        assert_parity(e[b], g[b], e1, g1, 1.0, 1e-12)


@pytest.mark.parametrize("G", [2, 4, 8])
def test_sharded_24_qubits_vs_oracle_fixture(G):
    """Sharded engine well above the sizes of the CPU tier: McClean 24 qubits x 6 layers over G virtual shards on one
    GPU (swap engine: exchange passes + cross-shard ladder tiles), observable ZZ(0,1) + 0.5 X_2 + 0.25 Y_23 (x term on a
    rank-held qubit for G = 8), against the oracle fixture gv20; and against the one-GPU path."""
    from qradient_b200.circuit_logic import McClean
    from qradient_b200.sharded import ShardedMcClean, LocalComm
    d = load_golden("gv20_mcclean_24x6")
    n, L = int(d["n"]), int(d["L"])
    zz = np.full((n, n), None)
    zz[0, 1] = 1.0
    x = np.array([None] * n, dtype=object)
    x[2] = 0.5
    y = np.array([None] * n, dtype=object)
    y[n - 1] = 0.25
    obs = {"zz": zz, "x": x, "y": y}
    sh = ShardedMcClean(n, obs, L, LocalComm(G), d["axes"], d["angles"])
    try:
        assert sh.mode == "swap"
        e, g = sh.grad_run()
        assert_parity(e, g, float(d["e"]), d["grad"], 1.75, 1e-10)
        if G == 2:   # the last local pass and the exchange pass issued in 8 slices of the index bits 9..11 (QR_OPT_SHARD_SLICES)
            sh.set_option("shard_slices", 8)
            launches = sh.perf["kernel_launches"]
            e8, g8 = sh.grad_run()
            assert sh.perf["kernel_launches"] > launches
            assert_parity(e8, g8, e, g, 1.75, 1e-12)
    finally:
        sh.close()
    if G == 2:
        one = McClean(n, obs, L, axes=d["axes"], angles=d["angles"])
        e1, g1 = one.grad_run()
        assert_parity(e1, g1, float(d["e"]), d["grad"], 1.75, 1e-10)


@pytest.mark.parametrize("G", [2, 8])
def test_sharded_qaoa_24_qubits_matches_one_gpu(G):
    """Sharded QAOA (swap engine without a ladder, per-layout H tables) against the one-GPU Qaoa on 24 qubits: E, gradient and
    the sampled bitstring indices."""
    from qradient_b200.circuit_logic import Qaoa
    from qradient_b200.optimization_problems import MaxCut
    from qradient_b200.sharded import ShardedQaoa, LocalComm
    n, p = 24, 4
    edges = MaxCut.random_regular(n, 3, seed=3)
    obs = MaxCut(n, edge_set=edges).to_observable()
    rng = np.random.default_rng(24)
    betas, gammas = rng.random(p), rng.random(p)
    one = Qaoa(n, obs, p)
    e1, g1 = one.grad_run(betas, gammas)
    one.run_expec_val(betas, gammas)
    u = np.random.RandomState(1).uniform(size=100)
    idx1 = one.sample_bitstrings(100, u)
    q = ShardedQaoa(n, obs, p, LocalComm(G))
    try:
        e, g = q.grad_run(betas, gammas)
        assert_parity(e, g, e1, g1, float(len(edges)), 1e-10)
        assert abs(q.run_expec_val(betas, gammas) - e1) < 1e-10 * len(edges)
        idx = q.sample_bitstrings(100, u)
        assert np.mean(idx == idx1) >= 0.99
    finally:
        q.close()


def test_two_devices_in_one_process():
    """One process driving two GPUs (LocalComm(devices=[0, 1]) and the `device=` keyword): the dynamic shared-memory
    opt-in of the tile kernels is per device (it used to be remembered per process).  Needs two visible devices."""
    import ctypes
    from qradient_b200 import _lib
    from qradient_b200.circuit_logic import McClean
    from qradient_b200.sharded import ShardedMcClean, LocalComm
    ndev = ctypes.c_int(0)
    _lib.lib().call("qr_device_count", ctypes.byref(ndev))
    if ndev.value < 2:
        pytest.skip("needs two CUDA devices")
    n, L = 16, 3
    rng = np.random.default_rng(2)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    e_ref, g_ref = orc.mcclean_grad_run(n, zz01(n), axes, angles)
    for dev in (1, 0):      # device 1 FIRST: it must get its own opt-in
        c = McClean(n, zz01(n), L, axes=axes, angles=angles, device=dev)
        e, g = c.grad_run()
        assert_parity(e, g, e_ref, g_ref, 1.0, 1e-10)
    sh = ShardedMcClean(n, zz01(n), L, LocalComm(2, devices=[0, 1]), axes, angles)
    try:
        e, g = sh.grad_run()
        assert_parity(e, g, e_ref, g_ref, 1.0, 1e-10)
    finally:
        sh.close()


def test_mcclean_30_qubits_properties():
    """North-star size (16 GiB state).  No oracle can run here (31 history vectors of 16 GiB), so
    this is a property test by necessity: E consistency, unit norm, finite differences on two angles; bench.py compares
    the sharded engine with this path at 30 x 30 in every multi-GPU run."""
    from qradient_b200.circuit_logic import McClean
    n, L = 30, 2
    rng = np.random.default_rng(30)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = McClean(n, zz01(n), L, axes=axes, angles=angles)
    e, g = c.grad_run()
    assert c.perf()["passes_per_layer"] == 3
    assert abs(c.run_expec_val() - e) < 1e-10
    assert abs(c.state.norm_error()) < 1e-11
    eps = 1e-5
    for (i, q) in ((0, 0), (1, 17)):
        a = angles.copy()
        a[i, q] += eps
        c.angles = a
        ep = c.run_expec_val()
        a[i, q] -= 2 * eps
        c.angles = a
        em = c.run_expec_val()
        assert abs((ep - em) / (2 * eps) - g[i, q]) < 1e-7


@pytest.mark.parametrize("n,L", [(21, 6), (23, 6), (24, 6), (26, 4)])
@pytest.mark.parametrize("p_z", [0.0, 0.15, 1.0 / 3.0, 0.6, 0.9, 1.0])
def test_axis_aware_plans_many_axis_patterns(n, L, p_z):
    """Planner stress: layers from all-X/Y to all-Rz (pass counts, splits, fillers, Rz slots, trades with the contiguous pass,
    the three-round contiguous pass and the fallback to the static plan all occur) -- axis-aware plans against the static plan
    of the same library, 12- and 11-bit tiles."""
    from qradient_b200.circuit_logic import McClean
    rng = np.random.default_rng(int(1000 * p_z) + n)
    axes = np.where(rng.random((L, n)) < p_z, 2, rng.integers(0, 2, (L, n)))
    angles = rng.uniform(0, 2 * np.pi, (L, n))
    zz = np.full((n, n), None)
    zz[0, 1], zz[2, n - 1] = 1.0, -0.7
    obs = {"zz": zz, "x": np.array([0.3] + [None] * (n - 1)), "y": np.array([None] * (n - 1) + [0.2])}
    c = McClean(n, obs, L, axes=axes, angles=angles)
    for tile_bits in (12, 11):
        c.state.set_option("tile_bits", tile_bits)
        c.state.set_option("axis_plan", 0)
        e0, g0 = c.grad_run()
        v0 = np.array(c.state.vec) if n <= 23 else None
        for mode in (15, 13, 1):
            c.state.set_option("axis_plan", mode)
            e1, g1 = c.grad_run()
            assert_parity(e1, g1, e0, g0, 2.2, 1e-12)
            assert abs(c.run_expec_val() - e0) <= 1e-12
            if v0 is not None and mode == 15:
                c.grad_run()
                np.testing.assert_allclose(c.state.vec, v0, atol=1e-13)


@pytest.mark.parametrize("n", [20, 22, 24])
def test_classifier_sublayers_with_axis_aware_plans(n):
    """Layered circuits (qr_layered_grad: sub-layers of one axis each, a ladder before every third one): all-Rz sub-layers,
    contiguous passes on general tiles without a ladder gather -- axis-aware plans against the static plan."""
    from qradient_b200.circuit_logic import MeynardClassifier
    rng = np.random.default_rng(n)
    Le, Lc = 1, 2
    data, enc, cls = rng.uniform(0, np.pi, (Le, n)), rng.uniform(0, 2 * np.pi, (Le, n, 2)), rng.uniform(0, 2 * np.pi, (Lc, n, 3))
    c = MeynardClassifier(n, Le, Lc)
    c.state.set_option("axis_plan", 0)
    e0, ge0, gc0 = c.grad_run(data, enc, cls)
    for mode in (15, 1):
        c.state.set_option("axis_plan", mode)
        e1, ge1, gc1 = c.grad_run(data, enc, cls)
        assert abs(e1 - e0) <= 1e-12
        np.testing.assert_allclose(ge1, ge0, rtol=0, atol=1e-12)
        np.testing.assert_allclose(gc1, gc0, rtol=0, atol=1e-12)
        assert abs(c.run_expec_val(data, enc, cls) - e0) <= 1e-12


def test_ini_state_with_axis_aware_plans():
    """`ini_state` given (mc_clean.py:30-36: the Ry(pi/4) layer is applied as gates, on the static plan) followed by layers on
    the axis-aware plans."""
    from qradient_b200.circuit_logic import McClean
    n, L = 21, 3
    rng = np.random.default_rng(21)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    ini = rng.normal(size=2 ** n) + 1j * rng.normal(size=2 ** n)
    ini /= np.linalg.norm(ini)
    zz = np.full((n, n), None)
    zz[0, 1], zz[4, n - 1] = 1.0, 0.5
    c = McClean(n, {"zz": zz}, L, axes=axes, angles=angles)
    c.state.set_option("axis_plan", 0)
    e0, g0 = c.grad_run(ini_state=ini)
    c.state.set_option("axis_plan", 15)
    e1, g1 = c.grad_run(ini_state=ini)
    assert_parity(e1, g1, e0, g0, 1.5, 1e-12)
    assert abs(c.run_expec_val(ini_state=ini) - e0) <= 1e-12


def test_batched_14_qubits_matches_single():
    from qradient_b200.circuit_logic import McClean
    n, L, B = 14, 4, 64
    rng = np.random.default_rng(4)
    axes, angles = rng.integers(0, 3, (B, L, n)), rng.uniform(0, 2 * np.pi, (B, L, n))
    c = McClean(n, zz01(n), L, axes=axes[0], angles=angles[0])
    e, g = c.grad_run_batch(angles, axes)
    for b in (0, 17, 63):
        e_ref, g_ref = orc.mcclean_grad_run(n, zz01(n), axes[b], angles[b])
        assert_parity(e[b], g[b], e_ref, g_ref, 1.0, 1e-10)


@pytest.mark.parametrize("n,L", [(12, 3), (14, 6), (20, 8), (24, 3)])
def test_programmatic_dependent_launch_is_bit_identical(n, L):
    """QR_OPT_PDL: tile passes launched with programmatic stream serialization (the next pass's CTAs queue up while
    the previous pass drains; every pass orders its memory accesses with griddepcontrol.wait).  Same kernels, same
    grids, same reduction order => results must be IDENTICAL to the serialized launches; a missing dependency would
    show up as a difference (or a run-to-run variation) because consecutive passes read what the previous one wrote."""
    from qradient_b200.circuit_logic import McClean, Qaoa
    from qradient_b200.optimization_problems import MaxCut
    rng = np.random.default_rng(200 + n)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    zz = np.full((n, n), None)
    zz[0, 1] = 1.0
    obs = {"zz": zz, "x": np.array([0.3] + [None] * (n - 1), dtype=object)}
    c = McClean(n, obs, L, axes=axes, angles=angles)
    c.state.set_option("pdl", 0)
    e0, g0 = c.grad_run()
    r0 = c.run_expec_val()
    for mode in (1, 2):
        c.state.set_option("pdl", mode)
        for _ in range(4):
            e1, g1 = c.grad_run()
            assert e1 == e0 and np.array_equal(g1, g0)
        assert c.run_expec_val() == r0
    if n <= 20:
        B = 8
        ax, an = rng.integers(0, 3, (B, L, n)), rng.uniform(0, 2 * np.pi, (B, L, n))
        c.state.set_option("pdl", 0)
        eb0, gb0 = c.grad_run_batch(an, ax)
        c.state.set_option("pdl", 2)
        eb1, gb1 = c.grad_run_batch(an, ax)
        assert np.array_equal(eb0, eb1) and np.array_equal(gb0, gb1)
    q = Qaoa(n, MaxCut(n, edge_set=[(i, i + 1) for i in range(n - 1)]).to_observable(), 3)
    b, gm = rng.random(3), rng.random(3)
    q.state.set_option("pdl", 0)
    e0, g0 = q.grad_run(b, gm)
    q.state.set_option("pdl", 2)
    for _ in range(3):
        e1, g1 = q.grad_run(b, gm)
        assert e1 == e0 and np.array_equal(g1, g0)
