"""Generate tests/golden/*.npz by running the REFERENCE's own algorithm files.

Run in the build container only (needs /root/reference; the GPU box does not have it):

    python tests/golden/make_golden.py

The reference package does not import at HEAD (`circuit_logic/mc_clean.py:2` imports a
`Gates` class that `physical_components/__init__.py` does not define), so this script
loads `circuit_logic/base.py`, `mc_clean.py` and `qaoa.py` UNMODIFIED through importlib and
supplies the missing `qradient.physical_components` dialect (`Gates`, `State`, `Observable`)
as a thin adapter around the reference's own `physical_components/state.py` and
`observable.py` (their sparse generators, ladder products and Kronecker sums are built by
the reference code itself; the adapter only wires attribute names).  Nothing here is
product code and no reference source is copied into the repository: only the numeric
inputs/outputs are saved.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.sparse as sp

REF = os.environ.get("QRADIENT_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


ref_state = _load("_ref_state", os.path.join(REF, "qradient/physical_components/state.py"))
ref_obs = _load("_ref_observable", os.path.join(REF, "qradient/physical_components/observable.py"))
RefState = ref_state.State


class Gates:
    """Dialect-B gate container; every matrix is produced by the reference State loaders."""

    def __init__(self, qnum):
        self.qnum = qnum
        self._ref = RefState(qnum)

    def add_xrots(self):
        self._ref.load_xrots()
        self.xrot = self._ref._State__xrot
        return self

    def add_yrots(self):
        self._ref.load_yrots()
        self.yrot = self._ref._State__yrot
        return self

    def add_zrots(self):
        self._ref.load_zrots()
        self.zrot_pos = self._ref._State__zrot_pos
        self.zrot_neg = self._ref._State__zrot_neg
        return self

    def add_cnot_ladder(self, periodic=False):
        self._ref.load_cnot_ladder(periodic)
        self.cnot_ladder = self._ref._State__cnot_ladder
        return self

    def add_x_summed(self):
        self._ref.load_xrot_all()
        self.x_summed = self._ref._State__xrot_all
        return self

    def add_classical_ham(self, observable, include_individual_components=False):
        # state.py:273-292 with `self.info` read as `observable.info` (HEAD typo at :278)
        n = self.qnum
        z = np.array([1.0, -1.0])
        ones = lambda i: np.full(2 ** i, 1.0)
        ham = np.zeros(2 ** n)
        comps = []
        info = observable.info
        if "z" in info and info["z"] is not None:
            for i, w in enumerate(info["z"]):
                if w is not None:
                    c = w * np.kron(ones(i), np.kron(z, ones(n - i - 1)))
                    ham += c
                    comps.append(c)
        if "zz" in info and info["zz"] is not None:
            for i in range(n):
                for j in range(i + 1, n):
                    if info["zz"][i, j] is not None:
                        c = info["zz"][i, j] * np.kron(
                            ones(i), np.kron(z, np.kron(ones(j - i - 1), np.kron(z, ones(n - j - 1)))))
                        ham += c
                        comps.append(c)
        self.classical_ham = ham
        if include_individual_components:
            self.classical_ham_components = np.array(comps)
        return self


class State:
    """Dialect-B State: same update formulas as reference state.py, reading from `gates`."""

    def __init__(self, qnum, ini="0"):
        self._qnum, self._ini = qnum, ini
        self.reset()

    def reset(self, ini=None):
        if ini is not None:
            self._ini = ini
        r = RefState(self._qnum, self._ini)   # reference reset(): state.py:61-71
        self.vec = r.vec

    def _ref_with_vec(self):
        # run a reference State method on our vector with our gates' matrices
        r = RefState.__new__(RefState)
        r._State__qnum, r._State__ini, r.vec = self._qnum, self._ini, self.vec
        g = self.gates
        for attr in ("xrot", "yrot", "zrot_pos", "zrot_neg"):
            if hasattr(g, attr):
                setattr(r, "_State__" + attr, getattr(g, attr))
        return r

    def _call(self, name, *a):
        r = self._ref_with_vec()
        getattr(r, name)(*a)
        self.vec = r.vec

    def xrot(self, angle, i): self._call("xrot", angle, i)
    def yrot(self, angle, i): self._call("yrot", angle, i)
    def zrot(self, angle, i): self._call("zrot", angle, i)
    def dxrot(self, angle, i): self._call("dxrot", angle, i)
    def dyrot(self, angle, i): self._call("dyrot", angle, i)
    def dzrot(self, angle, i): self._call("dzrot", angle, i)

    def cnot_ladder(self, stacking):            # state.py:251 intent
        self.vec = self.gates.cnot_ladder[stacking].dot(self.vec)

    def multiply_matrix(self, m):               # mc_clean.py:65
        self.vec = m.dot(self.vec)

    def exp_ham_classical(self, angle):         # state.py:301
        self.vec = self.vec * np.exp(-1.0j * angle * self.gates.classical_ham)

    def exp_ham_classical_component(self, angle, i):   # state.py:311
        self.vec = self.vec * np.exp(-1.0j * angle * self.gates.classical_ham_components[i])

    def ham_classical(self):                    # state.py:321
        self.vec = self.vec * (-1.0j * self.gates.classical_ham)

    def x_summed(self):                         # state.py:123
        self.vec = self.gates.x_summed.dot(self.vec)

    def norm_error(self):
        return 1.0 - np.linalg.norm(self.vec)


class Observable(ref_obs.Observable):
    """Reference Observable with the attributes its ctor forgets to set (observable.py:30-40)."""

    def __init__(self, qnum, observable, store_components=False):
        self.qnum = qnum
        self.info = observable
        self.dict = observable
        self.store_components = store_components
        self.has_loaded_projectors = False
        self.load_matrix(observable)

    def check_observable(self, known_keys, warning=None):
        pass


shim = types.ModuleType("qradient.physical_components")
shim.Gates, shim.State, shim.Observable = Gates, State, Observable
pkg = types.ModuleType("qradient"); pkg.__path__ = []
cl = types.ModuleType("qradient.circuit_logic"); cl.__path__ = []
sys.modules.update({"qradient": pkg, "qradient.physical_components": shim, "qradient.circuit_logic": cl})
import tqdm  # noqa: E402  (base.py imports tnrange)
base = _load("qradient.circuit_logic.base", os.path.join(REF, "qradient/circuit_logic/base.py"))
mc = _load("qradient.circuit_logic.mc_clean", os.path.join(REF, "qradient/circuit_logic/mc_clean.py"))
qa = _load("qradient.circuit_logic.qaoa", os.path.join(REF, "qradient/circuit_logic/qaoa.py"))
problems = _load("_ref_problems", os.path.join(REF, "qradient/optimization_problems.py"))
McClean, Qaoa, MaxCut = mc.McClean, qa.Qaoa, problems.MaxCut


def zz01(n):
    m = np.full((n, n), None)
    m[0, 1] = 1.0
    return {"zz": m}


def obs_to_arrays(n, obs):
    """Serialise an observable dict: NaN marks None."""
    out = {}
    for k in ("x", "y", "z"):
        a = np.full(n, np.nan)
        if k in obs:
            for i, w in enumerate(obs[k]):
                if w is not None:
                    a[i] = w
        out["obs_" + k] = a
    zz = np.full((n, n), np.nan)
    if "zz" in obs:
        for i in range(n):
            for j in range(n):
                if obs["zz"][i, j] is not None:
                    zz[i, j] = obs["zz"][i, j]
    out["obs_zz"] = zz
    return out


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote", path, {k: np.asarray(v).shape for k, v in kw.items()})


def gv_mcclean(name, n, L, obs, axes=None, angles=None, seed=None, keep_state=True, ini_state=None):
    if seed is not None:
        np.random.seed(seed)
    kw = {}
    if axes is not None:
        kw = dict(axes=np.array(axes), angles=np.array(angles, dtype=float))
    c = McClean(n, obs, L, **kw)
    e_run = c.run_expec_val(ini_state=None if ini_state is None else ini_state.copy())
    psi_final = c.state.vec.copy()
    e, grad = c.grad_run(ini_state=None if ini_state is None else ini_state.copy())
    costate = c.state.vec.copy()
    d = dict(n=n, L=L, axes=c.axes, angles=c.angles, E=e, E_run=e_run, grad=grad, **obs_to_arrays(n, obs))
    if keep_state:
        d.update(psi_final=psi_final, costate=costate)
    if ini_state is not None:
        d["ini_state"] = ini_state
    save(name, **d)
    return c


def gv_qaoa(name, n, p, edges, betas, gammas, shots_seed=0, n_bits=100, keep_state=True):
    obs = MaxCut(n, edge_set=np.array(edges)).to_observable()
    c = Qaoa(n, obs, p)
    betas, gammas = np.asarray(betas, float), np.asarray(gammas, float)
    e, grad = c.grad_run(betas, gammas)
    e_run = c.run_expec_val(betas, gammas)
    psi = c.state.vec.copy()
    # finite-shot expectation through the reference's own sample_expec_val (base.py:22-33)
    np.random.seed(shots_seed)
    e_shots = c.run_expec_val(betas, gammas, exact_expec_val=False, shot_num=100)
    # bitstring sampling through the reference's Qaoa.__sample mechanism (qaoa.py:196-198)
    from scipy import stats
    dist = np.abs(psi) ** 2
    np.random.seed(shots_seed)
    rv = stats.rv_discrete(values=(np.arange(2 ** n), dist))
    bits = rv.rvs(size=n_bits)
    uniforms = np.random.RandomState(shots_seed).uniform(size=n_bits)
    c.eigenvalues = c.state.gates.classical_ham
    d = dict(n=n, p=p, edges=np.array(edges), betas=betas, gammas=gammas, E=e, E_run=e_run, grad=grad,
             E_shots100_seed=e_shots, bitstrings=bits, uniforms=uniforms,
             bit_mean=c.eigenvalues[bits].mean(), **obs_to_arrays(n, obs))
    if keep_state:
        d.update(psi_final=psi, ham=c.state.gates.classical_ham)
    save(name, **d)


if __name__ == "__main__":
    # GV1: README example (README.md:30-37), global RNG seeded as in SURVEY appendix C
    gv_mcclean("gv1_mcclean_3x3", 3, 3, zz01(3), seed=0)
    # GV2: mixed x/y/z/zz observable, fixed parameters
    n = 4
    obs = {"x": np.array([0.3, None, None, None], dtype=object),
           "y": np.array([None, 0.7, None, None], dtype=object),
           "z": np.array([None, None, -1.1, None], dtype=object),
           "zz": np.full((4, 4), None)}
    obs["zz"][0, 1] = 1.0
    obs["zz"][1, 3] = -0.5
    gv_mcclean("gv2_mcclean_4x2_mixed", 4, 2, obs, axes=[[0, 1, 2, 0], [2, 0, 1, 1]],
               angles=[[.1, .2, .3, .4], [.5, .6, .7, .8]])
    # GV2b: same circuit started from a random normalised ini_state (mc_clean.py:32)
    rng = np.random.default_rng(7)
    ini = rng.normal(size=16) + 1j * rng.normal(size=16)
    ini /= np.linalg.norm(ini)
    gv_mcclean("gv2b_mcclean_4x2_ini", 4, 2, obs, axes=[[0, 1, 2, 0], [2, 0, 1, 1]],
               angles=[[.1, .2, .3, .4], [.5, .6, .7, .8]], ini_state=ini)
    # GV3: QAOA 4 qubits
    gv_qaoa("gv3_qaoa_4x2", 4, 2, [[0, 1], [1, 2], [0, 2], [2, 3]], [0.3, 0.7], [0.2, 0.9], n_bits=10)
    # GV4: QAOA 12 qubits, 3-regular graph (stand-in for config 3)
    edges12 = [(0, 2), (0, 6), (0, 9), (1, 2), (1, 3), (1, 4), (2, 8), (3, 5), (3, 11), (4, 7), (4, 10),
               (5, 7), (5, 10), (6, 7), (6, 8), (8, 9), (9, 11), (10, 11)]
    rng = np.random.default_rng(10)
    gammas = rng.random(3); betas = rng.random(3)
    gv_qaoa("gv4_qaoa_12x3", 12, 3, edges12, betas, gammas)
    # GV5: McClean 12x6, default_rng(1234)
    rng = np.random.default_rng(1234)
    gv_mcclean("gv5_mcclean_12x6", 12, 6, zz01(12), axes=rng.integers(0, 3, (6, 12)),
               angles=rng.uniform(0, 2 * np.pi, (6, 12)))
    # GV6: odd qubit number, every axis on every position, mixed observable incl. x/y on edge qubits
    n = 7
    rng = np.random.default_rng(77)
    obs = {"x": np.array([0.5] + [None] * 5 + [-0.25], dtype=object),
           "y": np.array([None, 0.4, None, None, None, None, 1.5], dtype=object),
           "z": np.array([0.1 * (i + 1) for i in range(7)], dtype=object),
           "zz": np.full((7, 7), None)}
    for (a, b, w) in [(0, 6, 0.9), (2, 3, -0.3), (1, 5, 0.2)]:
        obs["zz"][a, b] = w
    gv_mcclean("gv6_mcclean_7x5_mixed", 7, 5, obs, axes=rng.integers(0, 3, (5, 7)),
               angles=rng.uniform(0, 2 * np.pi, (5, 7)))
    # GV7: single-gate kernels, one reference call each (state.py gate formulas)
    n = 5
    rng = np.random.default_rng(5)
    v0 = rng.normal(size=32) + 1j * rng.normal(size=32)
    v0 /= np.linalg.norm(v0)
    g = Gates(n).add_xrots().add_yrots().add_zrots().add_cnot_ladder().add_x_summed()
    out = {"v0": v0}
    for name in ("xrot", "yrot", "zrot", "dxrot", "dyrot", "dzrot"):
        for q in range(n):
            s = State(n); s.gates = g; s.vec = v0.copy()
            getattr(s, name)(0.37 + 0.11 * q, q)
            out["%s_q%d" % (name, q)] = s.vec
    for st in (0, 1):
        s = State(n); s.gates = g; s.vec = v0.copy()
        s.cnot_ladder(st)
        out["ladder%d" % st] = s.vec
    s = State(n); s.gates = g; s.vec = v0.copy(); s.x_summed(); out["x_summed"] = s.vec
    r = RefState(n)
    for (ci, ti) in [(0, 1), (1, 0), (0, 4), (4, 0), (2, 3), (3, 1)]:
        m = r._State__cnot(ci, ti)
        out["cnot_%d_%d" % (ci, ti)] = m.dot(v0)
    save("gv7_gates_5", **out)
    # GV8: ladder basis maps n=2..10 (scatter form: image of basis state j), both stackings
    maps = {}
    for n in range(2, 11):
        r = RefState(n); r.load_cnot_ladder()
        for st in (0, 1):
            m = r._State__cnot_ladder[st].tocsc()
            maps["n%d_s%d" % (n, st)] = m.indices.astype(np.int64)  # column j -> row index
    save("gv8_ladder_maps", **maps)
    # GV9: McClean 10x4 finite-shot expectation through the reference's sample_expec_val
    n = 10
    rng = np.random.default_rng(9)
    obs = {"x": np.array([0.5] + [None] * 9, dtype=object), "z": np.array([None, 0.8] + [None] * 8, dtype=object),
           "zz": np.full((10, 10), None)}
    obs["zz"][0, 1] = 1.0; obs["zz"][3, 7] = -0.6
    c = McClean(n, obs, 4, axes=rng.integers(0, 3, (4, 10)), angles=rng.uniform(0, 2 * np.pi, (4, 10)))
    e_exact = c.run_expec_val()
    np.random.seed(3)
    e_shots = c.run_expec_val(exact_expec_val=False, shot_num=50)
    save("gv9_mcclean_10x4_shots", n=n, L=4, axes=c.axes, angles=c.angles, E=e_exact, E_shots50_seed3=e_shots,
         **obs_to_arrays(n, obs))
    # GV10: config 2 scalars only (McClean 20x20, default_rng(1234)); takes ~1 min and ~3 GiB
    if os.environ.get("QR_GOLDEN_BIG", "0") == "1":
        rng = np.random.default_rng(1234)
        axes = rng.integers(0, 3, (20, 20)); angles = rng.uniform(0, 2 * np.pi, (20, 20))
        c = McClean(20, zz01(20), 20, axes=axes, angles=angles)
        e, grad = c.grad_run()
        save("gv10_mcclean_20x20", n=20, L=20, axes=axes, angles=angles, E=e, grad=grad, **obs_to_arrays(20, zz01(20)))


def extra_goldens():
    """GV11: Qaoa.sample_grad_dense (qaoa.py:83-158); GV12: McClean.grad_run_with_component_sampling."""
    n, p = 4, 2
    edges = [[0, 1], [1, 2], [0, 2], [2, 3]]
    c = Qaoa(n, MaxCut(n, edge_set=np.array(edges)).to_observable(), p)
    betas, gammas = np.array([0.3, 0.7]), np.array([0.2, 0.9])
    np.random.seed(5)
    e, grad = c.sample_grad_dense(betas, gammas, shot_num=25)
    save("gv11_qaoa_sample_grad_dense", n=n, p=p, edges=np.array(edges), betas=betas, gammas=gammas, E=e, grad=grad, shot_num=25, seed=5)
    n, L = 5, 3
    rng = np.random.default_rng(12)
    obs = {"x": np.array([0.5, None, None, None, None], dtype=object), "z": np.array([None, 0.8, None, None, 0.3], dtype=object),
           "zz": np.full((5, 5), None)}
    obs["zz"][0, 1] = 1.0
    obs["zz"][2, 4] = 0.6
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = McClean(n, obs, L, use_observable_components=True, axes=axes, angles=angles)
    np.random.seed(2)
    e, grad = c.grad_run_with_component_sampling()
    save("gv12_mcclean_component_sampling", n=n, L=L, axes=axes, angles=angles, E=e, grad=grad, seed=2, **obs_to_arrays(n, obs))


if __name__ == "__main__" and os.environ.get("QR_GOLDEN_EXTRA", "1") == "1":
    extra_goldens()
