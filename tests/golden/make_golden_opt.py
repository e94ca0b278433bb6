"""GV13: the reference's optimiser loops (optimization.py: McCleanOpt / QaoaOpt with Adam, GradientDescent,
RateDecayOnPlateau) run UNMODIFIED on the reference's circuit classes (same shim as make_golden.py).

    python tests/golden/make_golden_opt.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import make_golden as mg  # noqa: E402  (loads the reference modules; writes nothing on import)

opt = mg._load("_ref_optimization", os.path.join(mg.REF, "qradient/optimization.py"))


def run_mcclean(name, optimizer, steps, n=5, L=3, seed=13):
    rng = np.random.default_rng(seed)
    obs = {"x": np.array([0.5] + [None] * (n - 1), dtype=object), "zz": np.full((n, n), None)}
    obs["zz"][0, 1] = 1.0
    obs["zz"][2, 4] = -0.6
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = mg.McClean(n, obs, L, axes=axes, angles=angles.copy())
    o = opt.McCleanOpt(c, dict(optimizer), max_iter=steps + 1, ini_parameters=angles.copy())
    for _ in range(steps):
        o.step()
    return {name + "_cost": o.cost_history[:steps].copy(), name + "_params": o.param_history[:steps + 1].copy(),
            "mc_axes": axes, "mc_angles": angles, **{"mc_" + k: v for k, v in mg.obs_to_arrays(n, obs).items()}}


def run_qaoa(name, optimizer, steps, n=6, p=2):
    edges = np.array([[0, 1], [1, 2], [2, 3], [3, 4], [4, 5], [0, 5], [1, 4]])
    c = mg.Qaoa(n, mg.MaxCut(n, edge_set=edges).to_observable(), p)
    betas, gammas = np.array([0.3, 0.7]), np.array([0.2, 0.9])
    o = opt.QaoaOpt(c, dict(optimizer), betas.copy(), gammas.copy(), max_iter=steps + 1)
    for _ in range(steps):
        o.step()
    return {name + "_cost": o.cost_history[:steps].copy(), name + "_params": o.param_history[:steps + 1].copy(),
            "qa_edges": edges, "qa_betas": betas, "qa_gammas": gammas}


if __name__ == "__main__" and (len(sys.argv) < 2 or "gv13" in sys.argv[1:]):
    d = {"steps": 6}
    d.update(run_mcclean("mc_adam", {"name": "Adam", "step_size": 0.05}, 6))
    d.update(run_mcclean("mc_gd", {"name": "GradientDescent", "step_size": 0.1}, 6))
    d.update(run_mcclean("mc_plateau", {"name": "RateDecayOnPlateau", "step_size": 0.8, "plateau_length": 1, "decay_rate": 0.5}, 6))
    d.update(run_qaoa("qa_adam", {"name": "Adam", "step_size": 0.05, "beta1": 0.8}, 6))
    d.update(run_qaoa("qa_gd", {"name": "GradientDescent", "step_size": 0.02}, 6))
    mg.save("gv13_optimizers", **d)


def gv14_mcclean_sample_grad_dense():
    """GV14: McClean.sample_grad_dense (mc_clean.py:207-268) for diagonal observables: a non-degenerate one and
    ZZ(0,1) (two 8-fold degenerate eigenvalues)."""
    out = {}
    for tag, n, L, shots, seed in (("a", 4, 3, 20, 3), ("b", 4, 2, 15, 8)):
        rng = np.random.default_rng(21 + seed)
        obs = {"zz": np.full((n, n), None)}
        obs["zz"][0, 1] = 1.0
        if tag == "a":
            obs["z"] = np.array([0.4, None, -0.7, None], dtype=object)
            obs["zz"][1, 3] = 0.5
        axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
        c = mg.McClean(n, obs, L, axes=axes, angles=angles)
        np.random.seed(seed)
        e, g = c.sample_grad_dense(shot_num=shots)
        out.update({tag + "_n": n, tag + "_L": L, tag + "_shots": shots, tag + "_seed": seed, tag + "_axes": axes, tag + "_angles": angles,
                    tag + "_E": e, tag + "_grad": g, tag + "_eigenvalues": c.eigenvalues,
                    **{tag + "_" + k: v for k, v in mg.obs_to_arrays(n, obs).items()}})
    mg.save("gv14_mcclean_sample_grad_dense", **out)


def gv15_mcclean_sample_grad_dense_xy():
    """GV15: McClean.sample_grad_dense (mc_clean.py:207-268) for observables with x / y terms (dense eigensystem):
    a generic one and X(0) + ZZ(0,1) (degenerate spectrum)."""
    out = {}
    for tag, n, L, shots, seed in (("a", 4, 3, 20, 4), ("b", 5, 2, 12, 9)):
        rng = np.random.default_rng(31 + seed)
        obs = {"zz": np.full((n, n), None), "x": np.array([0.5] + [None] * (n - 1), dtype=object)}
        obs["zz"][0, 1] = 1.0
        if tag == "a":
            obs["y"] = np.array([None, None, 0.3, None], dtype=object)
            obs["z"] = np.array([0.4, None, -0.7, 0.15], dtype=object)
            obs["zz"][1, 3] = 0.5
        axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
        c = mg.McClean(n, obs, L, axes=axes, angles=angles)
        np.random.seed(seed)
        e, g = c.sample_grad_dense(shot_num=shots)
        out.update({tag + "_n": n, tag + "_L": L, tag + "_shots": shots, tag + "_seed": seed, tag + "_axes": axes, tag + "_angles": angles,
                    tag + "_E": e, tag + "_grad": g, tag + "_eigenvalues": c.eigenvalues,
                    **{tag + "_" + k: v for k, v in mg.obs_to_arrays(n, obs).items()}})
    mg.save("gv15_mcclean_sample_grad_dense_xy", **out)


def gv16_mcclean_sample_grad_dense_component_sampling():
    """GV16: McClean.sample_grad_dense_with_component_sampling (mc_clean.py:277-350): one observable component drawn with
    np.random.choice, measured in ITS eigenbasis."""
    out = {}
    for tag, n, L, shots, seed in (("a", 4, 2, 16, 6), ("b", 4, 2, 10, 11)):
        rng = np.random.default_rng(41 + seed)
        obs = {"zz": np.full((n, n), None), "x": np.array([0.5] + [None] * (n - 1), dtype=object),
               "z": np.array([None, 0.8, None, 0.3], dtype=object)}
        obs["zz"][0, 1] = 1.0
        obs["zz"][1, 3] = 0.6
        if tag == "b":
            obs["y"] = np.array([None, None, 0.7, None], dtype=object)
        axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
        c = mg.McClean(n, obs, L, use_observable_components=True, axes=axes, angles=angles)
        np.random.seed(seed)
        e, g = c.sample_grad_dense_with_component_sampling(shot_num=shots)
        out.update({tag + "_n": n, tag + "_L": L, tag + "_shots": shots, tag + "_seed": seed, tag + "_axes": axes, tag + "_angles": angles,
                    tag + "_E": e, tag + "_grad": g, **{tag + "_" + k: v for k, v in mg.obs_to_arrays(n, obs).items()}})
    mg.save("gv16_mcclean_sample_grad_dense_component_sampling", **out)


def gv17_mcclean_sample_grad_with_component_sampling():
    """GV17: McClean.sample_grad_with_component_sampling (mc_clean.py:158-198): per parameter one component drawn with
    np.random.choice, per-term Bernoulli estimates (base.py:35-46) of the two shifted circuits."""
    n, L, shots, seed = 4, 2, 30, 14
    rng = np.random.default_rng(55)
    obs = {"zz": np.full((n, n), None), "x": np.array([0.5] + [None] * (n - 1), dtype=object),
           "y": np.array([None, None, 0.7, None], dtype=object), "z": np.array([None, 0.8, None, 0.3], dtype=object)}
    obs["zz"][0, 1] = 1.0
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = mg.McClean(n, obs, L, use_observable_components=True, axes=axes, angles=angles)
    np.random.seed(seed)
    e, g = c.sample_grad_with_component_sampling(shot_num=shots)
    mg.save("gv17_mcclean_sample_grad_component_sampling", n=n, L=L, shots=shots, seed=seed, axes=axes, angles=angles, E=e, grad=g,
            **mg.obs_to_arrays(n, obs))


if __name__ == "__main__":
    import sys as _sys
    which = _sys.argv[1:] or ["gv14", "gv15", "gv16", "gv17"]
    if "gv14" in which:
        gv14_mcclean_sample_grad_dense()
    if "gv15" in which:
        gv15_mcclean_sample_grad_dense_xy()
    if "gv16" in which:
        gv16_mcclean_sample_grad_dense_component_sampling()
    if "gv17" in which:
        gv17_mcclean_sample_grad_with_component_sampling()
