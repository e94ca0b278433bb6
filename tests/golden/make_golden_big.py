"""Full-size oracle fixtures for the BASELINE configs the reference-as-written cannot run
(memory wall): generated ONCE in the build container by the repo's oracle
(oracle/qr_oracle.py, itself pinned to reference outputs by tests/test_oracle.py at n <= 20).

    python tests/golden/make_golden_big.py [cfg3] [cfg4] [cfg5s]

  gv18_qaoa_config3_26x10.npz   BASELINE config 3: QAOA MaxCut 3-regular, 26 qubits, p = 10
        inputs : bench.CONFIG3_EDGES, rng = default_rng(10); gammas = rng.random(p); betas = rng.random(p)
        outputs: E, grad[10, 2] (qaoa.py:40-70), the 100 bitstring indices drawn from |psi_final|^2 with
                 U = RandomState(0).uniform(size=100) (qaoa.py:196-198), mean H over those indices, H at those
                 indices; needs ~27 GiB of host memory (21 history vectors) and ~15 min on one core
  gv19_mcclean_config4_14x14.npz BASELINE config 4: 8 of the 8192 parameter sets of the batched 14 x 14 workload
        inputs : rng = default_rng(4); axes = rng.integers(0, 3, (8192, 14, 14)); angles = rng.uniform(0, 2 pi, same)
        outputs: indices[8], E[8], grad[8, 14, 14] (mc_clean.py:47-78)
  gv20_mcclean_24x6.npz         a sharded-path anchor above the sizes the CPU tier tests: McClean 24 qubits x 6 layers,
        rng = default_rng(24), ZZ(0,1) + 0.5 X_2 + 0.25 Y_23 (an x term on a rank-bit qubit, a y term on the lowest bit)
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import qr_oracle as orc   # noqa: E402
import bench                          # noqa: E402


def cfg3():
    n, p = 26, 10
    rng = np.random.default_rng(10)
    gammas, betas = rng.random(p), rng.random(p)
    obs = orc.OracleObservable(n, orc.maxcut_observable(n, bench.CONFIG3_EDGES))
    t = time.time()
    e, grad, psi = orc.qaoa_grad_run(n, obs, betas, gammas, return_state=True)
    u = np.random.RandomState(0).uniform(size=100)
    idx = orc.sample_bitstrings(psi, u)
    ham = orc.classical_ham_vector(obs)
    cdf = np.cumsum(np.abs(psi) ** 2)
    # distance of every uniform from the cdf steps next to its index: how close a parallel scan may come to flipping it
    lo = np.where(idx > 0, cdf[np.maximum(idx - 1, 0)], 0.0)
    margin = np.minimum(u - lo, cdf[idx] - u)
    np.savez(os.path.join(HERE, "gv18_qaoa_config3_26x10.npz"), n=n, p=p, betas=betas, gammas=gammas,
             edges=np.array(bench.CONFIG3_EDGES), e=e, grad=grad, uniforms=u, idx=idx.astype(np.int64),
             ham_at_idx=ham[idx], mean_cost=float(ham[idx].mean()), cdf_margin=margin, norm2=float(cdf[-1]))
    print("cfg3: E = %.15g, |grad|max = %.3g, min cdf margin = %.3g, %.0f s" % (e, np.abs(grad).max(), margin.min(), time.time() - t))


def cfg4():
    n, L, B = 14, 14, 8192
    rng = np.random.default_rng(4)
    axes, angles = rng.integers(0, 3, (B, L, n)), rng.uniform(0, 2 * np.pi, (B, L, n))
    zz = np.full((n, n), None)
    zz[0, 1] = 1.0
    sel = np.array([0, 1, 2, 3, 4095, 4096, 8190, 8191])
    es, gs = [], []
    for b in sel:
        e, g = orc.mcclean_grad_run(n, {"zz": zz}, axes[b], angles[b])
        es.append(e)
        gs.append(g)
    np.savez(os.path.join(HERE, "gv19_mcclean_config4_14x14.npz"), n=n, L=L, B=B, seed=4, indices=sel, e=np.array(es), grad=np.array(gs))
    print("cfg4: E =", es)


def cfg5s():
    n, L = 24, 6
    rng = np.random.default_rng(24)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    zz = np.full((n, n), None)
    zz[0, 1] = 1.0
    x = np.array([None] * n, dtype=object)
    x[2] = 0.5
    y = np.array([None] * n, dtype=object)
    y[n - 1] = 0.25
    t = time.time()
    e, g = orc.mcclean_grad_run(n, {"zz": zz, "x": x, "y": y}, axes, angles)
    np.savez(os.path.join(HERE, "gv20_mcclean_24x6.npz"), n=n, L=L, axes=axes, angles=angles, e=e, grad=g)
    print("cfg5s: E = %.15g, %.0f s" % (e, time.time() - t))


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg4", "cfg5s", "cfg3"]
    for w in which:
        {"cfg3": cfg3, "cfg4": cfg4, "cfg5s": cfg5s}[w]()
