"""GPU check of the pair kernel (QR_OPT_PAIR) against the default kernels."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qradient_b200.circuit_logic import McClean, Qaoa
from qradient_b200.optimization_problems import MaxCut

def mixed(n):
    zz = np.full((n, n), None); zz[0, 1] = 1.0; zz[2, n - 1] = -0.5
    return {"zz": zz, "x": np.array([0.3] + [None] * (n - 1), dtype=object), "y": np.array([None] * (n - 1) + [0.2], dtype=object)}

ok = True
for n, L in ((12, 3), (13, 2), (15, 2), (16, 2), (19, 2), (20, 3), (22, 2), (25, 2)):
    rng = np.random.default_rng(n)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = McClean(n, mixed(n), L, axes=axes, angles=angles)
    c.state.set_option("tile_bits", 12)
    e0, g0 = c.grad_run(); v0 = np.array(c.state.vec) if n <= 20 else None
    r0 = c.run_expec_val()
    for mode in (1, 2, 3):
        c.state.set_option("pair", mode)
        for rep in range(2):
            e1, g1 = c.grad_run()
            de, dg = abs(e1 - e0), np.abs(g1 - g0).max()
            dv = np.abs(np.array(c.state.vec) - v0).max() if v0 is not None else 0.0
            r1 = c.run_expec_val()
            good = de < 1e-12 and dg < 1e-12 and dv < 1e-12 and abs(r1 - r0) < 1e-12
            ok &= good
            print("mcclean n=%d L=%d pair=%d rep=%d dE=%.1e dg=%.1e dvec=%.1e dErun=%.1e %s" % (n, L, mode, rep, de, dg, dv, abs(r1 - r0), "OK" if good else "FAIL"), flush=True)
    c.state.set_option("pair", 0)
for n in (13, 18, 22):
    rng = np.random.default_rng(n)
    q = Qaoa(n, MaxCut(n, edge_set=MaxCut.random_regular(n, 3, seed=n) if n % 2 == 0 else [(i, (i + 1) % n) for i in range(n - 1)]).to_observable(), 2)
    q.state.set_option("tile_bits", 12)
    b, g = rng.random(2), rng.random(2)
    e0, g0 = q.grad_run(b, g)
    q.state.set_option("pair", 3)
    e1, g1 = q.grad_run(b, g)
    good = abs(e1 - e0) < 1e-11 and np.abs(g1 - g0).max() < 1e-11
    ok &= good
    print("qaoa n=%d dE=%.1e dg=%.1e %s" % (n, abs(e1 - e0), np.abs(g1 - g0).max(), "OK" if good else "FAIL"), flush=True)
print("ALL OK" if ok else "SOME FAIL")
