"""Optimiser loop: host loop (McCleanOpt.step, one grad_run + host update per step) vs device loop (McCleanOpt.run)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qradient_b200.circuit_logic import McClean
from qradient_b200.optimization import McCleanOpt

for n, L, steps in ((8, 8, 200), (12, 12, 200), (16, 16, 100), (20, 20, 50), (24, 10, 10)):
    rng = np.random.default_rng(n)
    zz = np.full((n, n), None); zz[0, 1] = 1.0
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    res = {}
    for mode in ("host", "device"):
        c = McClean(n, {"zz": zz}, L, axes=axes, angles=angles.copy())
        o = McCleanOpt(c, {"name": "Adam", "step_size": 0.05}, max_iter=steps + 8, ini_parameters=angles.copy())
        (o.run if mode == "device" else (lambda k: [o.step() for _ in range(k)]))(3)     # warm-up
        t0 = time.perf_counter()
        if mode == "device":
            o.run(steps)
        else:
            for _ in range(steps):
                o.step()
        res[mode] = (time.perf_counter() - t0, o.cost_history[:steps + 3].copy())
    d = np.abs(res["host"][1] - res["device"][1]).max()
    print("n=%d L=%d steps=%d host %.3f ms/step device %.3f ms/step (x%.2f) max|dcost|=%.1e" % (
        n, L, steps, 1e3 * res["host"][0] / steps, 1e3 * res["device"][0] / steps, res["host"][0] / res["device"][0], d), flush=True)

# QAOA MaxCut (QaoaOpt.step vs QaoaOpt.run = qr_qaoa_optimize)
from qradient_b200.circuit_logic import Qaoa  # noqa: E402
from qradient_b200.optimization import QaoaOpt  # noqa: E402
from qradient_b200.optimization_problems import MaxCut  # noqa: E402

for n, p, steps in ((8, 4, 200), (12, 6, 200), (16, 8, 100), (20, 10, 50), (24, 10, 10)):
    rng = np.random.default_rng(n)
    obs = MaxCut(n, edge_num=min(2 * n, n * (n - 1) // 2)).to_observable()
    b0, g0 = rng.random(p), rng.random(p)
    res = {}
    for mode in ("host", "device"):
        q = Qaoa(n, obs, p)
        o = QaoaOpt(q, {"name": "Adam", "step_size": 0.02}, b0.copy(), g0.copy(), max_iter=steps + 8)
        (o.run if mode == "device" else (lambda k: [o.step() for _ in range(k)]))(3)     # warm-up
        t0 = time.perf_counter()
        if mode == "device":
            o.run(steps)
        else:
            for _ in range(steps):
                o.step()
        res[mode] = (time.perf_counter() - t0, o.cost_history[:steps + 3].copy())
    d = np.abs(res["host"][1] - res["device"][1]).max()
    print("QAOA n=%d p=%d steps=%d host %.3f ms/step device %.3f ms/step (x%.2f) max|dcost|=%.1e" % (
        n, p, steps, 1e3 * res["host"][0] / steps, 1e3 * res["device"][0] / steps, res["host"][0] / res["device"][0], d), flush=True)
