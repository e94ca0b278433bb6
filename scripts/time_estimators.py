"""GPU timings of the sampled gradient estimators at the sizes the reference publishes for its laptop CPU
(tutorials/timings/timing-test.ipynb cell 3, 100 layers):  McClean.sample_grad n = 11/12/13: 35.76 / 68.48 / 130.21 s;
sample_grad_dense ("sample_grad_observable") n = 9/10/11: 4.416 / 14.713 / 80.638 s; qaoa.sample_grad_dense n = 7/8/9:
3.909 / 5.898 / 10.816 s; McClean.grad_run n = 17/18/19: 9.545 / 22.091 / 46.849 s; qaoa.grad_run: 6.029 / 14.512 / 30.302 s."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qradient_b200.circuit_logic import McClean, Qaoa  # noqa: E402
from qradient_b200.optimization_problems import MaxCut  # noqa: E402

QUICK = bool(os.environ.get("QUICK"))        # CPU check of the script itself on the emulated library
L = 2 if QUICK else 100
SHRINK = 4 if QUICK else 0
out = {}


def zz01(n):
    m = np.full((n, n), None)
    m[0, 1] = 1.0
    return {"zz": m}


def timed(fn, reps=1):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps


rng = np.random.default_rng(0)
for n in (11 - SHRINK, 12 - SHRINK, 13 - SHRINK):
    c = McClean(n, zz01(n), L, axes=rng.integers(0, 3, (L, n)), angles=rng.uniform(0, 2 * np.pi, (L, n)))
    out["mcclean.sample_grad n=%d" % n] = timed(lambda: c.sample_grad(shot_num=100))
for n in (9 - SHRINK, 10 - SHRINK, 11 - SHRINK):
    c = McClean(n, zz01(n), L, axes=rng.integers(0, 3, (L, n)), angles=rng.uniform(0, 2 * np.pi, (L, n)))
    out["mcclean.sample_grad_dense n=%d" % n] = timed(lambda: c.sample_grad_dense(shot_num=100))
for n in (7 - SHRINK, 8 - SHRINK, 9 - SHRINK):
    q = Qaoa(n, MaxCut(n, edge_num=min(10, n * (n - 1) // 2)).to_observable(), L)
    b, g = rng.random(L), rng.random(L)
    out["qaoa.sample_grad_dense n=%d" % n] = timed(lambda: q.sample_grad_dense(b, g, shot_num=100))
for n in (17 - 2 * SHRINK, 18 - 2 * SHRINK, 19 - 2 * SHRINK):
    c = McClean(n, zz01(n), L, axes=rng.integers(0, 3, (L, n)), angles=rng.uniform(0, 2 * np.pi, (L, n)))
    out["mcclean.grad_run n=%d" % n] = timed(lambda: c.grad_run(), reps=5)
    q = Qaoa(n, MaxCut(n, edge_num=min(10, n * (n - 1) // 2)).to_observable(), L)
    b, g = rng.random(L), rng.random(L)
    out["qaoa.grad_run n=%d" % n] = timed(lambda: q.grad_run(b, g), reps=5)
print(json.dumps({k: round(v, 5) for k, v in out.items()}, indent=1))
