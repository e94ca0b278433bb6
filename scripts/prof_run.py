"""Small driver for ncu captures: one McClean grad_run at (n, L) on one GPU, or -- with --shards G -- the same circuit
through the swap engine on G virtual shards of that GPU (same kernels as a multi-GPU run; the peer pointers are local)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qradient_b200.circuit_logic import McClean  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=28)
ap.add_argument("--L", type=int, default=2)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--shards", type=int, default=1)
ap.add_argument("--seed", type=int, default=None)
ap.add_argument("--opt", action="append", default=[], help="name=value library option")
args = ap.parse_args()
rng = np.random.default_rng(args.n if args.seed is None else args.seed)
zz = np.full((args.n, args.n), None)
zz[0, 1] = 1.0
axes, angles = rng.integers(0, 3, (args.L, args.n)), rng.uniform(0, 2 * np.pi, (args.L, args.n))
if args.shards > 1:
    from qradient_b200.sharded import ShardedMcClean, LocalComm
    c = ShardedMcClean(args.n, {"zz": zz}, args.L, LocalComm(args.shards), axes, angles, mode="swap")
    for o in args.opt:
        k, v = o.split("=")
        c.set_option(k, int(v))
else:
    c = McClean(args.n, {"zz": zz}, args.L, axes=axes, angles=angles)
    for o in args.opt:
        k, v = o.split("=")
        c.state.set_option(k, int(v))
for _ in range(args.reps):
    e, g = c.grad_run()
print("E = %.15g" % e)
perf = c.perf if args.shards > 1 else c.state.perf()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in perf.items()})
