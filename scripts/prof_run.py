"""Small driver for ncu captures: one McClean grad_run at (n, L), optional warm-up call."""
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
ap.add_argument("--prefetch", type=int, default=1)
ap.add_argument("--ctas-bwd", type=int, default=1)
ap.add_argument("--ctas-fwd", type=int, default=2)
ap.add_argument("--r-fwd", type=int, default=3)
ap.add_argument("--r-bwd", type=int, default=3)
ap.add_argument("--tile-bits", type=int, default=0)
ap.add_argument("--async-fwd", type=int, default=0)
ap.add_argument("--tile-bits-x", type=int, default=0)
ap.add_argument("--decoupled", type=int, default=0)
ap.add_argument("--min-row-bits", type=int, default=3)
ap.add_argument("--async-bwd", type=int, default=0)
ap.add_argument("--lean", type=int, default=3)
ap.add_argument("--opt", action="append", default=[], help="name=value library option")
args = ap.parse_args()
rng = np.random.default_rng(args.n)
zz = np.full((args.n, args.n), None)
zz[0, 1] = 1.0
c = McClean(args.n, {"zz": zz}, args.L, axes=rng.integers(0, 3, (args.L, args.n)),
            angles=rng.uniform(0, 2 * np.pi, (args.L, args.n)))
c.state.set_option("prefetch", args.prefetch)
c.state.set_option("ctas_per_sm_bwd", args.ctas_bwd)
c.state.set_option("ctas_per_sm_fwd", args.ctas_fwd)
c.state.set_option("reg_bits_fwd", args.r_fwd)
c.state.set_option("reg_bits_bwd", args.r_bwd)
c.state.set_option("tile_bits", args.tile_bits)
c.state.set_option("async_fwd", args.async_fwd)
c.state.set_option("tile_bits_strided", args.tile_bits_x)
c.state.set_option("decoupled", args.decoupled)
c.state.set_option("min_row_bits", args.min_row_bits)
c.state.set_option("async_bwd", args.async_bwd)
c.state.set_option("lean", args.lean)
for o in args.opt:
    k, v = o.split("=")
    c.state.set_option(k, int(v))
for _ in range(args.reps):
    e, g = c.grad_run()
p = c.perf()
print(" ".join("%s=%s" % kv for kv in vars(args).items()))
print("E=%.12f ms_total=%.3f fwd_pass=%.3f ms (%.0f GB/s) bwd_pass=%.3f ms (%.0f GB/s) launches=%d" % (
    e, p["ms_total"], p["fwd_pass_ms_avg"], p["fwd_pass_bytes"] / max(p["fwd_pass_ms_avg"], 1e-9) / 1e6,
    p["bwd_pass_ms_avg"], p["bwd_pass_bytes"] / max(p["bwd_pass_ms_avg"], 1e-9) / 1e6, p["kernel_launches"]))
