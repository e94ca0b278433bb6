"""A/B of the axis-aware plans (QR_OPT_AXIS_PLAN) on one GPU: same circuit, option 0 / 1 / 2, device times from the library's
own CUDA events, results compared."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qradient_b200.circuit_logic import McClean  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cases", default="30x6,28x6,26x10,24x10,22x20,20x20")
ap.add_argument("--modes", default="0,1,2")
ap.add_argument("--tile-bits", default="0,12")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
for case in args.cases.split(","):
    n, L = (int(v) for v in case.split("x"))
    rng = np.random.default_rng(1234)
    zz = np.full((n, n), None)
    zz[0, 1] = 1.0
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    c = McClean(n, {"zz": zz}, L, axes=axes, angles=angles)
    ref = None
    for tb in (int(v) for v in args.tile_bits.split(",")):
        for mode in (int(v) for v in args.modes.split(",")):
            c.state.set_option("tile_bits", tb)
            c.state.set_option("axis_plan", mode)
            best = None
            for _ in range(args.reps):
                e, g = c.grad_run()
                p = c.state.perf()
                if best is None or p["ms_total"] < best["ms_total"]:
                    best = p
            if ref is None:
                ref = (e, g)
            print("n=%d L=%d tile_bits=%d axis_plan=%d: total %.3f ms (fwd %.3f, bwd %.3f), launches %d, fwd pass %.3f ms, bwd pass %.3f ms, "
                  "|dE| %.1e, max|dg| %.1e" % (n, L, tb, mode, best["ms_total"], best["ms_forward"], best["ms_backward"], best["kernel_launches"],
                                               best["fwd_pass_ms_avg"], best["bwd_pass_ms_avg"], abs(e - ref[0]), np.abs(g - ref[1]).max()), flush=True)
    del c
