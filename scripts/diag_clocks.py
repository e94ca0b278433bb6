"""Diagnostic: McClean grad_run at (n, L) with clocks / power sampled during the run (nvidia-smi -lms 20)."""
import argparse
import os
import subprocess
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qradient_b200.circuit_logic import McClean  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=30)
ap.add_argument("--L", type=int, default=6)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--opt", action="append", default=[], help="name=value library option")
args = ap.parse_args()
rng = np.random.default_rng(args.n)
zz = np.full((args.n, args.n), None)
zz[0, 1] = 1.0
c = McClean(args.n, {"zz": zz}, args.L, axes=rng.integers(0, 3, (args.L, args.n)), angles=rng.uniform(0, 2 * np.pi, (args.L, args.n)))
for o in args.opt:
    k, v = o.split("=")
    c.state.set_option(k, int(v))
c.grad_run()   # warm-up: allocations
rows = []
proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active",
                         "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
threading.Thread(target=lambda: [rows.append((time.time(), l.strip())) for l in proc.stdout], daemon=True).start()
time.sleep(0.3)
marks = []
for _ in range(args.reps):
    t0 = time.time()
    e, g = c.grad_run()
    marks.append((t0, time.time()))
    p = c.perf()
    print("n=%d L=%d %s E=%.12f ms_total=%.2f fwd_pass=%.3f ms (%.0f GB/s) bwd_pass=%.3f ms (%.0f GB/s)" % (
        args.n, args.L, args.opt, e, p["ms_total"], p["fwd_pass_ms_avg"], p["fwd_pass_bytes"] / max(p["fwd_pass_ms_avg"], 1e-9) / 1e6,
        p["bwd_pass_ms_avg"], p["bwd_pass_bytes"] / max(p["bwd_pass_ms_avg"], 1e-9) / 1e6))
time.sleep(0.2)
proc.terminate()
for t, l in rows:
    tag = "RUN " if any(a <= t <= b for a, b in marks) else "idle"
    print(tag, "%.3f" % (t - rows[0][0]), l)
