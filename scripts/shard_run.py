"""Sharded McClean across ranks (one process per GPU): torchrun --nproc-per-node G scripts/shard_run.py

QR_SHARD_BACKEND=emul runs the same code on CPU (gloo + the host-emulation library) for tests."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", dest="n", type=int, default=31)
ap.add_argument("--layers", dest="L", type=int, default=3)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--check", action="store_true", help="compare with the CPU oracle (small n only)")
ap.add_argument("--tile-bits", type=int, default=0)
ap.add_argument("--check-single", action="store_true", help="compare with the single-GPU path on rank 0")
ap.add_argument("--opt", action="append", default=[], help="library option name=value")
ap.add_argument("--mode", default=None, help="sharded engine: swap | peer (default: auto)")
ap.add_argument("--circuit", default="mcclean", choices=["mcclean", "qaoa"])
args = ap.parse_args()

import torch
import torch.distributed as dist

emul = os.environ.get("QR_SHARD_BACKEND") == "emul"
rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
from qradient_b200 import _lib
if emul:
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import build_emul
    _lib._load_for_testing(build_emul.build())
    dist.init_process_group("gloo")
    device = 0
else:
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
from qradient_b200.sharded import ShardedMcClean, ShardedQaoa, TorchDistComm

n, L = args.n, args.L
if args.circuit == "qaoa":
    # Qaoa.grad_run + sampling on the sharded register, checked against the oracle (small n) on rank 0
    rng = np.random.default_rng(5)
    edges = [(i, i + 1) for i in range(n - 1)] + [(0, n - 1), (0, n // 2)]
    from oracle import qr_oracle as orc
    obs = orc.maxcut_observable(n, edges)
    betas, gammas = rng.random(L), rng.random(L)
    q = ShardedQaoa(n, obs, L, TorchDistComm(), device=device)
    t0 = time.perf_counter()
    e, g = q.grad_run(betas, gammas)
    dt = time.perf_counter() - t0
    e_fwd = q.run_expec_val(betas, gammas)
    u = np.random.RandomState(0).uniform(size=40)
    idx = q.sample_bitstrings(40, u)
    if rank == 0:
        print(json.dumps({"circuit": "qaoa", "n": n, "p": L, "world": world, "E": e, "s_per_gradient": dt, "perf": q.perf}))
        if args.check:
            e_ref, g_ref, psi = orc.qaoa_grad_run(n, obs, betas, gammas, return_state=True)
            ok = abs(e - e_ref) < 1e-10 * len(edges) and np.allclose(g, g_ref, rtol=1e-10, atol=1e-10 * len(edges)) and \
                abs(e_fwd - e_ref) < 1e-10 * len(edges) and np.array_equal(idx, orc.sample_bitstrings(psi, u))
            print("PARITY OK" if ok else "PARITY FAIL")
    q.close()
    dist.destroy_process_group()
    sys.exit(0)
rng = np.random.default_rng(5)
axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
zz = np.full((n, n), None)
zz[0, 1] = 1.0
obs = {"zz": zz, "x": np.array([0.25] + [None] * (n - 1), dtype=object)}
circ = ShardedMcClean(n, obs, L, TorchDistComm(), axes, angles, device=device, mode=args.mode)
if args.tile_bits:
    circ.set_option("tile_bits", args.tile_bits)
for o in args.opt:
    k_, v_ = o.split("=")
    circ.set_option(k_, int(v_))
times = []
for _ in range(args.reps):
    dist.barrier()
    t0 = time.perf_counter()
    e, g = circ.grad_run()
    dist.barrier()
    times.append(time.perf_counter() - t0)
if rank == 0:
    out = {"n": n, "L": L, "world": world, "mode": circ.mode, "perf": circ.perf, "E": e, "grad_norm": float(np.linalg.norm(g)), "s_per_gradient": min(times),
           "bytes_sched_per_gpu": 16.0 * 2.0 ** n / world * (1 + 2 * 3 * L + 2 + 4 * 3 * L),
           "step_seconds": {k: round(v, 4) for k, v in circ.step_seconds.items()}, "opts": args.opt,
           "nvlink_bytes_per_direction_per_global_vector_step": 16.0 * 2.0 ** n / world * (world - 1) / world,
           "nvlink_bytes_per_direction_per_gradient_counted": getattr(circ, "link_bytes", None)}
    print(json.dumps(out))
    if args.check:
        from oracle import qr_oracle as orc
        e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
        ok = abs(e - e_ref) < 1.25e-10 and np.allclose(g, g_ref, rtol=1e-10, atol=1.25e-10)
        print("PARITY OK" if ok else "PARITY FAIL %g %g" % (abs(e - e_ref), np.abs(g - g_ref).max()))
if args.check_single:
    if rank == 0:
        from qradient_b200.circuit_logic import McClean
        single = McClean(n, obs, L, axes=axes, angles=angles, device=device)
        t0 = time.perf_counter()
        e1, g1 = single.grad_run()
        t1 = time.perf_counter() - t0
        ok = abs(e - e1) < 1.25e-10 and np.allclose(g, g1, rtol=1e-10, atol=1.25e-10)
        print("SINGLE-GPU %.3f s; %s |dE|=%.2e max|dg|=%.2e" % (t1, "PARITY OK" if ok else "PARITY FAIL", abs(e - e1), np.abs(g - g1).max()))
        del single
    dist.barrier()
circ.close()
dist.destroy_process_group()
