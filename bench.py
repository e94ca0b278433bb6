#!/usr/bin/env python
"""bench.py -- McClean grad_run full gradients/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

Workloads (synthetic random-parameter circuits, explicit seeded inputs):
    mcclean20  : McClean 20 qubits x 20 layers, ZZ(0,1), default_rng(1234)   (BASELINE config 2, default)
    mcclean30  : McClean 30 qubits x 30 layers, ZZ(0,1), default_rng(30)     (north-star HBM target)
    mcclean3   : McClean 3 x 3 README example                                  (BASELINE config 1)
    qaoa26     : QAOA MaxCut 3-regular 26 qubits p=10                          (BASELINE config 3)
    batch14    : McClean 14 x 14, 8192 parameter sets (split across ranks)     (BASELINE config 4)

One "step" = one full gradient (forward sweep, observable, adjoint backward sweep) of one
circuit (batch14: of the rank's share of the 8192 parameter sets).  `value` is the device-timed
throughput (CUDA events on the library's stream, parameters already uploaded);  `e2e` is the
same metric through the public Python API with host buffers: axes/angles H2D and E/grad D2H
inside the timed region.  N > 1: every rank runs its own independent circuits (parameter-set
sharding, no data-path collective) -> weak scaling; time = max over ranks.

--impl reference times the CPU oracle (numpy port of the reference algorithm; the reference is
pure Python and does not import at HEAD) on a bounded sample of the same workload:
the algorithm is single threaded, so every host core runs its own gradient and the rates are added.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------
def zz01(n):
    m = np.full((n, n), None)
    m[0, 1] = 1.0
    return {"zz": m}


CONFIG3_EDGES = [(0, 3), (0, 24), (0, 25), (1, 10), (1, 21), (1, 22), (2, 5), (2, 6), (2, 11), (3, 16), (3, 17), (4, 7),
                 (4, 12), (4, 20), (5, 6), (5, 9), (6, 23), (7, 13), (7, 16), (8, 13), (8, 15), (8, 18), (9, 14), (9, 20),
                 (10, 11), (10, 22), (11, 15), (12, 17), (12, 21), (13, 18), (14, 20), (14, 22), (15, 19), (16, 17),
                 (18, 21), (19, 24), (19, 25), (23, 24), (23, 25)]   # SURVEY.md appendix C (networkx seed 26)

WORKLOADS = {
    "mcclean20": dict(kind="mcclean", n=20, L=20, seed=1234, name="McClean 20 qubits x 20 layers grad_run, ZZ(0,1)"),
    "mcclean30": dict(kind="mcclean", n=30, L=30, seed=30, name="McClean 30 qubits x 30 layers grad_run, ZZ(0,1)"),
    "mcclean26": dict(kind="mcclean", n=26, L=20, seed=26, name="McClean 26 qubits x 20 layers grad_run, ZZ(0,1)"),
    "mcclean3": dict(kind="mcclean", n=3, L=3, seed=1234, name="McClean 3 qubits x 3 layers grad_run, ZZ(0,1)"),
    "qaoa26": dict(kind="qaoa", n=26, L=10, seed=10, name="QAOA MaxCut 3-regular 26 qubits p=10 grad_run"),
    "batch14": dict(kind="batch", n=14, L=14, seed=4, B=8192, name="McClean 14x14, 8192 parameter sets, grad_run_batch"),
    "shard": dict(kind="shard", n=30, L=20, seed=5, name="McClean (30+log2 G) qubits x 20 layers, state sharded over G GPUs"),
}


def mcclean_inputs(w, rank=0):
    rng = np.random.default_rng(w["seed"] + 1000 * rank)
    return rng.integers(0, 3, (w["L"], w["n"])), rng.uniform(0, 2 * np.pi, (w["L"], w["n"]))


def sched_bytes(n, L, passes):
    """B_sched of SURVEY.md 8(d): 16 N [2 P (L+1) + 2 + 4 P L]  (the init pass writes only)."""
    return 16.0 * 2.0 ** n * (2 * passes * (L + 1) + 2 + 4 * passes * L)


# ------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax = float(r[2])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# per-launch DRAM traffic of the backward tile pass at n = 30 (ncu dram__bytes_read.sum + dram__bytes_write.sum,
# averaged over the 9 backward launches of profiles/r1_launches_mcclean30_L3_tile12_final.csv; algorithmic: 68.72e9)
TRAFFIC20_BWD = 33.64e6   # n = 20 (L2 resident): ncu cold-cache capture, reads 33.6 MB, writes stay in L2 (profiles/r1_launches_mcclean20_dram.csv)
TRAFFIC30_BWD = 69.3e9
TRAFFIC30_SRC = ("ncu dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 9 backward launches in "
                 "profiles/r1_launches_mcclean30_L3_tile12_final.csv (reads 35.0 GB incl. ~2 % L2-prefetch over-fetch, writes 34.3 GB)")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
def cpu_oracle_time(w, layers_sample):
    """Oracle (numpy port of mc_clean.py:47-78 / qaoa.py:40-70) on a bounded sample; returns
    (seconds for the full workload, extrapolated linearly in the layer count; description)."""
    from oracle import qr_oracle as orc
    n, L = w["n"], w["L"]
    if w["kind"] in ("mcclean", "batch"):
        n_cpu = min(n, 22)
        Ls = min(L, layers_sample)
        axes, angles = mcclean_inputs(dict(w, n=n_cpu))
        def run(layers):
            t = time.perf_counter()
            orc.mcclean_grad_run(n_cpu, zz01(n_cpu), axes[:layers], angles[:layers])
            return time.perf_counter() - t
        t1 = run(1)
        ts = run(Ls) if Ls > 1 else t1
        slope = (ts - t1) / (Ls - 1) if Ls > 1 else t1
        full = (t1 + slope * (L - 1)) * 2.0 ** (n - n_cpu)   # cost is affine in the layer count, x2 per qubit
        desc = "numpy oracle (1 thread), gradients with 1 and %d of %d layers at n=%d, extrapolated affinely in layers" % (Ls, L, n_cpu)
        if n != n_cpu:
            desc += " and x2^%d in qubits" % (n - n_cpu)
        if w["kind"] == "batch":
            full *= w["B"]
            desc += " x%d parameter sets" % w["B"]
        return full, desc
    n_cpu = min(n, 20)
    rng = np.random.default_rng(w["seed"])
    gammas, betas = rng.random(L), rng.random(L)
    edges = [e for e in CONFIG3_EDGES if e[0] < n_cpu and e[1] < n_cpu]
    Ls = min(L, max(1, layers_sample // 2))
    t = time.perf_counter()
    orc.qaoa_grad_run(n_cpu, orc.maxcut_observable(n_cpu, edges), betas[:Ls], gammas[:Ls])
    dt = time.perf_counter() - t
    return dt * (L / Ls) * 2.0 ** (n - n_cpu), "numpy oracle, %d of %d layers at n=%d (induced subgraph), scaled" % (Ls, L, n_cpu)


def _cpu_worker(job):
    """One host core: the bounded oracle sample of workload `job[0]` (runs in a spawned process)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    name, layers = job
    return cpu_oracle_time(WORKLOADS[name], layers)


def cpu_oracle_throughput(name, layers_sample, cores):
    """The reference algorithm is single threaded by construction (numpy ufuncs + permutation gathers), so "all the host
    threads it can use" = `cores` independent gradients side by side, one per core, exactly how the GPU arm scales over
    GPUs (independent circuits, no exchange).  Returns (units per second over all cores, cores used, description)."""
    w = WORKLOADS[name]
    units = w.get("B", 1)
    if cores <= 1:
        full, desc = cpu_oracle_time(w, layers_sample)
        return units / full, 1, desc
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(name, layers_sample)] * cores)
    rate = sum(units / full for full, _ in res)
    # the sweeps are memory bound: on a host whose cores share one memory system the concurrent runs can add up to
    # LESS than one undisturbed core -- report whichever is better for the CPU
    full1, desc1 = cpu_oracle_time(w, layers_sample)
    if units / full1 >= rate:
        return units / full1, 1, desc1 + "; best of this and %d concurrent gradients (%.3g/s in total)" % (cores, rate)
    return rate, cores, res[0][1] + "; %d such gradients concurrently, one per host core, rates added (one core alone: %.3g/s)" % (
        cores, units / full1)


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def run_reference(args, w, rank, world):
    """--impl reference: the CPU oracle on all host cores (one single-threaded gradient per core)."""
    if rank != 0:
        return
    cores = args.cpu_cores if args.cpu_cores > 0 else host_cores()
    rates = []
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        rate, used, desc = cpu_oracle_throughput(args.workload, args.cpu_layers, cores)
        if i >= args.warmup:
            rates.append(rate)
        # every step is the same bounded sample; the whole run must end within a few minutes whatever --steps says
        if rates and time.perf_counter() - t_start > args.cpu_budget_s:
            break
    val = sum(rates) / len(rates)
    sec = w.get("B", 1) / val
    if len(rates) < args.steps:
        desc += "; %d of the %d requested steps timed (time budget %d s)" % (len(rates), args.steps, args.cpu_budget_s)
    line = {"impl": "reference", "metric": "McClean grad_run full gradients/sec", "value": val, "unit": "gradients/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
            "data": "synthetic", "config": {"workload": w["name"], "n_qubits": w["n"], "layers": w["L"]},
            "cpu_baseline": {"value": val, "unit": "gradients/s", "cores": used, "kind": "port", "sample": desc,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": val, "unit": "gradients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="mcclean20", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-layers", type=int, default=4, help="layers of the CPU sample (cpu_baseline / reference arm)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cores", type=int, default=0, help="host cores of the CPU arm (0 = all available)")
    ap.add_argument("--cpu-budget-s", type=int, default=150, help="the CPU arm stops repeating its sample after this many seconds")
    ap.add_argument("--prefetch", type=int, default=None)
    ap.add_argument("--tile-bits", type=int, default=None)
    ap.add_argument("--ctas-bwd", type=int, default=None)
    ap.add_argument("--ctas-fwd", type=int, default=None)
    ap.add_argument("--batch-chunk-mb", type=int, default=None)
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (see include/qradient_b200.h QR_OPT_*)")
    ap.add_argument("--hbm-target", type=int, default=1, help="also measure the 30x30 HBM-bound target (N=1 only)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else {"mcclean30": 3, "qaoa26": 5, "batch14": 3, "shard": 2}.get(args.workload, 20)
    if args.impl == "reference":
        args.warmup = min(args.warmup, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from qradient_b200.circuit_logic import McClean, Qaoa
    from qradient_b200.optimization_problems import MaxCut

    n, L = w["n"], w["L"]
    units_per_step = 1
    if w["kind"] == "shard" and world > 1:
        # BASELINE config 5: one register sharded over all ranks (NVLink P2P for the global qubits)
        from qradient_b200.sharded import ShardedMcClean, TorchDistComm
        n = w["n"] + int(np.log2(world))
        axes, angles = mcclean_inputs(dict(w, n=n))
        sh = ShardedMcClean(n, zz01(n), L, TorchDistComm(), axes, angles, device=local_rank)
        for _ in range(max(args.warmup - 2, 1)):
            sh.grad_run()
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e_sh, _ = sh.grad_run()
        dist.barrier(); torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if rank == 0:
            sec = dt.item() / args.steps
            P = 3
            bsched = 16.0 * 2.0 ** n / world * (1 + 2 * P * L + 2 + 4 * P * L)
            nvl = getattr(sh, "link_bytes", 0.0) or 2 * 16.0 * 2.0 ** n / world * (world - 1) / world * 3 * L   # bytes per direction per GPU (counted by the library: peer loads in + peer stores out, less for Rz on global qubits)
            peak, peak_src = measured_peak()
            print(json.dumps({"metric": "McClean grad_run full gradients/sec", "value": 1.0 / sec, "unit": "gradients/s",
                              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec,
                              "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "complex128 (f64)",
                              "data": "synthetic", "E": e_sh,
                              "config": {"workload": "McClean %d qubits x %d layers, state sharded over %d GPUs" % (n, L, world),
                                         "n_qubits": n, "layers": L, "parallelism": "state sharded on the top %d qubits" % int(np.log2(world))},
                              "e2e": {"value": 1.0 / sec, "unit": "gradients/s", "h2d_bytes_per_step": int(axes.size * 12),
                                      "d2h_bytes_per_step": int((L * n + 1) * 8)},
                              "roofline": {"bound": "hbm", "achieved": bsched / sec / 1e9, "peak": peak, "unit": "GB/s",
                                           "frac": bsched / sec / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                                           "note": "whole-gradient B_sched per GPU; NVLink bytes/direction/GPU = %.3g" % nvl}}))
        sh.close()
        dist.destroy_process_group()
        return
    if w["kind"] == "shard":
        w = dict(w, kind="mcclean")
    if w["kind"] == "mcclean":
        axes, angles = mcclean_inputs(w, rank)
        circ = McClean(n, zz01(n), L, axes=axes, angles=angles, device=local_rank)
        step = lambda: circ.grad_run()
        h2d, d2h = axes.size * 4 + angles.size * 8, (L * n + 1) * 8
    elif w["kind"] == "qaoa":
        rng = np.random.default_rng(w["seed"] + 1000 * rank)
        gammas, betas = rng.random(L), rng.random(L)
        circ = Qaoa(n, MaxCut(n, edge_set=CONFIG3_EDGES).to_observable(), L, device=local_rank)
        step = lambda: circ.grad_run(betas, gammas)
        h2d, d2h = 2 * L * 8, (2 * L + 1) * 8
    else:
        B = w["B"] // world
        rng = np.random.default_rng(w["seed"] + 1000 * rank)
        axes, angles = rng.integers(0, 3, (B, L, n)), rng.uniform(0, 2 * np.pi, (B, L, n))
        circ = McClean(n, zz01(n), L, axes=axes[0], angles=angles[0], device=local_rank)
        step = lambda: circ.grad_run_batch(angles, axes)
        units_per_step = B
        h2d, d2h = axes.size * 4 + angles.size * 8, B * (L * n + 1) * 8
    for name, val in (("prefetch", args.prefetch), ("tile_bits", args.tile_bits), ("ctas_per_sm_bwd", args.ctas_bwd),
                      ("ctas_per_sm_fwd", args.ctas_fwd), ("batch_chunk_mb", args.batch_chunk_mb)):
        if val is not None:
            circ.state.set_option(name, val)
    for o in args.opt:
        k_, v_ = o.split("=")
        circ.state.set_option(k_, int(v_))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, wall_ms, launches = [], [], 0
    perf = None
    for _ in range(args.steps):
        flush.zero_()                       # evict the previous step's state from L2 (untimed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()                              # synchronous: returns after E/grad reached the host
        wall_ms.append(1e3 * (time.perf_counter() - t0))
        perf = circ.perf()
        dev_ms.append(perf["ms_total"])
        launches += perf["kernel_launches"]
    barrier()
    clocks = sampler.stop()
    t_dev, t_wall = sum(dev_ms), sum(wall_ms)
    if world > 1:
        tt = torch.tensor([t_dev, t_wall], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_dev, t_wall = tt.tolist()
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    total_units = units_per_step * args.steps * world
    if t_dev <= 0.0:          # gate-at-a-time path (n < 4) records no device events: use the wall clock
        t_dev = t_wall
        perf = dict(perf, ms_total=t_wall / args.steps)
    value = total_units / (t_dev / 1e3)
    e2e = total_units / (t_wall / 1e3)
    peak, peak_src = measured_peak()
    P = perf["passes_per_layer"]
    bwd_gbs = perf["bwd_pass_bytes"] / (perf["bwd_pass_ms_avg"] * 1e-3) / 1e9 if perf["bwd_pass_ms_avg"] else 0.0
    line = {
        "metric": "McClean grad_run full gradients/sec" if w["kind"] != "qaoa" else "QAOA grad_run full gradients/sec",
        "value": value, "unit": "gradients/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": w["name"], "n_qubits": n, "layers": L, "units_per_step_per_gpu": units_per_step,
                   "parallelism": "independent parameter sets per GPU" if world > 1 else "1 GPU",
                   "l2": "256 MB flush between timed steps (state %.0f MiB)" % (16 * 2.0 ** n / 2 ** 20),
                   "passes_per_layer": P, "tile_bits": perf["tile_bits"]},
        "e2e": {"value": e2e, "unit": "gradients/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": t_wall / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_tile12<2,*> (backward tile pass: psi and lambda, qr_tile12.cuh)",
                     "achieved": bwd_gbs, "peak": peak, "unit": "GB/s", "frac": bwd_gbs / peak, "peak_source": peak_src,
                     "bytes_per_launch": perf["bwd_pass_bytes"], "ms_per_launch": perf["bwd_pass_ms_avg"],
                     "traffic": TRAFFIC20_BWD if (n == 20 and w["kind"] == "mcclean") else None,
                     "note": "state vector is L2-resident at n<=21 (2 x %.0f MiB): fraction of the HBM peak is reported "
                             "but launch latency / L2 bound; see hbm_target for the HBM-bound size" % (16 * 2.0 ** n / 2 ** 20)
                     if n <= 21 else "HBM-bound size"},
        "sched": {"B_sched_bytes": perf["algorithmic_bytes"], "achieved_GBps": perf["algorithmic_bytes"] / (max(perf["ms_total"], 1e-9) * 1e-3) / 1e9,
                  "frac_of_peak": perf["algorithmic_bytes"] / (max(perf["ms_total"], 1e-9) * 1e-3) / 1e9 / peak,
                  "ms_forward": perf["ms_forward"], "ms_observable": perf["ms_observable"], "ms_backward": perf["ms_backward"]},
    }
    # ---- BASELINE config 3: "... exact grad_run plus 100-shot sampling": the sampling leg, timed after the gradient steps ----
    if w["kind"] == "qaoa" and rank == 0:
        u = np.random.RandomState(0).uniform(size=100)
        circ.run_expec_val(betas, gammas)
        circ.sample_cost(100, u)                       # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e_fwd = circ.run_expec_val(betas, gammas)      # psi_final (grad_run leaves the co-state in state.vec)
        t1 = time.perf_counter()
        mean_cost = circ.sample_cost(100, u)           # prefix-sum sampler + H gather on the device
        t2 = time.perf_counter()
        line["sampling"] = {"shots": 100, "ms_forward_sweep": 1e3 * (t1 - t0), "ms_sampling": 1e3 * (t2 - t1),
                            "mean_cost_of_samples": float(mean_cost), "E_exact": float(e_fwd), "uniforms": "RandomState(0).uniform(size=100)"}
    # ---- the HBM-bound north-star size, measured in the same run (N = 1, default workload only) ----
    if rank == 0 and world == 1 and args.hbm_target and args.workload == "mcclean20":
        try:
            del circ
            w30 = WORKLOADS["mcclean30"]
            a30, g30 = mcclean_inputs(w30)
            c30 = McClean(30, zz01(30), 30, axes=a30, angles=g30, device=local_rank)
            for name, val in (("prefetch", args.prefetch), ("ctas_per_sm_bwd", args.ctas_bwd), ("ctas_per_sm_fwd", args.ctas_fwd)):
                if val is not None:
                    c30.state.set_option(name, val)
            for o in args.opt:
                k_, v_ = o.split("=")
                c30.state.set_option(k_, int(v_))
            c30.grad_run()
            t0 = time.perf_counter()
            e30, _ = c30.grad_run()
            wall30 = time.perf_counter() - t0
            p30 = c30.perf()
            b30 = p30["bwd_pass_bytes"] / (p30["bwd_pass_ms_avg"] * 1e-3) / 1e9
            f30 = p30["fwd_pass_bytes"] / (p30["fwd_pass_ms_avg"] * 1e-3) / 1e9
            line["hbm_target"] = {
                "workload": w30["name"], "gradients_per_s": 1e3 / p30["ms_total"], "e2e_gradients_per_s": 1.0 / wall30,
                "ms_per_gradient": p30["ms_total"], "E": e30, "passes_per_layer": p30["passes_per_layer"],
                "roofline": {"bound": "hbm", "kernel": "k_tile12<2,false,*> backward (qr_tile12.cuh)", "achieved": b30, "peak": peak, "unit": "GB/s",
                             "frac": b30 / peak, "bytes_per_launch": p30["bwd_pass_bytes"], "ms_per_launch": p30["bwd_pass_ms_avg"],
                             "traffic": TRAFFIC30_BWD, "traffic_source": TRAFFIC30_SRC},
                "forward_pass": {"achieved": f30, "frac": f30 / peak, "ms_per_launch": p30["fwd_pass_ms_avg"]},
                "sched": {"B_sched_bytes": p30["algorithmic_bytes"], "achieved_GBps": p30["algorithmic_bytes"] / (p30["ms_total"] * 1e-3) / 1e9,
                          "frac_of_peak": p30["algorithmic_bytes"] / (p30["ms_total"] * 1e-3) / 1e9 / peak}}
            del c30
        except Exception as exc:
            line["hbm_target"] = {"error": str(exc)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # a separate interpreter (no CUDA context to fork): the reference arm's own measurement, one gradient per host core
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
               "--warmup", "0", "--cpu-layers", str(args.cpu_layers), "--cpu-cores", str(args.cpu_cores)]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=900).stdout.strip().splitlines()
            line["cpu_baseline"] = json.loads(out[-1])["cpu_baseline"]
        except Exception as exc:
            full, desc = cpu_oracle_time(w, args.cpu_layers)
            line["cpu_baseline"] = {"value": units_per_step / full, "unit": "gradients/s", "cores": 1, "kind": "port",
                                    "sample": desc + " (multi-core run failed: %s)" % exc, "host_cores_available": os.cpu_count()}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
