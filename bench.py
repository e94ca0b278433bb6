#!/usr/bin/env python
"""bench.py -- McClean grad_run full gradients/s on B200 (BASELINE.json metric) and the fraction of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Default workload at every N: ONE circuit, McClean 30 qubits x 30 layers, ZZ(0,1), default_rng(30) -- the north-star
size (HBM bound: 16 GiB per vector).  N = 1 runs it through `McClean.grad_run`; N > 1 runs the SAME circuit through
`ShardedMcClean.grad_run` with the register sharded on the top log2(N) qubits (NVLink exchanges for the global qubits),
so value(N) / (N value(1)) is a strong-scaling efficiency ("scaling": "strong").  One "step" = one full gradient
(forward sweep, observable, adjoint backward sweep).

    value : gradients/s, device time: CUDA events on the library's stream (N = 1) / barrier + synchronize bracket, max
            over ranks (N > 1), parameters already resident
    e2e   : the same through the public Python API with host buffers: axes / angles H2D and E / grad D2H inside the
            timed region (wall clock)

Extra keys of the same JSON line (BASELINE.json configs; each carries its own parity check against a committed fixture):
    config2_20x20  : McClean 20 x 20 (L2 resident: latency bound), E / grad against tests/golden/gv10 (reference output)
    config3_qaoa26 : QAOA MaxCut 26 qubits p = 10 + 100-shot sampling against tests/golden/gv18 (oracle, full size)
    config4_batch14: 8192 parameter sets of McClean 14 x 14 split over the ranks, 8 of them against tests/golden/gv19
    config5_33x20  : (N = 8) McClean 33 qubits x 20 layers sharded over 8 GPUs

--impl reference times the CPU oracle (numpy port of the reference algorithm; the reference is pure Python and its HEAD
does not import) on all host cores: one single-threaded gradient per core, rates added.  30 x 30 cannot
run on a host (31 history vectors of 16 GiB), so each step is a bounded sample -- a FULL 30-layer gradient at 20
qubits, scaled by 2^10 in the amplitude count (every sweep is linear in it) -- and the 20 x 20 configuration is timed
in full, un-extrapolated, next to it.
"""
import argparse
import glob
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
GOLDEN = os.path.join(ROOT, "tests", "golden")
METRIC = "McClean grad_run full gradients/sec"


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
    "mcclean30": dict(kind="mcclean", n=30, L=30, seed=30, name="McClean 30 qubits x 30 layers grad_run, ZZ(0,1)"),
    "mcclean20": dict(kind="mcclean", n=20, L=20, seed=1234, name="McClean 20 qubits x 20 layers grad_run, ZZ(0,1)"),
    "mcclean26": dict(kind="mcclean", n=26, L=20, seed=26, name="McClean 26 qubits x 20 layers grad_run, ZZ(0,1)"),
    "mcclean3": dict(kind="mcclean", n=3, L=3, seed=1234, name="McClean 3 qubits x 3 layers grad_run, ZZ(0,1)"),
    "qaoa26": dict(kind="qaoa", n=26, L=10, seed=10, name="QAOA MaxCut 3-regular 26 qubits p=10 grad_run"),
    "batch14": dict(kind="batch", n=14, L=14, seed=4, B=8192, name="McClean 14x14, 8192 parameter sets, grad_run_batch"),
    "mcclean33": dict(kind="mcclean", n=33, L=20, seed=5, name="McClean 33 qubits x 20 layers grad_run, ZZ(0,1)"),
}


def mcclean_inputs(w):
    rng = np.random.default_rng(w["seed"])
    return rng.integers(0, 3, (w["L"], w["n"])), rng.uniform(0, 2 * np.pi, (w["L"], w["n"]))


def sched_bytes(n, L, passes):
    """B_sched at the SURVEY.md 8(d) design point: 16 N [2 P (L+1) + 2 + 4 P L]."""
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


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(n):
    """DRAM bytes per backward tile-pass launch from the committed ncu launch list of the current kernels
    (profiles/r*_traffic_bwd_n<n>.json: written by scripts/ncu_traffic.py from an `ncu --metrics dram__bytes_*` pass;
    the file names the library source hash it was captured on).  None when no capture exists for this size."""
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic_bwd_n%d.json" % n)))
    if not cands:
        return None, None
    try:
        d = json.load(open(cands[-1]))
        return float(d["dram_bytes_per_launch"]), "%s (captured on source hash %s; this build: %s)" % (
            os.path.relpath(cands[-1], ROOT), d.get("source_hash"), source_hash())
    except Exception:
        return None, None


def source_hash():
    import hashlib
    h = hashlib.sha1()
    for f in sorted(glob.glob(os.path.join(ROOT, "qradient_b200", "csrc", "*"))):
        h.update(open(f, "rb").read())
    return h.hexdigest()[:12]


NVLINK_SM_PULL_GBPS = 673.0   # measured: SM loads from the peer, both directions busy (profiles/r2_p2pbench.log; copy engines: 777)


def nvlink_block(perf, link_bytes, sec, nl, world, step_seconds):
    """Link-side roofline of a sharded gradient: the exchange passes are the only kernels that cross NVLink; their rate
    (bytes this rank pulls per pass / average duration of the pass, CUDA events on the launching stream) is compared with
    what SM loads get from the link on this hardware."""
    out = {"bytes_in_per_gpu_per_gradient": link_bytes, "achieved_GBps_in_per_gpu_whole_gradient": link_bytes / sec / 1e9,
           "nominal_GBps_per_direction": 900.0, "measured_sm_pull_GBps": NVLINK_SM_PULL_GBPS,
           "peak_source": "profiles/r2_p2pbench.log (scripts/p2pbench.cu, 2 GPUs, both directions)", "step_seconds": step_seconds}
    per_pass = 16.0 * 2.0 ** nl * (world - 1) / world          # bytes of ONE vector crossing the link in one exchange pass
    for key, nv in (("exchange_forward", 1), ("exchange_backward", 2)):
        ms = perf.get("ms_%s_avg" % key, 0.0)
        if ms:
            gbs = nv * per_pass / (ms * 1e-3) / 1e9
            out[key] = {"ms_per_pass": ms, "bytes_per_pass": nv * per_pass, "achieved_GBps": gbs, "frac_of_measured_pull": gbs / NVLINK_SM_PULL_GBPS,
                        "frac_of_nominal": gbs / 900.0}
    return out


# ------------------------------------------------------------------------------------------
# CPU arm
# ------------------------------------------------------------------------------------------
def _oracle_grad_seconds(job):
    """One host core: one FULL oracle gradient of McClean n x L (numpy port of mc_clean.py:47-78)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import qr_oracle as orc
    n, L, seed = job
    rng = np.random.default_rng(seed)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    t = time.perf_counter()
    e, _ = orc.mcclean_grad_run(n, zz01(n), axes, angles)
    return time.perf_counter() - t, e


def cpu_gradients_per_s(n, L, seed, cores):
    """`cores` full gradients side by side (the reference algorithm is single threaded: numpy ufuncs + permutation
    gathers), rates added; the one-core rate if that is higher (the sweeps are memory bound).  Returns
    (gradients/s at n x L, cores used, one-core seconds, E)."""
    t1, e = _oracle_grad_seconds((n, L, seed))
    if cores <= 1:
        return 1.0 / t1, 1, t1, e
    import multiprocessing as mp
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_oracle_grad_seconds, [(n, L, seed)] * cores)
    rate = sum(1.0 / t for t, _ in res)
    if 1.0 / t1 >= rate:
        return 1.0 / t1, 1, t1, e
    return rate, cores, t1, e


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def reference_as_written_20x20():
    """Comparator A: the reference's own circuit_logic/*.py, unmodified, through the adapter of
    tests/golden/make_golden.py -- only where /root/reference (or QRADIENT_REFERENCE) exists, i.e. in the build
    container; the GPU box does not have it."""
    ref = os.environ.get("QRADIENT_REFERENCE", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "qradient")):
        return None
    try:
        sys.path.insert(0, GOLDEN)
        import make_golden as mg
        w = WORKLOADS["mcclean20"]
        axes, angles = mcclean_inputs(w)
        t0 = time.perf_counter()
        c = mg.McClean(w["n"], zz01(w["n"]), w["L"], axes=axes, angles=angles)
        t1 = time.perf_counter()
        e, _ = c.grad_run()
        t2 = time.perf_counter()
        return {"kind": "reference", "cores": 1, "value": 1.0 / (t2 - t1), "unit": "gradients/s", "E": float(e),
                "seconds_per_gradient": t2 - t1, "constructor_seconds": t1 - t0,
                "sample": "reference circuit_logic/mc_clean.py grad_run as written (one thread), one full 20 x 20 gradient"}
    except Exception as exc:   # the adapter needs scipy / tqdm; never fatal for the bench line
        return {"kind": "reference", "error": str(exc)}


def run_reference(args, rank):
    """--impl reference: the CPU oracle on all host cores, on the headline configuration (30 x 30: bounded sample)."""
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = args.cpu_cores if args.cpu_cores > 0 else host_cores()
    n, L = w["n"], w["L"]
    n_cpu = min(n, args.cpu_qubits)
    scale = 2.0 ** (n - n_cpu)
    rates, used, t_one, e_cpu = [], cores, None, None
    t_start = time.perf_counter()
    for i in range(args.warmup + args.steps):
        rate, used, t_one, e_cpu = cpu_gradients_per_s(n_cpu, L, w["seed"], cores)
        if i >= args.warmup:
            rates.append(rate / scale)
        if rates and time.perf_counter() - t_start > args.cpu_budget_s:   # the whole run must end within a few minutes
            break
    val = sum(rates) / len(rates)
    if n_cpu == n:
        desc = "numpy oracle (port of mc_clean.py:47-78, one thread per gradient): FULL %d x %d gradients, un-extrapolated" % (n, L)
    else:
        desc = ("numpy oracle (port of mc_clean.py:47-78, one thread per gradient): FULL %d-layer gradients at %d qubits, "
                "scaled by 2^%d in the amplitude count (a %d-qubit history does not fit a host)" % (L, n_cpu, n - n_cpu, n))
    desc += "; %d gradients side by side, one per host core, rates added (one core alone: %.3g s per sample gradient)" % (used, t_one)
    if len(rates) < args.steps:
        desc += "; %d of the %d requested steps timed (time budget %d s)" % (len(rates), args.steps, args.cpu_budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "gradients/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / val,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128 (f64)",
            "data": "synthetic", "config": {"workload": w["name"], "n_qubits": n, "layers": L},
            "cpu_baseline": {"value": val, "unit": "gradients/s", "cores": used, "kind": "port", "sample": desc,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": val, "unit": "gradients/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.workload == "mcclean30" and not args.no_extras:
        # BASELINE config 2 in full: un-extrapolated 20 x 20 gradients (and the reference as written where it exists)
        w2 = WORKLOADS["mcclean20"]
        r2, u2, t2, e2 = cpu_gradients_per_s(w2["n"], w2["L"], w2["seed"], cores)
        line["config2_20x20"] = {"workload": w2["name"], "value": r2, "unit": "gradients/s", "cores": u2, "kind": "port",
                                 "E": e2, "seconds_per_gradient_one_core": t2,
                                 "sample": "FULL 20 x 20 oracle gradients, un-extrapolated, one per host core"}
        raw = reference_as_written_20x20()
        if raw is not None:
            line["config2_20x20"]["reference_as_written"] = raw
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# parity checks against the committed fixtures (each returns a small dict for the JSON line)
# ------------------------------------------------------------------------------------------
def check_config2(circ_cls, device):
    import torch
    d = np.load(os.path.join(GOLDEN, "gv10_mcclean_20x20.npz"))
    w = WORKLOADS["mcclean20"]
    axes, angles = mcclean_inputs(w)
    assert np.array_equal(axes, d["axes"]) and np.array_equal(angles, d["angles"]), "bench inputs differ from the GV10 recipe"
    circ = circ_cls(w["n"], zz01(w["n"]), w["L"], axes=axes, angles=angles, device=device)
    for _ in range(3):
        e, g = circ.grad_run()
    de, dg = abs(e - float(d["E"])), float(np.abs(g - d["grad"]).max())
    assert de <= 1e-10 and dg <= 1e-10, "20 x 20 differs from the reference output (GV10): |dE| = %g, max|dgrad| = %g" % (de, dg)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    dev, wall, launches, K = [], [], 0, 20
    for _ in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        circ.grad_run()
        wall.append(time.perf_counter() - t0)
        p = circ.perf()
        dev.append(p["ms_total"])
        launches += p["kernel_launches"]
    del flush
    return {"workload": w["name"], "value": K / (sum(dev) * 1e-3), "unit": "gradients/s", "ms_per_gradient": sum(dev) / K,
            "e2e": {"value": K / sum(wall), "unit": "gradients/s", "h2d_bytes_per_step": int(axes.size * 12), "d2h_bytes_per_step": int((axes.size + 1) * 8)},
            "steps": K, "gpu_launches": int(launches), "l2": "256 MB flush between timed steps (state 16 MiB: L2 resident, latency bound)",
            "E": e, "parity_checked": True, "parity": {"fixture": "tests/golden/gv10_mcclean_20x20.npz (reference output)", "abs_dE": de, "max_abs_dgrad": dg, "tol": 1e-10}}


def check_config3(device):
    from qradient_b200.circuit_logic import Qaoa
    from qradient_b200.optimization_problems import MaxCut
    d = np.load(os.path.join(GOLDEN, "gv18_qaoa_config3_26x10.npz"))
    w = WORKLOADS["qaoa26"]
    n, p = w["n"], w["L"]
    q = Qaoa(n, MaxCut(n, edge_set=CONFIG3_EDGES).to_observable(), p, device=device)
    betas, gammas = d["betas"], d["gammas"]
    q.grad_run(betas, gammas)
    dev, K = [], 3
    for _ in range(K):
        e, g = q.grad_run(betas, gammas)
        dev.append(q.perf()["ms_total"])
    scale = float(len(CONFIG3_EDGES))
    de, dg = abs(e - float(d["e"])), float(np.abs(g - d["grad"]).max())
    assert de <= 1e-10 * scale and np.allclose(g, d["grad"], rtol=1e-10, atol=1e-10 * scale), "QAOA-26 differs from the oracle fixture: %g %g" % (de, dg)
    t0 = time.perf_counter()
    q.run_expec_val(betas, gammas)          # psi_final (grad_run leaves the co-state in state.vec)
    t1 = time.perf_counter()
    idx = q.sample_bitstrings(100, d["uniforms"])
    t2 = time.perf_counter()
    same = int(np.sum(idx == d["idx"]))
    assert same == 100, "bitstring indices differ from the oracle fixture (%d of 100 equal)" % same
    return {"workload": w["name"], "value": K / (sum(dev) * 1e-3), "unit": "gradients/s", "ms_per_gradient": sum(dev) / K,
            "sampling": {"shots": 100, "ms_forward_sweep": 1e3 * (t1 - t0), "ms_sampling": 1e3 * (t2 - t1), "indices_equal": same},
            "E": e, "parity_checked": True,
            "parity": {"fixture": "tests/golden/gv18_qaoa_config3_26x10.npz (oracle, full size)", "abs_dE": de, "max_abs_dgrad": dg, "tol": 1e-10 * scale}}


def check_config4(circ_cls, device, rank, world):
    d = np.load(os.path.join(GOLDEN, "gv19_mcclean_config4_14x14.npz"))
    w = WORKLOADS["batch14"]
    n, L, B = w["n"], w["L"], w["B"]
    rng = np.random.default_rng(w["seed"])
    axes, angles = rng.integers(0, 3, (B, L, n)), rng.uniform(0, 2 * np.pi, (B, L, n))
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    circ = circ_cls(n, zz01(n), L, axes=axes[0], angles=angles[0], device=device)
    circ.grad_run_batch(angles[lo:hi], axes[lo:hi])
    dev, wall, K = [], [], 3
    for _ in range(K):
        t0 = time.perf_counter()
        e, g = circ.grad_run_batch(angles[lo:hi], axes[lo:hi])
        wall.append(time.perf_counter() - t0)
        dev.append(circ.perf()["ms_total"])
    worst = 0.0
    for k, b in enumerate(d["indices"]):
        if lo <= b < hi:
            de, dg = abs(e[b - lo] - d["e"][k]), float(np.abs(g[b - lo] - d["grad"][k]).max())
            assert de <= 1e-10 and dg <= 1e-10, "batched 14 x 14, parameter set %d differs from the oracle fixture: %g %g" % (b, de, dg)
            worst = max(worst, de, dg)
    return sum(dev) / K * 1e-3, sum(wall) / K, worst, (hi - lo)


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="mcclean30", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra configurations (config2/3/4/5 keys)")
    ap.add_argument("--cpu-cores", type=int, default=0, help="host cores of the CPU arm (0 = all available)")
    ap.add_argument("--cpu-qubits", type=int, default=20, help="qubits of the CPU sample when the workload does not fit a host")
    ap.add_argument("--cpu-budget-s", type=int, default=150, help="the CPU arm stops repeating its sample after this many seconds")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (see include/qradient_b200.h QR_OPT_*)")
    ap.add_argument("--shard-mode", default=None, help="sharded engine: 'swap' (one NVLink crossing per layer) or 'peer' (round-1 exchange)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.steps is None:
        args.steps = 2 if args.impl == "reference" else {"mcclean30": 5, "mcclean33": 2, "qaoa26": 5, "batch14": 3}.get(args.workload, 20)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        args.warmup = min(args.warmup, 1)
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from qradient_b200.circuit_logic import McClean, Qaoa
    from qradient_b200.optimization_problems import MaxCut
    peak, peak_src = measured_peak()
    opts = [(o.split("=")[0], int(o.split("=")[1])) for o in args.opt]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return t.item()

    n, L = w["n"], w["L"]
    line = None
    sampler = ClockSampler(local_rank)
    # ======================================================================================
    if w["kind"] == "mcclean" and world > 1:
        # ---- one register sharded over all ranks: strong scaling of the headline circuit ----
        from qradient_b200.sharded import ShardedMcClean, TorchDistComm
        axes, angles = mcclean_inputs(w)
        kw = {"mode": args.shard_mode} if args.shard_mode else {}
        sh = ShardedMcClean(n, zz01(n), L, TorchDistComm(), axes, angles, device=local_rank, **kw)
        for k_, v_ in opts:
            sh.set_option(k_, v_)
        for _ in range(max(1, min(args.warmup, 3))):
            sh.grad_run()
        barrier()
        sampler.start()
        t0 = time.perf_counter()
        launches = 0
        for _ in range(args.steps):
            e_sh, g_sh = sh.grad_run()
            launches += sh.perf["kernel_launches"]
        barrier()
        sec = allmax(time.perf_counter() - t0) / args.steps
        clocks = sampler.stop()
        launches = int(allsum(launches))
        perf = sh.perf
        P = perf["sweeps_per_layer"]
        g_ = int(np.log2(world))
        b_hbm = 16.0 * 2.0 ** (n - g_) * (1 + 2 * P * L + 2 + 4 * P * L)          # per GPU, by construction of the schedule
        link = perf["link_bytes"]                                                   # bytes this rank pulled over NVLink per gradient
        line = {"metric": METRIC, "value": 1.0 / sec, "unit": "gradients/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "complex128 (f64)",
                "data": "synthetic", "E": e_sh,
                "config": {"workload": w["name"], "n_qubits": n, "layers": L,
                           "parallelism": "state vector sharded on the top %d qubits over %d GPUs (%s)" % (g_, world, sh.mode),
                           "l2": "no flush: %.1f GiB per vector and GPU, far above the 126 MB L2" % (16.0 * 2.0 ** (n - g_) / 2 ** 30),
                           "sweeps_per_layer": P},
                "e2e": {"value": 1.0 / sec, "unit": "gradients/s", "h2d_bytes_per_step": int(axes.size * 12), "d2h_bytes_per_step": int((L * n + 1) * 8),
                        "note": "the timed call IS the public API call with host buffers (ShardedMcClean.grad_run)"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "whole gradient, per GPU (tile passes + exchange passes)", "achieved": b_hbm / sec / 1e9, "peak": peak,
                             "unit": "GB/s", "frac": b_hbm / sec / 1e9 / peak, "peak_source": peak_src, "traffic": None,
                             "bytes_per_gradient_per_gpu": b_hbm},
                "nvlink": nvlink_block(perf, link, sec, n - g_, world, sh.step_seconds)}
        # ---- parity inside the run: the same circuit on ONE GPU (rank 0), compared with the sharded result ----
        sh.close()
        del sh
        barrier()
        if rank == 0 and not args.no_extras and n <= 30:
            try:
                one = McClean(n, zz01(n), L, axes=axes, angles=angles, device=local_rank)
                e1, g1 = one.grad_run()
                de, dg = abs(e1 - e_sh), float(np.abs(g1 - g_sh).max())
                line["parity_checked"] = bool(de <= 1e-10 and dg <= 1e-10)
                line["parity"] = {"against": "McClean.grad_run on one GPU, same inputs, same run", "E_one_gpu": e1, "abs_dE": de, "max_abs_dgrad": dg,
                                  "tol": 1e-10, "ms_one_gpu": one.perf()["ms_total"]}
                del one
            except Exception as exc:
                line["parity"] = {"error": str(exc)}
        barrier()
        if not args.no_extras:
            # ---- BASELINE config 4: the 8192 parameter sets split over the ranks (no collective on the data path) ----
            try:
                t_dev, t_wall, worst, nb = check_config4(McClean, local_rank, rank, world)
                t_dev, t_wall, worst = allmax(t_dev), allmax(t_wall), allmax(worst)
                line["config4_batch14"] = {"workload": WORKLOADS["batch14"]["name"], "value": WORKLOADS["batch14"]["B"] / t_dev, "unit": "gradients/s",
                                           "e2e": WORKLOADS["batch14"]["B"] / t_wall, "sets_per_gpu": nb, "scaling": "strong", "parity_checked": True,
                                           "parity": {"fixture": "tests/golden/gv19_mcclean_config4_14x14.npz (oracle)", "max_abs_diff": worst, "tol": 1e-10}}
            except Exception as exc:
                line["config4_batch14"] = {"error": str(exc)}
            barrier()
            # ---- BASELINE config 5: 33 qubits x 20 layers on 8 GPUs ----
            if world == 8:
                try:
                    w5 = WORKLOADS["mcclean33"]
                    a5, g5 = mcclean_inputs(w5)
                    sh5 = ShardedMcClean(w5["n"], zz01(w5["n"]), w5["L"], TorchDistComm(), a5, g5, device=local_rank, **kw)
                    sh5.grad_run()
                    barrier()
                    t0 = time.perf_counter()
                    e5, _ = sh5.grad_run()
                    barrier()
                    s5 = allmax(time.perf_counter() - t0)
                    p5 = sh5.perf
                    b5 = 16.0 * 2.0 ** 30 * (1 + 2 * p5["sweeps_per_layer"] * w5["L"] + 2 + 4 * p5["sweeps_per_layer"] * w5["L"])
                    line["config5_33x20"] = {"workload": w5["name"] + ", sharded over 8 GPUs", "value": 1.0 / s5, "unit": "gradients/s", "seconds_per_gradient": s5,
                                             "E": e5, "sweeps_per_layer": p5["sweeps_per_layer"],
                                             "roofline": {"bound": "hbm", "achieved": b5 / s5 / 1e9, "peak": peak, "unit": "GB/s", "frac": b5 / s5 / 1e9 / peak},
                                             "nvlink": nvlink_block(p5, p5["link_bytes"], s5, 30, 8, sh5.step_seconds),
                                             "parity": "no oracle at 33 qubits: the sharded engine is compared with the one-GPU path at 30 qubits in this run and "
                                                       "with the oracle at <= 24 qubits in tests/"}
                    sh5.close()
                except Exception as exc:
                    line["config5_33x20"] = {"error": str(exc)}
                barrier()
        if rank == 0:
            print(json.dumps(line))
        dist.destroy_process_group()
        return

    # ======================================================================================
    # one circuit per rank (N = 1: the headline; N > 1 with a non-default workload: independent replicas, weak scaling)
    units_per_step = 1
    if w["kind"] == "mcclean":
        axes, angles = mcclean_inputs(w)
        circ = McClean(n, zz01(n), L, axes=axes, angles=angles, device=local_rank)
        step = lambda: circ.grad_run()
        h2d, d2h = axes.size * 4 + angles.size * 8, (L * n + 1) * 8
    elif w["kind"] == "qaoa":
        rng = np.random.default_rng(w["seed"])
        gammas, betas = rng.random(L), rng.random(L)
        circ = Qaoa(n, MaxCut(n, edge_set=CONFIG3_EDGES).to_observable(), L, device=local_rank)
        step = lambda: circ.grad_run(betas, gammas)
        h2d, d2h = 2 * L * 8, (2 * L + 1) * 8
    else:
        B = w["B"] // world
        rng = np.random.default_rng(w["seed"])
        axes, angles = rng.integers(0, 3, (w["B"], L, n)), rng.uniform(0, 2 * np.pi, (w["B"], L, n))
        axes, angles = axes[rank * B:(rank + 1) * B], angles[rank * B:(rank + 1) * B]
        circ = McClean(n, zz01(n), L, axes=axes[0], angles=angles[0], device=local_rank)
        step = lambda: circ.grad_run_batch(angles, axes)
        units_per_step = B
        h2d, d2h = axes.size * 4 + angles.size * 8, B * (L * n + 1) * 8
    for k_, v_ in opts:
        circ.state.set_option(k_, v_)
    resident = n <= 22    # 4 buffers of <= 64 MiB: flush the L2 between timed steps; larger registers exceed it by themselves
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if resident else None
    for _ in range(max(args.warmup, 3)):
        res = step()
    barrier()
    sampler.start()
    dev_ms, wall_ms, launches, perf = [], [], 0, None
    for _ in range(args.steps):
        if flush is not None:
            flush.zero_()                   # evict the previous step's state from L2 (untimed)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = step()                        # synchronous: returns after E / grad reached the host
        wall_ms.append(1e3 * (time.perf_counter() - t0))
        perf = circ.perf()
        dev_ms.append(perf["ms_total"])
        launches += perf["kernel_launches"]
    barrier()
    clocks = sampler.stop()
    t_dev, t_wall = allmax(sum(dev_ms)), allmax(sum(wall_ms))
    launches = int(allsum(launches))
    total_units = units_per_step * args.steps * world
    if t_dev <= 0.0:          # gate-at-a-time path (n < 4) records no device events: use the wall clock
        t_dev = t_wall
        perf = dict(perf, ms_total=t_wall / args.steps)
    P = perf["passes_per_layer"]
    bwd_gbs = perf["bwd_pass_bytes"] / (perf["bwd_pass_ms_avg"] * 1e-3) / 1e9 if perf["bwd_pass_ms_avg"] else 0.0
    fwd_gbs = perf["fwd_pass_bytes"] / (perf["fwd_pass_ms_avg"] * 1e-3) / 1e9 if perf["fwd_pass_ms_avg"] else 0.0
    traffic, traffic_src = ncu_traffic(n)
    sched_gbs = perf["algorithmic_bytes"] / (max(perf["ms_total"], 1e-9) * 1e-3) / 1e9
    line = {
        "metric": METRIC if w["kind"] != "qaoa" else "QAOA grad_run full gradients/sec",
        "value": total_units / (t_dev / 1e3), "unit": "gradients/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_dev / args.steps, "higher_is_better": True, "scaling": "strong" if world == 1 else "weak", "vs_baseline": None,
        "dtype": "complex128 (f64)", "data": "synthetic",
        "config": {"workload": w["name"], "n_qubits": n, "layers": L, "units_per_step_per_gpu": units_per_step,
                   "parallelism": "independent circuits per GPU (replicas)" if world > 1 else "1 GPU",
                   "l2": ("256 MB flush between timed steps (state %.0f MiB)" % (16 * 2.0 ** n / 2 ** 20)) if resident
                         else "no flush: %.1f GiB per vector, far above the 126 MB L2" % (16 * 2.0 ** n / 2 ** 30),
                   "passes_per_layer": P, "tile_bits": perf["tile_bits"]},
        "e2e": {"value": total_units / (t_wall / 1e3), "unit": "gradients/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": t_wall / args.steps},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "k_tile12<2,*> / k_tile12_g<2,*> / k_tile12_gs<2> (backward tile passes: psi and lambda; contiguous pass, axis-aware strided passes; qr_tile12.cuh)",
                     "achieved": bwd_gbs, "peak": peak, "unit": "GB/s", "frac": bwd_gbs / peak, "peak_source": peak_src,
                     "bytes_per_launch": perf["bwd_pass_bytes"], "ms_per_launch": perf["bwd_pass_ms_avg"],
                     "traffic": traffic, "traffic_source": traffic_src,
                     "forward_pass": {"achieved": fwd_gbs, "frac": fwd_gbs / peak, "ms_per_launch": perf["fwd_pass_ms_avg"]},
                     "note": "HBM-bound size" if not resident else "state vector is L2-resident: launch latency / L2 bound, not an HBM number"},
        "sched": {"B_sched_bytes": perf["algorithmic_bytes"], "achieved_GBps": sched_gbs, "frac_of_peak": sched_gbs / peak,
                  "B_sched_design_point_bytes": sched_bytes(n, L, 3) if n == 30 else None,
                  "ms_forward": perf["ms_forward"], "ms_observable": perf["ms_observable"], "ms_backward": perf["ms_backward"]},
    }
    if w["kind"] == "mcclean":
        # the result itself: E of the timed call and E of a forward-only run of the same circuit must agree
        e_grad = float(res[0])
        e_fwd = float(circ.run_expec_val())
        line["E"] = e_grad
        line["self_check"] = {"abs_E_grad_minus_E_forward": abs(e_grad - e_fwd), "norm_error": float(abs(circ.state.norm_error())),
                              "note": "no oracle exists at this size (SURVEY.md section 6); oracle parity is asserted on the extra configurations below"}
        assert abs(e_grad - e_fwd) <= 1e-10, "grad_run and run_expec_val disagree on E"
        if world == 1 and not args.no_extras and hasattr(circ, "state"):
            # the timed runs use the per-layer axis-aware plans (QR_OPT_AXIS_PLAN); one more gradient with the static plan
            # (every index bit a tile bit of some pass) must give the same numbers
            g_axis = np.array(res[1])
            circ.state.set_option("axis_plan", 0)
            e_static, g_static = circ.grad_run()
            circ.state.set_option("axis_plan", dict(opts).get("axis_plan", 15))
            line["self_check"]["axis_plan_vs_static_plan"] = {"abs_dE": abs(e_grad - float(e_static)), "max_abs_dgrad": float(np.abs(g_axis - np.array(g_static)).max()),
                                                              "tol": 1e-10, "ms_static_plan": circ.perf()["ms_total"]}
            assert abs(e_grad - float(e_static)) <= 1e-10 and np.abs(g_axis - np.array(g_static)).max() <= 1e-10, "axis-aware and static plans disagree"
    del circ
    if rank == 0 and world == 1 and args.workload == "mcclean30" and not args.no_extras:
        for key, fn in (("config2_20x20", lambda: check_config2(McClean, local_rank)), ("config3_qaoa26", lambda: check_config3(local_rank))):
            try:
                line[key] = fn()
            except AssertionError:
                raise
            except Exception as exc:
                line[key] = {"error": str(exc)}
        try:
            t_dev4, t_wall4, worst, nb = check_config4(McClean, local_rank, 0, 1)
            line["config4_batch14"] = {"workload": WORKLOADS["batch14"]["name"], "value": nb / t_dev4, "unit": "gradients/s", "e2e": nb / t_wall4,
                                       "sets_per_gpu": nb, "parity_checked": True,
                                       "parity": {"fixture": "tests/golden/gv19_mcclean_config4_14x14.npz (oracle)", "max_abs_diff": worst, "tol": 1e-10}}
        except AssertionError:
            raise
        except Exception as exc:
            line["config4_batch14"] = {"error": str(exc)}
        line["parity_checked"] = all(line.get(k, {}).get("parity_checked", False) for k in ("config2_20x20", "config3_qaoa26", "config4_batch14"))
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # a separate interpreter (no CUDA context to fork): the reference arm's own measurement, one gradient per host core
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "1",
               "--warmup", "0", "--cpu-cores", str(args.cpu_cores), "--cpu-qubits", str(args.cpu_qubits), "--no-extras"]
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=900).stdout.strip().splitlines()
            line["cpu_baseline"] = json.loads(out[-1])["cpu_baseline"]
        except Exception as exc:
            line["cpu_baseline"] = {"error": "CPU arm failed: %s" % exc, "kind": "port"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
