"""State sharding over ranks (qradient_b200/sharded.py): virtual shards in one process on both
backends, and a real world_size-2 `gloo` run on CPU (two processes, IPC-mapped shard buffers)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from backends import backend  # noqa: F401
from conftest import assert_parity, obs_scale
from oracle import qr_oracle as orc
from qradient_b200.sharded import ShardedMcClean, ShardedQaoa, LocalComm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def mixed_obs(n):
    m = np.full((n, n), None)
    m[0, 1] = 1.0
    m[1, n - 1] = 0.5
    return {"zz": m, "x": np.array([0.3] + [None] * (n - 1), dtype=object),
            "y": np.array([None, 0.2] + [None] * (n - 3) + [0.7], dtype=object),
            "z": np.array([None, -0.4] + [None] * (n - 2), dtype=object)}


@pytest.mark.parametrize("n,L,G,tile_bits", [(7, 2, 2, 12), (8, 3, 4, 12), (9, 2, 8, 12), (10, 2, 4, 5), (9, 3, 2, 4),
                                             (12, 2, 16, 6)])
def test_virtual_shards_vs_oracle(backend, n, L, G, tile_bits):
    rng = np.random.default_rng(n * 10 + G)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
    c = ShardedMcClean(n, obs, L, LocalComm(G), axes, angles)
    try:
        c.set_option("tile_bits", tile_bits)
        e, g = c.grad_run()
        assert_parity(e, g, e_ref, g_ref, obs_scale(obs), 1e-10)
        assert abs(c.run_expec_val() - e_ref) <= 1e-10 * obs_scale(obs)
        c.angles = angles + 0.1          # parameters can be re-assigned between calls
        e2, g2 = c.grad_run()
        e_ref2, g_ref2 = orc.mcclean_grad_run(n, obs, axes, angles + 0.1)
        assert_parity(e2, g2, e_ref2, g_ref2, obs_scale(obs), 1e-10)
    finally:
        c.close()


@pytest.mark.parametrize("n,L,G,peeled", [(14, 3, 2, False), (15, 2, 2, True), (16, 3, 4, False), (17, 3, 4, True), (18, 3, 8, True)])
def test_swap_engine_virtual_shards_vs_oracle(backend, n, L, G, peeled):
    """Swap engine (qr_shard.cuh): exchange passes whose loads come from the peer shards and whose stores are local, the
    layout alternating between 'qubits 0..g-1 on the rank bits' and 'local bits [sigma, sigma+g) on the rank bits', ladder
    passes with cross-shard tiles in the swapped layout; x / y observable terms on a rank-held qubit and on the lowest bit."""
    if backend == "emul" and n >= 16:
        L = 2          # the host emulation runs every thread of every CTA as a fiber: keep the CPU tier short
    rng = np.random.default_rng(n * 10 + G)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    obs = mixed_obs(n)
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
    c = ShardedMcClean(n, obs, L, LocalComm(G), axes, angles, mode="swap")
    try:
        assert c.mode == "swap"
        e, g = c.grad_run()
        assert_parity(e, g, e_ref, g_ref, obs_scale(obs), 1e-10)
        assert c.perf["sweeps_per_layer"] == 3 and c.perf["kernel_launches"] > 0
        # every exchanged amplitude crosses the link once: (G-1)/G of a shard per exchange pass (+ the cross-shard ladder tiles)
        n_loc = 2 ** (n - int(np.log2(G)))
        exchange = 3 * L * 16.0 * n_loc * (G - 1) / G
        assert exchange <= c.link_bytes <= 2.0 * exchange
        # where the geometry allows, the CNOT that would make a ladder pass read another shard is folded into the
        # neighbouring exchange pass: exactly one crossing per exchange, nothing else on the link
        assert (c.link_bytes == exchange) == peeled
        assert abs(c.run_expec_val() - e_ref) <= 1e-10 * obs_scale(obs)
        if n > 16 and backend == "emul":
            return
        c.angles = angles + 0.1          # parameters can be re-assigned between calls
        e2, g2 = c.grad_run()
        e_ref2, g_ref2 = orc.mcclean_grad_run(n, obs, axes, angles + 0.1)
        assert_parity(e2, g2, e_ref2, g_ref2, obs_scale(obs), 1e-10)
        if n > 15 and backend == "emul":
            return
        # the round-1 engine on the same register
        p = ShardedMcClean(n, obs, L, LocalComm(G), axes, angles + 0.1, mode="peer")
        e3, g3 = p.grad_run()
        p.close()
        assert_parity(e3, g3, e2, g2, obs_scale(obs), 1e-12)
    finally:
        c.close()


@pytest.mark.parametrize("n,p,G,weighted", [(14, 3, 2, False), (15, 2, 2, True), (16, 2, 4, False), (18, 3, 8, False)])
def test_sharded_qaoa_and_sampling_vs_oracle(backend, n, p, G, weighted):
    """Qaoa.grad_run / run_expec_val (qaoa.py:23-70) and inverse-CDF bitstring sampling (qaoa.py:196-198) on a sharded register:
    the diagonal phase by the per-layout H tables (phase look-up table for integer weights, sincos otherwise), the beta /
    gamma reductions of all passes, the shard-local scans + gathered totals of the sampler (indices equal to the oracle's)."""
    if backend == "emul" and n >= 18:
        p = 2
    rng = np.random.default_rng(n + G)
    edges = [(i, i + 1) for i in range(n - 1)] + [(0, n - 1), (0, n // 2), (1, n - 2)]
    if weighted:
        zz = np.full((n, n), None)
        for (a, b) in edges:
            zz[a, b] = float(rng.uniform(0.5, 1.5))
        obs = {"zz": zz, "z": np.array([0.3, None, -0.7] + [None] * (n - 3), dtype=object)}
        scale = obs_scale(obs)
    else:
        obs = orc.maxcut_observable(n, edges)
        scale = float(len(edges))
    betas, gammas = rng.random(p), rng.random(p)
    e_ref, g_ref, psi = orc.qaoa_grad_run(n, obs, betas, gammas, return_state=True)
    q = ShardedQaoa(n, obs, p, LocalComm(G))
    try:
        e, g = q.grad_run(betas, gammas)
        assert_parity(e, g, e_ref, g_ref, scale, 1e-10)
        with pytest.raises(RuntimeError):
            q.sample_bitstrings(5)                          # the devices hold the co-state now
        assert abs(q.run_expec_val(betas, gammas) - e_ref) <= 1e-10 * scale
        u = np.concatenate([np.random.RandomState(0).uniform(size=60), [0.0, 0.5]])
        idx = q.sample_bitstrings(u.size, u)
        ref = orc.sample_bitstrings(psi, u)
        cdf = np.cumsum(np.abs(psi) ** 2)
        for a, b, uu in zip(idx, ref, u):
            if a != b:   # only where u sits within rounding of a cdf step
                assert abs(cdf[min(a, b)] - uu) < 1e-13, (a, b, uu)
        assert np.mean(idx == ref) > 0.95
        with pytest.raises(ValueError):
            q.grad_run(betas[:-1], gammas)                  # qaoa.py:186-191
    finally:
        q.close()
    with pytest.raises(ValueError):
        ShardedQaoa(n, {"x": np.array([1.0] + [None] * (n - 1), dtype=object)}, p, LocalComm(G))


def test_swap_engine_needs_enough_local_qubits(backend):
    n, G = 13, 4      # 11 local qubits < 12 + 2
    c = ShardedMcClean(n, mixed_obs(n), 1, LocalComm(G), np.zeros((1, n), int), np.zeros((1, n)))
    assert c.mode == "peer"
    c.close()
    c = ShardedMcClean(n, mixed_obs(n), 1, LocalComm(G), np.zeros((1, n), int), np.zeros((1, n)), mode="swap")
    with pytest.raises(ValueError):
        c.grad_run()
    c.close()


@pytest.mark.parametrize("n,G", [(8, 2), (8, 4), (9, 8), (10, 16)])
def test_diagonal_global_gates_skip_the_exchange(backend, n, G):
    """Rz on a global (rank-bit) qubit is applied as a per-subgroup phase and only X / Y rotations are exchanged
    (QR_OPT_SHARD_ZSKIP): every Z / non-Z pattern on the global qubits against the oracle and against the full exchange."""
    g = int(np.log2(G))
    patterns = [[2] * g, [0] * g, [2] + [1] * (g - 1), [1] * (g - 1) + [2], [(2 if i % 2 else 0) for i in range(g)],
                [(2 if i % 2 == 0 else 1) for i in range(g)]]
    L = len(patterns)
    rng = np.random.default_rng(77 + n)
    axes, angles = rng.integers(0, 3, (L, n)), rng.uniform(0, 2 * np.pi, (L, n))
    for i, pat in enumerate(patterns):
        axes[i, :g] = pat
    obs = mixed_obs(n)
    e_ref, g_ref = orc.mcclean_grad_run(n, obs, axes, angles)
    c = ShardedMcClean(n, obs, L, LocalComm(G), axes, angles)
    try:
        e, gr = c.grad_run()
        assert_parity(e, gr, e_ref, g_ref, obs_scale(obs), 1e-10)
        link_skip = c.link_bytes
        assert abs(c.run_expec_val() - e_ref) <= 1e-10 * obs_scale(obs)
        c.set_option("shard_zskip", 0)
        e0, g0 = c.grad_run()
        assert_parity(e, gr, e0, g0, obs_scale(obs), 1e-12)
        # counted NVLink volume per direction and rank: full exchange = 3 vector-steps per layer x 2 x 16 B x N_loc x (G-1)/G
        n_loc = 2 ** (n - g)
        assert c.link_bytes == pytest.approx(3 * L * 2 * 16.0 * n_loc * (G - 1) / G)
        assert link_skip < c.link_bytes
    finally:
        c.close()


def test_sharded_argument_checks(backend):
    with pytest.raises(ValueError):
        ShardedMcClean(6, mixed_obs(6), 1, LocalComm(3), np.zeros((1, 6), int), np.zeros((1, 6)))
    with pytest.raises(ValueError):
        ShardedMcClean(5, mixed_obs(5), 1, LocalComm(4), np.zeros((1, 5), int), np.zeros((1, 5)))   # < 4 local qubits
    c = ShardedMcClean(8, mixed_obs(8), 1, LocalComm(2), np.full((1, 8), 3), np.zeros((1, 8)))
    with pytest.raises(ValueError):
        c.grad_run()
    c.close()


def test_gloo_world_size_2_sharded_qaoa_on_cpu():
    """Two processes over gloo: ShardedQaoa.grad_run, run_expec_val and the sharded sampler (allreduced shard totals)."""
    script = os.path.join(ROOT, "scripts", "shard_run.py")
    env = dict(os.environ, QR_SHARD_BACKEND="emul", MASTER_ADDR="127.0.0.1", MASTER_PORT="29745")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29745", script, "--circuit", "qaoa", "--qubits", "14", "--layers", "3", "--check"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "PARITY OK" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("qubits,port", [(9, "29741"), (14, "29743")])
def test_gloo_world_size_2_on_cpu(qubits, port):
    """Two processes, torch.distributed gloo, shard buffers mapped across processes.  9 qubits: the peer engine with a host
    barrier after every step; 14 qubits: the swap engine, all steps enqueued at once, ranks ordered by device-side flags."""
    script = os.path.join(ROOT, "scripts", "shard_run.py")
    env = dict(os.environ, QR_SHARD_BACKEND="emul", MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", port, script, "--qubits", str(qubits), "--layers", "3", "--check"]
    res = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "PARITY OK" in res.stdout, res.stdout + res.stderr
