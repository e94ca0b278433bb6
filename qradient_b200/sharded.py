"""State-vector sharding over several GPUs for registers that do not fit one device.

The reference is single-process; this is the large-register extension named in BASELINE.json
(McClean 33 qubits on 8 GPUs).  Rank r holds the amplitudes whose top log2(G) index bits equal r
(qubits 0..log2(G)-1 are the "global" qubits, physical_components/state.py:84-88 bit order).

    * local qubits  : the same fused tile passes as on one GPU;
    * CNOT ladder   : a whole destination shard reads exactly one source shard, so the ladder is a
                      relabelling of shards plus a local gather -- no data movement;
    * global qubits : "swap" engine (default; >= 12 + log2 G local qubits): an EXCHANGE PASS -- an ordinary tile pass
                      whose loads come from the G peer shards over NVLink (CUDA IPC mappings) and whose stores are
                      local, so the rank bits trade places with g local bits and every exchanged amplitude crosses the
                      link once; the passes of all layers are enqueued at once and device-side flags order the ranks.
                      "peer" engine (small registers): one kernel per layer and vector reads the G peer shards,
                      applies the rotations in registers and writes them back; a host barrier after every step;
    * E and gradient: per-rank partial sums, one allreduce of L*n+1 doubles at the end.

Two communicators: `TorchDistComm` (one process per GPU, torch.distributed: NCCL on GPUs, gloo in
the CPU test tier) and `LocalComm` (all shards driven by one process; used by tests and by
single-process multi-GPU runs).
"""
import ctypes

import numpy as np

from . import _lib
from .physical_components import Observable

NBUF = 4          # state buffers per shard; index NBUF addresses the shard's ordering flags (swap engine)


class _Shard:
    """One rank's shard context."""

    def __init__(self, lib, n_total, log2_world, rank, device):
        self.lib, self.rank, self.device = lib, rank, device
        h = ctypes.c_void_p()
        lib.call('qr_shard_create', int(n_total), int(log2_world), int(rank), int(device), ctypes.byref(h))
        self.ctx = h

    def handles(self):
        out = []
        for b in range(NBUF + 1):
            buf = ctypes.create_string_buffer(64)
            self.lib.call('qr_shard_ipc_handle', self.ctx, b, buf)
            out.append(buf.raw)
        return out

    def pointers(self):
        out = []
        for b in range(NBUF + 1):
            p = ctypes.c_void_p()
            self.lib.call('qr_shard_buffer_ptr', self.ctx, b, ctypes.byref(p))
            out.append(p.value)
        return out

    def set_option(self, name, value):
        self.lib.call('qr_set_option', self.ctx, _lib.OPT[name], int(value))

    def close(self):
        if self.ctx is not None:
            self.lib.cdll.qr_ctx_destroy(self.ctx)
            self.ctx = None


class TorchDistComm:
    """One process per GPU; torch.distributed must be initialised by the caller."""
    lockstep = False   # the swap engine orders the ranks with device-side flags: no host barrier between steps

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)

    def local_ranks(self):
        return [self.rank]

    def exchange_handles(self, shards):
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, shards[0].handles(), group=self.group)
        for peer, hs in enumerate(gathered):
            if peer == self.rank:
                continue
            for b, h in enumerate(hs):
                shards[0].lib.call('qr_shard_ipc_open', shards[0].ctx, peer, b, h)

    def barrier(self):
        self.dist.barrier(group=self.group)

    def allreduce_sum(self, arrays):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arrays[0]))
        backend = self.dist.get_backend(self.group)
        if backend == 'nccl':
            tc = t.cuda()
            self.dist.all_reduce(tc, group=self.group)
            t = tc.cpu()
        else:
            self.dist.all_reduce(t, group=self.group)
        return t.numpy()


class LocalComm:
    """All G shards live in this process (device d per shard, or one device for tests)."""
    lockstep = True    # the shards are stepped one after another by this process

    def __init__(self, world, devices=None):
        self.world = world
        self.devices = devices if devices is not None else [0] * world
        self.rank = 0

    def local_ranks(self):
        return list(range(self.world))

    def exchange_handles(self, shards):
        ptrs = [s.pointers() for s in shards]
        for s in shards:
            for peer in range(self.world):
                if peer == s.rank:
                    continue
                for b in range(NBUF + 1):
                    s.lib.call('qr_shard_set_peer_ptr', s.ctx, peer, b, ctypes.c_void_p(ptrs[peer][b]), self.devices[peer])

    def barrier(self):
        pass   # every step is stream-synchronised and the shards are stepped in lockstep

    def allreduce_sum(self, arrays):
        return np.sum(arrays, axis=0)


class ShardedMcClean:
    """McClean circuit (circuit_logic/mc_clean.py) on a register sharded over G = 2^g GPUs.

    Same constructor arguments and return values as `McClean.grad_run` / `run_expec_val`; every
    rank must call the methods collectively with identical arguments.
    """

    def __init__(self, qubit_number, observable, layer_number, comm, axes, angles, device=None, mode=None):
        self.qnum, self.lnum, self.comm = int(qubit_number), int(layer_number), comm
        g = int(np.log2(comm.world))
        if 2 ** g != comm.world or g < 1:
            raise ValueError('world size must be a power of two >= 2')
        self.log2_world = g
        self._lib = _lib.lib()
        self.observable = Observable(self.qnum, observable)
        self.axes, self.angles = axes, angles
        self.shards = []
        for r in comm.local_ranks():
            dev = device if device is not None else (comm.devices[r] if hasattr(comm, 'devices') else 0)
            self.shards.append(_Shard(self._lib, self.qnum, g, r, dev))
        comm.barrier()
        comm.exchange_handles(self.shards)
        comm.barrier()
        nl = self.qnum - g
        if mode is None:
            mode = 'swap' if (1 <= g <= 3 and nl >= 12 + g) else 'peer'
        if mode not in ('swap', 'peer'):
            raise ValueError("mode must be 'swap' or 'peer'")
        self.mode = mode
        self.set_option('shard_mode', 2 if mode == 'swap' else 1)
        self.set_option('shard_lockstep', 1 if getattr(comm, 'lockstep', True) else 0)
        self.perf = {}

    def set_option(self, name, value):
        for s in self.shards:
            s.set_option(name, value)

    def _params(self):
        axes = np.asarray(self.axes)
        angles = np.asarray(self.angles, dtype=np.float64)
        if axes.shape != (self.lnum, self.qnum) or angles.shape != (self.lnum, self.qnum):
            raise ValueError('axes and angles must have shape ({}, {})'.format(self.lnum, self.qnum))
        if np.any((axes < 0) | (axes > 2)):
            raise ValueError('Invalid axis')
        return _lib.as_i32(axes), np.ascontiguousarray(angles)

    def _run(self, want_grad):
        axes, angles = self._params()
        nsteps = ctypes.c_int()
        for s in self.shards:
            self._lib.call('qr_shard_mcclean_begin', s.ctx, self.lnum, _lib.ptr(axes), _lib.ptr(angles),
                           self.observable._handle, int(want_grad), ctypes.byref(nsteps))
        self.comm.barrier()
        import time
        L = self.lnum
        if self.mode == 'swap':
            # swap engine: one process per GPU enqueues every step at once (device-side flags order the ranks); a process
            # that drives several shards runs them in lockstep, one step after another
            t0 = time.perf_counter()
            lockstep = getattr(self.comm, 'lockstep', True)
            for step in range(nsteps.value):
                for s in self.shards:
                    self._lib.call('qr_shard_step', s.ctx, step)
                if lockstep:
                    self.comm.barrier()
            self.step_seconds = {'enqueue': time.perf_counter() - t0}
        else:
            # peer engine: every step is stream-synchronised and followed by a barrier over the ranks
            self.step_seconds = {'fwd_local': 0.0, 'fwd_global': 0.0, 'observable': 0.0, 'bwd_local': 0.0, 'bwd_global': 0.0}
            for step in range(nsteps.value):
                t0 = time.perf_counter()
                for s in self.shards:
                    self._lib.call('qr_shard_step', s.ctx, step)
                self.comm.barrier()
                if step < 2 * L:
                    kind = 'fwd_local' if step % 2 == 0 else 'fwd_global'
                elif step == 2 * L:
                    kind = 'observable'
                else:
                    kind = 'bwd_local' if (step - 2 * L - 1) % 2 == 0 else 'bwd_global'
                self.step_seconds[kind] += time.perf_counter() - t0   # wall time of this rank incl. the barrier
        parts = []
        for s in self.shards:
            e = ctypes.c_double()
            grad = np.zeros(self.lnum * self.qnum + 1, dtype=np.float64)
            self._lib.call('qr_shard_mcclean_finish', s.ctx, ctypes.byref(e), _lib.ptr(grad[1:]) if want_grad else None)
            grad[0] = e.value
            parts.append(grad)
        # NVLink volume of this gradient as counted by the library (peer loads + peer stores of the first local shard):
        # smaller than the full exchange when global qubits carry Rz gates (QR_OPT_SHARD_ZSKIP)
        perf = _lib.QrPerf()
        self._lib.call('qr_perf_last', self.shards[0].ctx, ctypes.byref(perf))
        self.link_bytes = perf.link_bytes
        self.perf = {'kernel_launches': int(perf.kernel_launches), 'link_bytes': float(perf.link_bytes),
                     'sweeps_per_layer': int(perf.passes_per_layer) + (1 if self.mode == 'peer' else 0), 'mode': self.mode}
        if self.mode == 'swap':   # device times of rank 0's stream: local passes / exchange passes / waiting for peers
            nxf, nxb = perf.fwd_pass_bytes, perf.bwd_pass_bytes
            busy = perf.ms_forward + perf.ms_backward + perf.ms_observable + nxf * perf.fwd_pass_ms_avg + nxb * perf.bwd_pass_ms_avg
            self.perf.update({'ms_total': perf.ms_total, 'ms_local_forward': perf.ms_forward, 'ms_local_backward': perf.ms_backward,
                              'ms_observable': perf.ms_observable, 'ms_exchange_forward_avg': perf.fwd_pass_ms_avg,
                              'ms_exchange_backward_avg': perf.bwd_pass_ms_avg, 'exchange_passes': int(nxf + nxb),
                              'ms_waiting_and_flags': perf.ms_total - busy})
            self.step_seconds.update({k: v for k, v in self.perf.items() if k.startswith('ms_')})
        total = self.comm.allreduce_sum(parts)
        return float(total[0]), np.array(total[1:]).reshape(self.lnum, self.qnum)

    def run_expec_val(self):
        return self._run(False)[0]

    def grad_run(self):
        return self._run(True)

    def close(self):
        for s in self.shards:
            s.close()
        self.shards = []


class ShardedQaoa:
    """QAOA circuit (circuit_logic/qaoa.py) on a register sharded over G = 2^g GPUs: `run_expec_val`, `grad_run` and
    bitstring sampling with the signatures and return values of `Qaoa`; every rank calls the methods collectively with
    identical arguments.  Swap engine only (>= 12 + log2 G local qubits); the observable has z / zz terms (qaoa.py:11-14)."""

    def __init__(self, qubit_number, observable, layer_number, comm, device=None):
        self.qnum, self.lnum, self.comm = int(qubit_number), int(layer_number), comm
        g = int(np.log2(comm.world))
        if 2 ** g != comm.world or g < 1:
            raise ValueError('world size must be a power of two >= 2')
        self.log2_world = g
        self._lib = _lib.lib()
        self.observable = Observable(self.qnum, observable)
        if np.any(self.observable.term_kinds < 2):
            raise ValueError('the classical Hamiltonian of a QAOA circuit has z / zz terms only')
        self.shards = []
        for r in comm.local_ranks():
            dev = device if device is not None else (comm.devices[r] if hasattr(comm, 'devices') else 0)
            self.shards.append(_Shard(self._lib, self.qnum, g, r, dev))
        comm.barrier()
        comm.exchange_handles(self.shards)
        comm.barrier()
        self.mode = 'swap'
        for s in self.shards:
            s.set_option('shard_mode', 2)
            s.set_option('shard_lockstep', 1 if getattr(comm, 'lockstep', True) else 0)
        self.perf = {}
        self._sampled_state = False

    def set_option(self, name, value):
        for s in self.shards:
            s.set_option(name, value)

    def _check_parameters(self, betas, gammas):
        betas, gammas = np.asarray(betas, dtype=np.float64), np.asarray(gammas, dtype=np.float64)
        if (betas.size != self.lnum) or (gammas.size != self.lnum):   # qaoa.py:186-191
            raise ValueError('Wrong amount of parameters. Expected {0} and {0}, found {1} and {2}.'.format(
                self.lnum, betas.size, gammas.size))
        return np.ascontiguousarray(betas).ravel(), np.ascontiguousarray(gammas).ravel()

    def _run(self, betas, gammas, want_grad):
        betas, gammas = self._check_parameters(betas, gammas)
        nsteps = ctypes.c_int()
        for s in self.shards:
            self._lib.call('qr_shard_qaoa_begin', s.ctx, self.lnum, _lib.ptr(betas), _lib.ptr(gammas), self.observable._handle,
                           int(want_grad), ctypes.byref(nsteps))
        self.comm.barrier()
        lockstep = getattr(self.comm, 'lockstep', True)
        for step in range(nsteps.value):
            for s in self.shards:
                self._lib.call('qr_shard_step', s.ctx, step)
            if lockstep:
                self.comm.barrier()
        parts = []
        for s in self.shards:
            e = ctypes.c_double()
            grad = np.zeros(2 * self.lnum + 1, dtype=np.float64)
            self._lib.call('qr_shard_mcclean_finish', s.ctx, ctypes.byref(e), _lib.ptr(grad[1:]) if want_grad else None)
            grad[0] = e.value
            parts.append(grad)
        perf = _lib.QrPerf()
        self._lib.call('qr_perf_last', self.shards[0].ctx, ctypes.byref(perf))
        self.link_bytes = perf.link_bytes
        self.perf = {'kernel_launches': int(perf.kernel_launches), 'link_bytes': float(perf.link_bytes),
                     'sweeps_per_layer': int(perf.passes_per_layer), 'mode': self.mode, 'ms_total': perf.ms_total}
        self._sampled_state = not want_grad
        total = self.comm.allreduce_sum(parts)
        return float(total[0]), np.array(total[1:]).reshape(self.lnum, 2)

    def run_expec_val(self, betas, gammas, hide_progbar=True, exact_expec_val=True, shot_num=1, ini_state=None):
        """qaoa.py:23-38 (exact expectation value; the state stays on the devices for `sample_bitstrings`)."""
        if ini_state is not None or not exact_expec_val:
            raise NotImplementedError('sharded QAOA: exact expectation values from |+..+> only')
        return self._run(betas, gammas, False)[0]

    def grad_run(self, betas, gammas, hide_progbar=True, ini_state=None):
        """qaoa.py:40-70: (E, grad[p, 2]), column 0 = d/d beta, column 1 = d/d gamma."""
        if ini_state is not None:
            raise NotImplementedError('sharded QAOA starts from |+..+>')
        return self._run(betas, gammas, True)

    def sample_bitstrings(self, shot_num, uniforms=None):
        """Indices drawn from |psi|^2 of the LAST run_expec_val by inverse CDF (qaoa.py:196-198): first k with
        cumsum(|psi|^2)[k] >= u.  Every shard scans its own amplitudes; the shard totals are gathered, the shard whose range
        of the cumulative sum holds u searches u minus the total of the shards before it, and the global index is the
        shard's base plus the local one."""
        if not self._sampled_state:
            raise RuntimeError('sample_bitstrings needs the state of a preceding run_expec_val (grad_run leaves the co-state)')
        u = np.random.uniform(size=shot_num) if uniforms is None else np.asarray(uniforms, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        G, nl = self.comm.world, self.qnum - self.log2_world
        totals = np.zeros(G)
        for s in self.shards:
            t = ctypes.c_double()
            self._lib.call('qr_norm2', s.ctx, ctypes.byref(t))
            totals[s.rank] = t.value
        totals = np.asarray(self.comm.allreduce_sum([totals]))
        ends = np.cumsum(totals)
        owner = np.searchsorted(ends, u, side='left')          # first shard whose cumulative total reaches u
        out = np.zeros(u.size, dtype=np.float64)
        for s in self.shards:
            mine = np.nonzero(owner == s.rank)[0]
            if mine.size == 0:
                continue
            base = ends[s.rank] - totals[s.rank]
            ul = np.ascontiguousarray(np.minimum(u[mine] - base, totals[s.rank]))
            idx = np.empty(mine.size, dtype=np.int64)
            self._lib.call('qr_sample_bitstrings', s.ctx, int(mine.size), _lib.ptr(ul), _lib.ptr(idx))
            out[mine] = idx + float(s.rank) * 2.0 ** nl
        out = np.asarray(self.comm.allreduce_sum([out]))       # u above the last total: no shard owns it, index 0 (scipy's argmax)
        return out.astype(np.int64)

    def close(self):
        for s in self.shards:
            s.close()
        self.shards = []
