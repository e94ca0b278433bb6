"""ctypes binding of libqradient_b200.so (include/qradient_b200.h).

There is no CPU fallback: if the CUDA library is missing or no GPU is visible, importing a
circuit or creating a State raises.  `_load_for_testing` exists only so the CPU-only test tier
can execute the kernels' thread programs through tests/emul (never used by the package itself).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QRADIENT_B200_LIB", os.path.join(_HERE, "libqradient_b200.so"))   # override: A/B timing of library builds

QR_OK, QR_EINVAL, QR_ECUDA, QR_ENOMEM, QR_ESTATE = 0, 1, 2, 3, 4
TERM_KIND = {"x": 0, "y": 1, "z": 2, "zz": 3}
OPT = {"fusion": 0, "tile_bits": 1, "prefetch": 2, "ctas_per_sm_fwd": 3, "ctas_per_sm_bwd": 4,
       "final_ladder": 5, "ham_lut": 6, "tile_bits_strided": 11, "min_row_bits": 12, "batch_chunk_mb": 13,
       "staged": 19, "staged_min_bit": 20, "pdl": 26, "defer_reduce": 28, "shard_zskip": 29, "shard_mode": 30, "shard_lockstep": 31, "shard_slices": 32, "shard_xsms": 33, "axis_plan": 34, "loop_graph": 35}

c_int, c_double, c_void_p, c_size_t, c_ll = ctypes.c_int, ctypes.c_double, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_longlong
P = ctypes.POINTER


class QrPerf(ctypes.Structure):
    _fields_ = [("ms_total", c_double), ("ms_forward", c_double), ("ms_observable", c_double),
                ("ms_backward", c_double), ("algorithmic_bytes", c_double), ("bwd_pass_ms_avg", c_double),
                ("bwd_pass_bytes", c_double), ("fwd_pass_ms_avg", c_double), ("fwd_pass_bytes", c_double),
                ("kernel_launches", c_ll), ("passes_per_layer", c_int), ("tile_bits", c_int), ("link_bytes", c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_SIGNATURES = {
    "qr_last_error": (ctypes.c_char_p, []),
    "qr_version": (c_int, []),
    "qr_device_count": (c_int, [P(c_int)]),
    "qr_ctx_create": (c_int, [c_int, c_int, P(c_void_p)]),
    "qr_ctx_destroy": (c_int, [c_void_p]),
    "qr_set_option": (c_int, [c_void_p, c_int, c_ll]),
    "qr_get_option": (c_int, [c_void_p, c_int, P(c_ll)]),
    "qr_perf_last": (c_int, [c_void_p, P(QrPerf)]),
    "qr_sync": (c_int, [c_void_p]),
    "qr_state_init": (c_int, [c_void_p, c_int]),
    "qr_state_upload": (c_int, [c_void_p, c_void_p, c_size_t]),
    "qr_state_download": (c_int, [c_void_p, c_void_p, c_size_t]),
    "qr_state_device_ptr": (c_int, [c_void_p, P(c_void_p)]),
    "qr_state_save": (c_int, [c_void_p, c_int]),
    "qr_state_load": (c_int, [c_void_p, c_int]),
    "qr_state_free_snapshots": (c_int, [c_void_p]),
    "qr_apply_rot": (c_int, [c_void_p, c_int, c_double, c_int]),
    "qr_apply_drot": (c_int, [c_void_p, c_int, c_double, c_int]),
    "qr_apply_cnot": (c_int, [c_void_p, c_int, c_int]),
    "qr_apply_cnot_ladder": (c_int, [c_void_p, c_int, c_int]),
    "qr_apply_x_summed": (c_int, [c_void_p]),
    "qr_norm2": (c_int, [c_void_p, P(c_double)]),
    "qr_obs_create": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, P(c_void_p)]),
    "qr_obs_destroy": (c_int, [c_void_p]),
    "qr_apply_observable": (c_int, [c_void_p, c_void_p]),
    "qr_expec_val": (c_int, [c_void_p, c_void_p, P(c_double)]),
    "qr_term_expecs": (c_int, [c_void_p, c_void_p, c_void_p]),
    "qr_ham_load": (c_int, [c_void_p, c_void_p]),
    "qr_ham_download": (c_int, [c_void_p, c_void_p, c_size_t]),
    "qr_apply_exp_ham": (c_int, [c_void_p, c_double]),
    "qr_apply_exp_ham_component": (c_int, [c_void_p, c_void_p, c_int, c_double]),
    "qr_apply_ham": (c_int, [c_void_p, c_int]),
    "qr_mcclean_expec": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, P(c_double)]),
    "qr_mcclean_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, P(c_double), c_void_p]),
    "qr_layered_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, P(c_double), c_void_p]),
    "qr_mcclean_grad_batch": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "qr_qaoa_expec": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, P(c_double)]),
    "qr_qaoa_grad": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, P(c_double), c_void_p]),
    "qr_sample_bitstrings": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "qr_ham_gather": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "qr_mcclean_optimize": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                    c_void_p, c_void_p]),
    "qr_qaoa_optimize": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "qr_perm_load": (c_int, [c_void_p, c_void_p, c_size_t]),
    "qr_state_permute": (c_int, [c_void_p]),
    "qr_dense_load": (c_int, [c_void_p, c_void_p, c_size_t]),
    "qr_state_apply_dense": (c_int, [c_void_p]),
    "qr_shard_create": (c_int, [c_int, c_int, c_int, c_int, P(c_void_p)]),
    "qr_shard_ipc_handle": (c_int, [c_void_p, c_int, c_void_p]),
    "qr_shard_ipc_open": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "qr_shard_set_peer_ptr": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int]),
    "qr_shard_buffer_ptr": (c_int, [c_void_p, c_int, P(c_void_p)]),
    "qr_shard_info": (c_int, [c_void_p, P(c_int), P(c_int), P(c_int), P(c_int)]),
    "qr_shard_mcclean_begin": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, P(c_int)]),
    "qr_shard_qaoa_begin": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, P(c_int)]),
    "qr_shard_step": (c_int, [c_void_p, c_int]),
    "qr_shard_mcclean_finish": (c_int, [c_void_p, P(c_double), c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


class Library:
    """A loaded libqradient_b200 with typed entry points and error translation."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise RuntimeError(
                "qradient_b200: CUDA library %s not found. Build it with `python -m qradient_b200.build` "
                "(needs nvcc); there is no CPU fallback." % path)
        self.path = path
        self.cdll = ctypes.CDLL(path)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(self.cdll, name)
            fn.restype, fn.argtypes = res, args

    def last_error(self):
        return self.cdll.qr_last_error().decode("utf-8", "replace")

    def check(self, rc):
        if rc == QR_OK:
            return
        msg = self.last_error()
        if rc == QR_EINVAL:
            raise ValueError(msg)
        if rc == QR_ENOMEM:
            raise MemoryError(msg)
        raise RuntimeError("qradient_b200: " + msg)

    def call(self, name, *args):
        self.check(getattr(self.cdll, name)(*args))


_LIB = None


def lib():
    """The process-wide library handle (loaded on first use)."""
    global _LIB
    if _LIB is None:
        _LIB = Library(LIB_PATH)
    return _LIB


def _load_for_testing(path):
    """Swap in another build of the same C ABI (tests/emul only).  Returns the previous handle."""
    global _LIB
    prev = _LIB
    _LIB = Library(path) if path is not None else None
    return prev


def _restore(handle):
    global _LIB
    _LIB = handle


def ptr(a):
    return a.ctypes.data_as(c_void_p)


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
