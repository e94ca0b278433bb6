"""Backend selection for the parity tests.

`cuda`  : the product library qradient_b200/libqradient_b200.so on a real GPU (marker: gpu).
`emul`  : the same kernel sources compiled by g++ against tests/emul/cuda_emul.h, so the CPU-only
          tier executes every kernel's thread program (test infrastructure, never shipped).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))

from qradient_b200 import _lib  # noqa: E402

def _cuda_devices():
    """Number of visible CUDA devices (0 without a driver): plain `pytest tests` on a CPU-only box skips the cuda
    backend instead of failing in qr_ctx_create."""
    import ctypes
    try:
        n = ctypes.c_int(0)
        rc = _lib.Library(_lib.LIB_PATH).cdll.qr_device_count(ctypes.byref(n))
        return n.value if rc == 0 else 0
    except Exception:
        return 0


_NO_GPU = pytest.mark.skipif(_cuda_devices() == 0, reason="no CUDA device visible")
BACKENDS = [pytest.param("emul", id="emul"), pytest.param("cuda", id="cuda", marks=[pytest.mark.gpu, _NO_GPU])]
_handles = {}


def activate(name):
    if name not in _handles:
        if name == "emul":
            import build_emul
            _handles[name] = _lib.Library(build_emul.build())
        else:
            _handles[name] = _lib.Library(_lib.LIB_PATH)
    _lib._restore(_handles[name])
    return _handles[name]


@pytest.fixture(params=BACKENDS)
def backend(request):
    prev = _lib._LIB
    activate(request.param)
    yield request.param
    _lib._restore(prev)
