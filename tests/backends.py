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

BACKENDS = [pytest.param("emul", id="emul"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]
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
