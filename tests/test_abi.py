"""The C-ABI library: it loads, and exports every symbol include/qradient_b200.h declares.
No compute calls here (no GPU in the CPU tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "qradient_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qr_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    from qradient_b200 import build
    return build.build()


def test_header_declares_the_bound_symbols():
    from qradient_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == header_symbols()


def test_cuda_library_exports_every_header_symbol(built_lib):
    handle = ctypes.CDLL(built_lib)
    for sym in header_symbols():
        assert hasattr(handle, sym), sym


def test_library_is_sm100a_cuda_code(built_lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_gpu_means_loud_failure_not_fallback(built_lib):
    """Without a device, creating a context must raise (QR_ECUDA), never compute on the CPU."""
    from qradient_b200 import _lib
    lib = _lib.Library(built_lib)
    n = ctypes.c_int(-1)
    rc = lib.cdll.qr_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible")
    h = ctypes.c_void_p()
    with pytest.raises(RuntimeError):
        lib.call("qr_ctx_create", 4, 0, ctypes.byref(h))


def test_product_package_never_references_oracle_or_emulation():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "qradient_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "qr_oracle" not in text and "import oracle" not in text and "from oracle" not in text, f
                if f.endswith(".py"):
                    assert "libqr_emul" not in text and "build_emul" not in text, f
